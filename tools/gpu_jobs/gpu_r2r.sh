# round-2 call R (1 GPU, seconds): smoke() as the driver runs it, with the one-pass prefill step
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
