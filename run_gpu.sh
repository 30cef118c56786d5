timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "forward_logits or greedy_stream or decode_modes or sampling or wide_tier or tier_greedy" 2>&1 | tail -4
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
