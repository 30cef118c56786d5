# round-2 call S (1 GPU, under a minute): the quick tests outside the GEMM selection of call O that reach the one-pass prefill, with the tile-order planes as the default
timeout 55 python -m pytest tests -m gpu -q -x --timeout 50 -k "generate_stops or greedy_stream_identical or long_context or set_gamma_then or engine_device_sampling" 2>&1 | tail -3
