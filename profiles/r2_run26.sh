# Round-2, GPU call 26 (1 GPU, last minutes of the budget): render tests after the reciprocal-direction change of k_r_march.
timeout 100 python -m pytest tests/test_gpu_stage0.py tests/test_gpu_reference_parity.py -m gpu -q -x -k "render" 2>&1 | tail -3
