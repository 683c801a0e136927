set -x
timeout 900 python -m pytest tests/test_gpu_reference_parity.py -q -k "fused_step and garden" > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-300 | head -10
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_parity.py > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
