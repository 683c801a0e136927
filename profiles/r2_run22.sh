# Round-2, GPU call 22 (1 GPU): vertex-offset optimizer of stage 1 (test), stage-1 suite.
set -x
timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_antialias.py -m gpu -q > gpurun_out/t_s1.log 2>&1; tail -4 gpurun_out/t_s1.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_s1.log | cut -c1-500 | head
