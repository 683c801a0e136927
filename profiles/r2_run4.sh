# Round-2, GPU call 4: parity (full logs), fused-backward role probe, rasterizer + stage-1 tests.
set -x
timeout 300 python profiles/fusedprobe.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_stage1.py -q -x > gpurun_out/t_raster.log 2>&1; tail -15 gpurun_out/t_raster.log
timeout 1200 python -m pytest tests/test_gpu_reference_parity.py -q > gpurun_out/t_parity.log 2>&1; grep -E "^E  |passed|failed" gpurun_out/t_parity.log | head -40
timeout 900 python -m pytest tests/test_gpu_stage0.py -q -k "fused_backward" > gpurun_out/t_fused.log 2>&1; tail -5 gpurun_out/t_fused.log
