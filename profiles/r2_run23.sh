# Round-2, GPU call 23 (1 GPU): rasterizer with camera-plane-crossing triangles (tests), final ncu launch list + full capture.
set -x
timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_stage1.py -m gpu -q -x > gpurun_out/t_raster.log 2>&1; tail -4 gpurun_out/t_raster.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_raster.log | cut -c1-500 | head
bash profiles/r2_ncu_final.sh
