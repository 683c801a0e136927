# Round-2, GPU call 17 (1 GPU): antialias kernels vs the CPU oracle + antialiased stage-1 step; TV-inside-the-fused-backward parity cases (run 16 never
# reached the GPU: the session that prepared it was cut off) and its A/B.
set -x
timeout 900 python -m pytest tests/test_gpu_antialias.py tests/test_gpu_stage1.py tests/test_gpu_raster.py -m gpu -q --durations=8 > gpurun_out/t_aa.log 2>&1; tail -16 gpurun_out/t_aa.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_aa.log | cut -c1-400 | head -20
timeout 1200 python -m pytest tests/test_gpu_reference_parity.py -m gpu -q --durations=8 -k "tvfused or tv_gradient" > gpurun_out/t_tv.log 2>&1; tail -14 gpurun_out/t_tv.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_tv.log | cut -c1-400 | head
for extra in "--tv-in-bwd 1" "--tv-in-bwd 0" "--tv-in-bwd 1" "--tv-in-bwd 0"; do
timeout 600 python bench.py --steps 100 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
done
