# Round-2, GPU call 6: fused backward v2 with the corrected register budget, fused forward test alone, garden TV diagnosis, stage-1 bench.
set -x
timeout 300 python profiles/fusedprobe.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_stage0.py -q -k "fused_forward or tv_random" > gpurun_out/t_fused_fwd.log 2>&1; tail -3 gpurun_out/t_fused_fwd.log; grep -E "^E  " gpurun_out/t_fused_fwd.log | cut -c1-300 | head -5
timeout 900 python -m pytest tests/test_gpu_stage0.py -q -k "fused_backward" > gpurun_out/t_fused_bwd.log 2>&1; tail -3 gpurun_out/t_fused_bwd.log; grep -E "^E  " gpurun_out/t_fused_bwd.log | cut -c1-300 | head -5
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -q > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-400 | head -30
timeout 600 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_raster.py -q > gpurun_out/t_stage1.log 2>&1; tail -3 gpurun_out/t_stage1.log
timeout 600 python bench.py --steps 60 --warmup 10 --fused-bwd 1 --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG fused-bwd |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
timeout 600 python bench.py --workload lego_stage1 --steps 20 --warmup 5 > gpurun_out/bench_stage1.json 2> gpurun_out/bench_stage1.err; tail -c 1500 gpurun_out/bench_stage1.json; tail -3 gpurun_out/bench_stage1.err
