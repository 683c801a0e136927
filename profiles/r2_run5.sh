# Round-2, GPU call 5b: fused backward v2 + fused forward, parity diagnostics, mark_untrained, TV fallback, new bench.py on three workloads.
set -x
timeout 300 python profiles/fusedprobe.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_stage0.py -q -k "fused_backward or fused_forward or tv_random" > gpurun_out/t_fused.log 2>&1; tail -3 gpurun_out/t_fused.log; grep -E "^E  " gpurun_out/t_fused.log | cut -c1-300 | head
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -q > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-400 | head -30
for extra in "--fused-bwd 0" "--fused-bwd 1" "--fused-fwd 1" "--fused-fwd 1 --fused-bwd 1" "--parts 1"; do
timeout 600 python bench.py --steps 60 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
done
timeout 900 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload garden_stage0 --steps 40 --warmup 10 --skip-cpu > gpurun_out/bench_garden.json 2> gpurun_out/bench_garden.err; tail -c 2500 gpurun_out/bench_garden.json; tail -3 gpurun_out/bench_garden.err
timeout 600 python bench.py --workload lego_stage1 --steps 20 --warmup 5 > gpurun_out/bench_stage1.json 2> gpurun_out/bench_stage1.err; tail -c 1500 gpurun_out/bench_stage1.json; tail -3 gpurun_out/bench_stage1.err
