# Round-2, GPU call 8: TV-alone parity, parts x fused combinations, stage-1 bench at ~40 % coverage, full gpu test suite.
set -x
timeout 900 python -m pytest tests/test_gpu_reference_parity.py -q -k "tv_gradient_alone or garden or mark_untrained" > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-500 | head -30
for extra in "--parts 1" "--parts 2" "--parts 4" "--parts 1 --fused-fwd 1"; do
timeout 600 python bench.py --steps 60 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
done
timeout 600 python bench.py --workload lego_stage1 --steps 20 --warmup 5 > gpurun_out/bench_stage1.json 2> gpurun_out/bench_stage1.err; tail -c 1200 gpurun_out/bench_stage1.json; tail -3 gpurun_out/bench_stage1.err
timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_parity.py > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
