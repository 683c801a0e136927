# Round-2, GPU call 21 (1 GPU): stage-1 step as a CUDA graph (test + bench A/B), default stage-0 bench line with the eval_render leg.
set -x
timeout 900 python -m pytest tests/test_gpu_stage1.py -m gpu -q -x > gpurun_out/t_s1.log 2>&1; tail -3 gpurun_out/t_s1.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_s1.log | cut -c1-400 | head
for extra in "--antialias 1" "--antialias 1 --no-graph" "--antialias 0" "--antialias 0 --no-graph" "--antialias 1"; do
timeout 600 python bench.py --workload lego_stage1 --steps 100 --warmup 20 $extra > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; python -c "
import json
for l in open('gpurun_out/bench_s1.json'):
    if l.startswith('{'):
        d=json.loads(l); print('S1 $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", d['antialias_ms'], round(d['forward_ms'],4))"; tail -2 gpurun_out/bench_s1.err
done
timeout 900 python bench.py --skip-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
