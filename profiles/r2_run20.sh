# Round-2, GPU call 20 (1 GPU): marching cubes vs the oracle + 512^3 properties; render tests after the chunk / counters change; stage-1 bench in both orders.
set -x
timeout 900 python -m pytest tests/test_gpu_mcubes.py -m gpu -q -s --durations=6 > gpurun_out/t_mc.log 2>&1; tail -14 gpurun_out/t_mc.log; grep -E "marching cubes 512|^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_mc.log | cut -c1-400 | head
timeout 900 python -m pytest tests/test_gpu_stage0.py tests/test_gpu_reference_parity.py -m gpu -q -x -k "render" > gpurun_out/t_render.log 2>&1; tail -3 gpurun_out/t_render.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_render.log | cut -c1-400 | head
for aa in 0 1 0 1; do
timeout 600 python bench.py --workload lego_stage1 --steps 100 --warmup 10 --antialias $aa > gpurun_out/bench_s1_aa$aa.json 2> gpurun_out/bench_s1_aa$aa.err; python -c "
import json
for l in open('gpurun_out/bench_s1_aa$aa.json'):
    if l.startswith('{'):
        d=json.loads(l); print('S1 aa=$aa |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", d['antialias_ms'], round(d['forward_ms'],4))"; tail -2 gpurun_out/bench_s1_aa$aa.err
done
