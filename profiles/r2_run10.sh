# Round-2, GPU call 10 (1 GPU): full gpu test suite, defer-zero A/B, ncu launch list + full capture, final default bench with all legs.
set -x
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -6 gpurun_out/t_all.log; grep -E "^E  .*(Assertion|assert )|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head
for extra in "--defer-zero 1" "--defer-zero 0" "--defer-zero 1" "--defer-zero 0"; do
timeout 600 python bench.py --steps 100 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4))"; tail -2 gpurun_out/bench_x.err
done
bash profiles/r2_ncu.sh
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 4000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload garden_stage0 --steps 40 --warmup 10 --skip-cpu > gpurun_out/bench_garden.json 2> gpurun_out/bench_garden.err; tail -c 600 gpurun_out/bench_garden.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
