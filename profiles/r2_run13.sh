# Round-2, GPU call 13 (1 GPU): gpu suite after the scan-round / index-wrap changes; where to release the next batch's march (A/B); garden.
set -x
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -6 gpurun_out/t_all.log; grep -E "^E  .*(Assertion|assert )|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head
for extra in "--prefetch-at optimizer" "--prefetch-at start" "--prefetch-at optimizer" "--prefetch-at start" "--prefetch-at optimizer --defer-zero 0"; do
timeout 600 python bench.py --steps 100 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
done
for extra in "--prefetch-at optimizer" "--prefetch-at start"; do
timeout 600 python bench.py --workload garden_stage0 --steps 40 --warmup 10 --skip-cpu --skip-reference $extra > gpurun_out/bench_garden.json 2> gpurun_out/bench_garden.err; python -c "
import json
for l in open('gpurun_out/bench_garden.json'):
    if l.startswith('{'):
        d=json.loads(l); print('GARDEN $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_garden.err
done
