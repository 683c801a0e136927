# Round-2, GPU call 16 (1 GPU): TV gradient inside the fused backward kernel -- parity (vs the stand-alone TV launch and vs the reference) and A/B.
set -x
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py tests/test_gpu_stage0.py -m gpu -q -x -k "fused_step or tv_gradient or stage0" > gpurun_out/t_tv.log 2>&1; tail -4 gpurun_out/t_tv.log; grep -E "^E  .*(Assertion|assert )|^FAILED" gpurun_out/t_tv.log | cut -c1-300 | head
for extra in "--tv-in-bwd 1" "--tv-in-bwd 0" "--tv-in-bwd 1" "--tv-in-bwd 0" "--tv-in-bwd 1 --parts 1"; do
timeout 600 python bench.py --steps 100 --warmup 10 $extra --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
done
for extra in "--tv-in-bwd 1" "--tv-in-bwd 0"; do
timeout 600 python bench.py --workload garden_stage0 --steps 40 --warmup 10 --skip-cpu --skip-reference $extra > gpurun_out/bench_garden.json 2> gpurun_out/bench_garden.err; python -c "
import json
for l in open('gpurun_out/bench_garden.json'):
    if l.startswith('{'):
        d=json.loads(l); print('GARDEN $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_garden.err
done
