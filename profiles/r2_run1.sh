# Round-2, GPU call 1: validate + measure the variants compiled (never run) in round 1; decides what is kept.
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
N2M_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_stage0.py -q -x -k "scatter_level or compact_mlp or level_pipelined or two_issuer" 2>&1 | tail -8
for extra in "" "--mlp-bwd two-tile-2issuers" "--scatter-cuts 10" "--scatter-cuts 10 --level-pipe" "--scatter-cuts 10 --level-pipe --l2-persist-mb 64" "--scatter-cuts 8,12" "--mlp-fwd-compact" "--parts 1" "--parts 1 --scatter-cuts 10 --level-pipe --l2-persist-mb 64" "--parts 1 --scatter-cuts 6,10,13"; do
  timeout 200 python bench.py --steps 60 --warmup 10 --skip-cpu $extra 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); c=d['config']; print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"
  tail -2 gpurun_out/err.txt
done
