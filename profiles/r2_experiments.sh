# Round-2 first GPU call: validate and measure the variants that were compiled but not run in round 1.
#   gpurun --timeout 1200 -- 'bash profiles/r2_experiments.sh'
set -x
N2M_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_stage0.py -q -x -k "scatter_level or compact_mlp or level_pipelined or two_issuer" 2>&1 | tail -5
timeout 120 python profiles/redbench.py 2>&1 | tail -26
for extra in "" "--mlp-bwd two-tile" "--mlp-bwd two-tile-2issuers" "--scatter-cuts 10 --level-pipe" "--scatter-cuts 10 --level-pipe --l2-persist-mb 64" "--scatter-cuts 10 --level-pipe --mlp-fwd-compact" "--scatter-cuts 10" "--scatter-cuts 8,12" "--mlp-fwd-compact" "--scatter-cuts 10 --mlp-fwd-compact" "--parts 1" "--parts 1 --scatter-cuts 10 --mlp-fwd-compact"; do
  timeout 200 python bench.py --steps 100 --warmup 10 --skip-cpu $extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); c=d['config']; print('$extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"
done
