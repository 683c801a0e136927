# Round-2, GPU call 3: fused backward kernel validation + parity tests (GradScaler back-off) + bench A/B.
set -x
timeout 600 python -m pytest tests/test_gpu_stage0.py -q -x -k "fused_backward" 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_gpu_reference_parity.py -q 2>&1 | grep -v Warning | grep -v "^  " | tail -40
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_reference_parity.py 2>&1 | tail -8
for extra in "" "--no-fused-bwd" "--parts 1" "--parts 1 --no-fused-bwd" "--parts 4"; do
  timeout 200 python bench.py --steps 60 --warmup 10 --skip-cpu $extra 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); c=d['config']; print('CFG $extra |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"
  tail -2 gpurun_out/err.txt
done
