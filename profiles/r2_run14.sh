# Round-2, GPU call 14 (2 GPUs): data-parallel tests after the hybrid removal + sharded density-grid update; default 2-GPU bench line.
set -x
timeout 900 python -m pytest tests/test_gpu_dp.py -q -s > gpurun_out/t_dp.log 2>&1; tail -5 gpurun_out/t_dp.log; grep -E "^E  |NVLS" gpurun_out/t_dp.log | cut -c1-300 | head
for dp in auto; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --dp $dp > gpurun_out/bench_n2_$dp.json 2> gpurun_out/bench_n2_$dp.err
  python -c "
import json
for l in open('gpurun_out/bench_n2_$dp.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N2 $dp |', d['config']['parallelism'], round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d.get('dp_check'), d.get('dp_stage_ms'))"
  grep -E "unavailable|Error|error" gpurun_out/bench_n2_$dp.err | head -5
done
