# Round-2, GPU call 25 (2 GPUs): the driver's N = 2 command on the final code (data-parallel step + dp_check against the NCCL path).
set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/bench_n2.err
python -c "
import json
for l in open('gpurun_out/r2_bench_n2_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N2 |', d['config']['parallelism'], round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d.get('dp_check'), d.get('dp_stage_ms'))"
grep -E "unavailable|Error|error|Traceback" gpurun_out/bench_n2.err | head -5
