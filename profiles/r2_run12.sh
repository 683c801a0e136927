# Round-2, GPU call 12 (8 GPUs, charged 8x: kept short): the default data-parallel bench line as the driver launches it, then the hybrid optimizer.
set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
N=${N:-8}
for dp in ${MODES:-auto nvls hybrid}; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 10 --dp $dp --skip-reference --skip-cpu --psnr-iters 0 > gpurun_out/bench_n${N}_$dp.json 2> gpurun_out/bench_n${N}_$dp.err
  python -c "
import json
for l in open('gpurun_out/bench_n${N}_$dp.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N$N $dp |', d['config']['parallelism'], round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d.get('dp_check'), d.get('dp_stage_ms'))"
  grep -E "unavailable|Error|error" gpurun_out/bench_n${N}_$dp.err | head -5
done
