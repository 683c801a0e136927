# multi-GPU scaling sweep (one box, 8 GPUs): PeerAdam at 8/4/2 ranks, NCCL all-reduce at 8, plus the 2-rank PeerAdam parity test
for cfg in "8 peer" "4 peer" "2 peer" "8 nccl"; do
  set -- $cfg
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $1 --steps 40 --warmup 10 --skip-cpu --dp $2 > gpurun_out/dp$1_$2.log 2>&1
  grep "^{" gpurun_out/dp$1_$2.log | tee -a gpurun_out/scale_r1f.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['parallelism'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
  grep -v "^{" gpurun_out/dp$1_$2.log | grep -vE "^\s*$|Warning|warn|OMP_NUM" | tail -6
done
timeout 300 python -m pytest tests/test_gpu_dp.py -q -x 2>&1 | tail -3
