# Round-2, GPU call 7 (2 GPUs): NVLS / peer / NCCL data-parallel optimizers -- correctness (tests/test_gpu_dp.py) and step time.
set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_dp.py -q -s > gpurun_out/t_dp.log 2>&1; tail -5 gpurun_out/t_dp.log; grep -E "^E  |NVLS" gpurun_out/t_dp.log | cut -c1-300 | head
for dp in hybrid peer nvls nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 10 --dp $dp > gpurun_out/bench_n2_$dp.json 2> gpurun_out/bench_n2_$dp.err
  python -c "
import json
for l in open('gpurun_out/bench_n2_$dp.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N2 $dp |', d['config']['parallelism'], round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d.get('dp_check'))"
  grep -E "unavailable|Error|error" gpurun_out/bench_n2_$dp.err | head -5
done
timeout 900 python -m pytest tests/test_gpu_reference_parity.py -q -k "fused_step and garden" > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-300 | head -10
timeout 600 python bench.py --steps 60 --warmup 10 --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json
for l in open('gpurun_out/bench_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N1 default |', round(d['ms_per_step'],4), f\"{d['value']:.3e}\", 'e2e', round(d['e2e']['ms_per_step'],4), d['roofline']['stage_ms_cold_l2'])"; tail -2 gpurun_out/bench_x.err
timeout 600 python -m pytest tests/test_gpu_stage0.py -q -x > gpurun_out/t_stage0.log 2>&1; tail -3 gpurun_out/t_stage0.log
