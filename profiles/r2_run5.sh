# Round-2, GPU call 5: fused backward v2 (16 scatter warps), parity diagnostics, mark_untrained, new bench.py on three workloads.
set -x
timeout 300 python profiles/fusedprobe.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_stage0.py -q -k "fused_backward" > gpurun_out/t_fused.log 2>&1; tail -3 gpurun_out/t_fused.log
timeout 1500 python -m pytest tests/test_gpu_reference_parity.py -q > gpurun_out/t_parity.log 2>&1; grep -E "^E  .*(Assertion|assert )|passed|failed" gpurun_out/t_parity.log | cut -c1-400 | head -30
timeout 600 python bench.py --steps 60 --warmup 10 --fused-bwd 1 --skip-cpu --skip-reference --psnr-iters 0 > gpurun_out/bench_fused1.json 2> gpurun_out/bench_fused1.err; tail -c 1500 gpurun_out/bench_fused1.json; tail -3 gpurun_out/bench_fused1.err
timeout 900 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --workload garden_stage0 --steps 40 --warmup 10 --skip-cpu > gpurun_out/bench_garden.json 2> gpurun_out/bench_garden.err; tail -c 2500 gpurun_out/bench_garden.json; tail -3 gpurun_out/bench_garden.err
timeout 600 python bench.py --workload lego_stage1 --steps 20 --warmup 5 > gpurun_out/bench_stage1.json 2> gpurun_out/bench_stage1.err; tail -c 1500 gpurun_out/bench_stage1.json; tail -3 gpurun_out/bench_stage1.err
