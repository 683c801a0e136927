# Round-2, GPU call 24 (1 GPU): FINAL -- full gpu suite, smoke, default bench lines of the three workloads (saved under profiles/).
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_all.log | cut -c1-500 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --workload lego_stage1 > gpurun_out/r2_bench_stage1_final.json 2> gpurun_out/bench_s1.err; tail -c 1200 gpurun_out/r2_bench_stage1_final.json; tail -2 gpurun_out/bench_s1.err
timeout 600 python bench.py --workload garden_stage0 --skip-cpu > gpurun_out/r2_bench_garden_final.json 2> gpurun_out/bench_g.err; tail -c 900 gpurun_out/r2_bench_garden_final.json | head -c 600; tail -2 gpurun_out/bench_g.err
timeout 900 python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/bench_d.err; tail -c 2500 gpurun_out/r2_bench_default_final.json; tail -2 gpurun_out/bench_d.err
