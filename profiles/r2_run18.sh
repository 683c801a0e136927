# Round-2, GPU call 18 (1 GPU): full gpu suite after the tv_in_bwd removal + antialias tests with their final thresholds; stage-1 bench with and without antialias.
set -x
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/t_all.log 2>&1; tail -22 gpurun_out/t_all.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_all.log | cut -c1-400 | head -20
for aa in 1 0; do
timeout 600 python bench.py --workload lego_stage1 --steps 40 --warmup 5 --antialias $aa > gpurun_out/bench_s1_aa$aa.json 2> gpurun_out/bench_s1_aa$aa.err; tail -c 1500 gpurun_out/bench_s1_aa$aa.json; tail -2 gpurun_out/bench_s1_aa$aa.err
done
