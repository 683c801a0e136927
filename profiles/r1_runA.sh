set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 2500 gpurun_out/bench_r1f.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1f.json 2>&1; tail -c 800 gpurun_out/bench_ref_r1f.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 1 --skip-cpu --no-graph > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_s0_encode_bwd|k_adam_tables|k_s0_encode_fwd|k_mlp_bwd|k_mlp_fwd|k_s0_count_warp|k_s0_composite" -s 14 -c 9 -o gpurun_out/prof_r1f python bench.py --steps 1 --warmup 1 --skip-cpu --no-graph --parts 1 > /dev/null 2>&1
ls -la gpurun_out/
