# Round-2 final ncu evidence (1 GPU, final code): launch list of eager steps + full capture of the hot kernels, same command as the bench.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2f_launches_ncu.csv python bench.py --steps 4 --warmup 3 --skip-cpu --skip-reference --psnr-iters 0 --no-graph --no-prefetch > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_s0_bwd_fused|k_adam_tables|k_s0_encode_fwd|k_mlp_fwd|k_s0_composite|k_s0_count_warp|k_s0_encode_bwd" -s 24 -c 8 -o gpurun_out/r2f_prof python bench.py --steps 2 --warmup 3 --skip-cpu --skip-reference --psnr-iters 0 --no-graph --no-prefetch --parts 1 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | tail -5
