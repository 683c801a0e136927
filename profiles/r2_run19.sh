# Round-2, GPU call 19 (1 GPU): evaluation renderer with device-side alive-ray rounds (tests + timing), per-launch timing of the stage-1 step.
set -x
timeout 900 python -m pytest tests/test_gpu_stage0.py tests/test_gpu_reference_parity.py -m gpu -q -x -k "render" > gpurun_out/t_render.log 2>&1; tail -8 gpurun_out/t_render.log; grep -E "^E  .*(Assertion|assert |Error)|^FAILED" gpurun_out/t_render.log | cut -c1-400 | head
timeout 600 python profiles/render_probe.py 1500 > gpurun_out/render_probe.json 2> gpurun_out/render_probe.err; cat gpurun_out/render_probe.json; tail -3 gpurun_out/render_probe.err
timeout 600 python profiles/s1_stage_probe.py > gpurun_out/s1_stage_probe.json 2> gpurun_out/s1_stage_probe.err; cat gpurun_out/s1_stage_probe.json; tail -3 gpurun_out/s1_stage_probe.err
