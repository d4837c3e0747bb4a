#!/bin/bash
# Developer GPU session (run under gpurun): tests, smoke, bench, ncu launch list + full captures.
mkdir -p gpurun_out
R=${R:-r02}
(timeout 1800 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -15) > gpurun_out/${R}_pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/${R}_smoke.log
(timeout 600 python bench.py 2>&1 | tail -1) > gpurun_out/${R}_bench.json
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 118 -c 118 --csv \
    --log-file gpurun_out/${R}_launches.csv python tests/gpu_probe.py one > gpurun_out/${R}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 94 -c 5 \
    -o gpurun_out/${R}_prof_gemm -f python tests/gpu_probe.py one > gpurun_out/${R}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention3_kernel -s 21 -c 1 \
    -o gpurun_out/${R}_prof_attn -f python tests/gpu_probe.py one > gpurun_out/${R}_ncu_attn.log 2>&1
if [ "${VAE:-0}" = "1" ]; then
# the decoder's implicit-GEMM convolutions: second decode, three GEMMs from the 128^2 x 256-channel level
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 72 -c 3 \
    -o gpurun_out/${R}_prof_vae -f python tests/gpu_probe.py vae > gpurun_out/${R}_ncu_vae.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
    --log-file gpurun_out/${R}_vae_launches.csv python tests/gpu_probe.py vae > gpurun_out/${R}_ncu_vae_launch.log 2>&1
fi
fi
cat gpurun_out/${R}_pytest_gpu.log gpurun_out/${R}_smoke.log gpurun_out/${R}_bench.json
