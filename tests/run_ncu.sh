#!/bin/bash
# ncu captures only (run under gpurun): R=tag bash tests/run_ncu.sh
mkdir -p gpurun_out
R=${R:-ncu}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 118 -c 118 --csv \
    --log-file gpurun_out/${R}_launches.csv python tests/gpu_probe.py one > gpurun_out/${R}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 94 -c 5 \
    -o gpurun_out/${R}_prof_gemm -f python tests/gpu_probe.py one > gpurun_out/${R}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention3_kernel -s 21 -c 1 \
    -o gpurun_out/${R}_prof_attn -f python tests/gpu_probe.py one > gpurun_out/${R}_ncu_attn.log 2>&1
tail -3 gpurun_out/${R}_ncu_gemm.log gpurun_out/${R}_ncu_attn.log
