#!/bin/bash
# quick developer loop (run under gpurun): SECTIONS="gemm attn fwd perf prof" R=tag bash tests/run_quick.sh
mkdir -p gpurun_out
R=${R:-quick}
SECTIONS=${SECTIONS:-"gemm attn fwd perf prof"}
for s in $SECTIONS; do
  timeout 300 python tests/gpu_probe.py $s fp16 2>&1 | grep -v "^$" | tail -22
done > gpurun_out/${R}_probe.log 2>&1
cat gpurun_out/${R}_probe.log
