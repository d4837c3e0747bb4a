mkdir -p gpurun_out
for s in "ln fp16" "gemm fp16" "gemm bf16" "attn fp16" "attn bf16" "fwd fp16" "fwd bf16" "sample fp16" "perf fp16"; do
  timeout 300 python tests/gpu_probe.py $s 2>&1 | grep -v "^$" | tail -25
done > gpurun_out/probe1.log 2>&1
tail -100 gpurun_out/probe1.log
