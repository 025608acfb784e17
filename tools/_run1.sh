set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke1.log
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 64 430 gpurun_out/r2_trace_b64.json > gpurun_out/r2_trace_b64.log 2>&1
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so VB_KV_PREFETCH_PCT=0 timeout 300 python tools/trace_ar_step.py 64 430 gpurun_out/r2_trace_b64_pf0.json > gpurun_out/r2_trace_b64_pf0.log 2>&1
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 1 430 gpurun_out/r2_trace_b1.json > gpurun_out/r2_trace_b1.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_bench1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 12 -c 2 -o gpurun_out/r2_prof_gemm_tcgen05 python tools/bench_gemm.py 48000 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -3 gpurun_out/r2_pytest1.log; tail -2 gpurun_out/r2_smoke1.log; cat gpurun_out/r2_trace_b64.log | head -30; tail -c 600 gpurun_out/r2_bench1.err
timeout 600 python tools/sweep_decode.py 64 753 "" "VB_KV_PF_BULK=1" "VB_ATTN_PF_K_FROM=40,VB_ATTN_PF_V_FROM=40" "VB_KV_PF_BULK=1,VB_ATTN_PF_K_FROM=40,VB_ATTN_PF_V_FROM=40" "VB_KV_PREFETCH_PCT=0,VB_ATTN_PF_K_FROM=0,VB_ATTN_PF_V_FROM=0" "VB_ATTN_PF_K_FROM=40,VB_ATTN_PF_V_FROM=0" "VB_KV_PREFETCH_PCT=60,VB_KV_PF_BULK=1,VB_ATTN_PF_K_FROM=60,VB_ATTN_PF_V_FROM=60" "VB_KV_PREFETCH_PCT=0" > gpurun_out/r2_sweep1.log 2>&1
cat gpurun_out/r2_sweep1.log | tail -20
