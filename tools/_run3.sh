set -x
mkdir -p gpurun_out
for f in tests/test_backward_gpu.py tests/test_codec.py tests/test_parity_bf16_gpu.py tests/test_modules_gpu.py tests/test_parity_gpu.py tests/test_gemm_gpu.py tests/test_boundary.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x > gpurun_out/r3_$n.log 2>&1; echo "rc=$?" >> gpurun_out/r3_$n.log
  tail -4 gpurun_out/r3_$n.log
done
timeout 600 python tools/sweep_decode.py 64 753 "" "VB_ATTN_DECODE_MMA=0" "VB_KV_PF_BY_ROW=1" "VB_KV_PF_BY_ROW=1,VB_KV_PREFETCH_PCT=50" "VB_KV_PREFETCH_PCT=50" "VB_KV_PREFETCH_PCT=60" "VB_KV_PREFETCH_PCT=0" > gpurun_out/r3_sweep.log 2>&1
tail -16 gpurun_out/r3_sweep.log
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 64 430 gpurun_out/r3_trace_b64.json > gpurun_out/r3_trace_b64.log 2>&1; head -20 gpurun_out/r3_trace_b64.log
for bm in 2 1 0; do VB_GRID_BARRIER=$bm VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 1 430 gpurun_out/r3_trace_b1_bar$bm.json > gpurun_out/r3_trace_b1_bar$bm.log 2>&1; head -22 gpurun_out/r3_trace_b1_bar$bm.log; done
for bm in 2 0; do VB_GRID_BARRIER=$bm timeout 300 python tools/profile_codec.py 16 > gpurun_out/r3_codec_bar$bm.log 2>&1; grep -E "lstm|==|total" gpurun_out/r3_codec_bar$bm.log; done
