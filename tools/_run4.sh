set -x
mkdir -p gpurun_out
for f in tests/test_backward_gpu.py tests/test_parity_bf16_gpu.py tests/test_parity_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q > gpurun_out/r4_$n.log 2>&1; echo "rc=$?" >> gpurun_out/r4_$n.log
  tail -6 gpurun_out/r4_$n.log
done
timeout 400 python tools/sweep_decode.py 64 753 "" "VB_ATTN_DECODE_MMA=0" > gpurun_out/r4_sweep.log 2>&1; tail -4 gpurun_out/r4_sweep.log
timeout 300 python tools/sweep_decode.py 1 753 "" > gpurun_out/r4_b1.log 2>&1; tail -2 gpurun_out/r4_b1.log
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 1 430 gpurun_out/r4_trace_b1.json > gpurun_out/r4_trace_b1.log 2>&1; head -12 gpurun_out/r4_trace_b1.log
VB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attn_decode -s 4800 -c 2 -o gpurun_out/r4_prof_attn_decode_mma python tools/profile_decode.py 64 420 bf16 > gpurun_out/r4_ncu_attn_mma.log 2>&1; tail -2 gpurun_out/r4_ncu_attn_mma.log
VB_ATTN_DECODE_MMA=0 VB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attn_decode -s 4800 -c 2 -o gpurun_out/r4_prof_attn_decode_2phase python tools/profile_decode.py 64 420 bf16 > gpurun_out/r4_ncu_attn_2p.log 2>&1; tail -2 gpurun_out/r4_ncu_attn_2p.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05 -s 6 -c 1 -o gpurun_out/r4_prof_attn_tcgen05_pp python tools/bench_attention.py 64 1025 > gpurun_out/r4_ncu_fa.log 2>&1; tail -2 gpurun_out/r4_ncu_fa.log
timeout 300 python tools/profile_codec.py 32 > gpurun_out/r4_codec_b32.log 2>&1; grep -E "==|total|lstm" gpurun_out/r4_codec_b32.log
