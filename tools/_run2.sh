set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -40 gpurun_out/r2_pytest2.log
timeout 600 python tools/sweep_decode.py 64 753 "" "VB_KV_PF_BY_ROW=1" "VB_KV_PF_BY_ROW=1,VB_KV_PREFETCH_PCT=50" "VB_KV_PF_BY_ROW=1,VB_KV_PREFETCH_PCT=60" "VB_KV_PF_BY_ROW=1,VB_KV_PREFETCH_PCT=30" "VB_KV_PREFETCH_PCT=50" > gpurun_out/r2_sweep2.log 2>&1
cat gpurun_out/r2_sweep2.log | tail -14
timeout 300 python tools/profile_codec.py 16 > gpurun_out/r2_codec_prof.log 2>&1; tail -60 gpurun_out/r2_codec_prof.log
timeout 300 python tools/sweep_decode.py 1 753 "" > gpurun_out/r2_b1.log 2>&1; tail -3 gpurun_out/r2_b1.log
