set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -k "optimizer or forward_backward" > gpurun_out/r6_bwd.log 2>&1; grep -E "^E  |^FAILED|passed|failed" gpurun_out/r6_bwd.log | head
timeout 600 python -m pytest tests/test_parity_bf16_gpu.py -m gpu -q -k "small_batch or finished or free_running" > gpurun_out/r6_bf16.log 2>&1; grep -E "^E  |^FAILED|passed|failed" gpurun_out/r6_bf16.log | head
timeout 300 python tools/sweep_decode.py 1 753 "" > gpurun_out/r6_b1.log 2>&1; tail -2 gpurun_out/r6_b1.log | cut -c1-150
timeout 300 python tools/sweep_decode.py 4 753 "" > gpurun_out/r6_b4.log 2>&1; tail -2 gpurun_out/r6_b4.log | cut -c1-150
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 1 430 gpurun_out/r6_trace_b1.json > gpurun_out/r6_trace_b1.log 2>&1; head -16 gpurun_out/r6_trace_b1.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r6_trace_b1.json'))
print([(s['kernel'], round(s['us'],2)) for s in j['stage_us'][:12]])
PY
