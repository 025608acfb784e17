set -x
mkdir -p gpurun_out
for f in tests/test_backward_gpu.py tests/test_parity_bf16_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q > gpurun_out/r5_$n.log 2>&1; echo "rc=$?" >> gpurun_out/r5_$n.log
  grep -E "^E  |^FAILED|passed|failed" gpurun_out/r5_$n.log | head -20
done
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 1 430 gpurun_out/r5_trace_b1.json > gpurun_out/r5_trace_b1.log 2>&1; head -16 gpurun_out/r5_trace_b1.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r5_trace_b1.json'))
print([(s['kernel'], round(s['us'],2)) for s in j['stage_us'][:12]])
PY
timeout 300 python tools/sweep_decode.py 1 753 "" > gpurun_out/r5_b1.log 2>&1; tail -2 gpurun_out/r5_b1.log | cut -c1-150
timeout 600 python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r5_bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r5_bench.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','frac','ms','fp32_exact','bf16_match_rate','first_divergence','encode_ms_per_utt','decode_ms_per_utt')}) for k,v in j.items() if k in ('value','ms_per_step','roofline','roofline_nar','roofline_b1','p50_utt_latency_ms','parity','config2','config3','config4','parity_mode','phase_ms_per_step')})
PY
