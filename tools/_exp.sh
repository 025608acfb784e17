set +e
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r1_v3.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v3.json').read())
print(d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline_nar']['frac'], d['p50_utt_latency_ms'])
PY
