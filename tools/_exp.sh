set +e
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r1_v5.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v5.json').read())
print(d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_nar']['frac'], d['p50_utt_latency_ms'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
