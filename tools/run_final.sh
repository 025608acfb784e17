#!/bin/bash
# Final validation of the tree on one box: GPU test-suite, smoke, bench line (+ the reference arm), device timeline.
mkdir -p gpurun_out
(timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/final_tests.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5) > gpurun_out/final_smoke.log 2>&1
(timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -3 gpurun_out/final_bench.err) > gpurun_out/final_bench.log 2>&1
T=valle_b200/lib/libvalle_b200_trace.so
(VB_LIB_PATH=$T timeout 200 python tools/trace_ar_step.py 64 430 gpurun_out/final_trace_b64.json 2>&1 | tail -11) > gpurun_out/final_trace_b64.log 2>&1
for f in final_tests final_smoke final_bench final_trace_b64; do echo "== $f"; cat gpurun_out/$f.log | cut -c1-300; done
