#!/bin/bash
# Final validation of the tree on one box: GPU test-suite, smoke, (bench line, device timeline when time allows).
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/final_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5) > gpurun_out/final_smoke.log 2>&1
(SWEEP_REPS=2 timeout 120 python tools/sweep_decode.py 64 753 "" 2>&1 | tail -2) > gpurun_out/final_sweep.log 2>&1
for f in final_tests final_smoke final_sweep; do echo "== $f"; cat gpurun_out/$f.log | cut -c1-300; done
