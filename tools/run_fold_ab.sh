#!/bin/bash
mkdir -p gpurun_out
(SWEEP_REPS=4 timeout 700 python tools/sweep_decode.py 64 753 "VB_L2_HINTS=1" "VB_L2_HINTS=0" "VB_L2_HINTS=3" "VB_L2_HINTS=1,VB_KV_PREFETCH_PCT=30" "VB_L2_HINTS=1,VB_KV_PREFETCH_PCT=50" "VB_DECODE_FOLD=0,VB_L2_HINTS=0" "VB_DECODE_FOLD=0,VB_L2_HINTS=1" 2>&1 | tail -28) > gpurun_out/fold_sweep15_b64.log 2>&1
(SWEEP_REPS=3 timeout 300 python tools/sweep_decode.py 1 400 "VB_L2_HINTS=1" "VB_L2_HINTS=0" 2>&1 | tail -6) > gpurun_out/fold_sweep15_b1.log 2>&1
for f in fold_sweep15_b64 fold_sweep15_b1; do echo "== $f"; cat gpurun_out/$f.log; done
