#!/bin/bash
mkdir -p gpurun_out
(SWEEP_REPS=4 timeout 700 python tools/sweep_decode.py 64 753 "" "VB_SPLITS_QKV=6" "VB_SPLITS_QKV=8" "VB_SPLITS_OUT=16" "VB_SPLITS_FFN1=6" "VB_KV_PREFETCH_PCT=30" "VB_KV_PREFETCH_PCT=35" "VB_SPLITS_FFN2=13" 2>&1 | tail -32) > gpurun_out/fold_sweep16_b64.log 2>&1
cat gpurun_out/fold_sweep16_b64.log
