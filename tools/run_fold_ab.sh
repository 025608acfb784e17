#!/bin/bash
mkdir -p gpurun_out
T=valle_b200/lib/libvalle_b200_trace.so
tr() {  # name, B, frames, env...
  local name=$1 B=$2 F=$3; shift 3
  (env VB_LIB_PATH=$T "$@" timeout 200 python tools/trace_ar_step.py $B $F gpurun_out/trace_$name.json 2>&1 | tail -11) > gpurun_out/trace_$name.log 2>&1
}
(timeout 300 python -m pytest tests/test_parity_bf16_gpu.py -x -q -k "fold or big_short or config1" 2>&1 | tail -4) > gpurun_out/fold_t8.log 2>&1
(VB_RED_WAIT_FULL=0 timeout 300 python -m pytest tests/test_parity_bf16_gpu.py -x -q -k "fold or big_short" 2>&1 | tail -4) > gpurun_out/fold_t8_nowait.log 2>&1
tr fold_v8 64 430 A=1
tr fold_nowait_v8 64 430 VB_RED_WAIT_FULL=0
(SWEEP_REPS=4 timeout 500 python tools/sweep_decode.py 64 753 "" "VB_DECODE_FOLD=0" "VB_RED_WAIT_FULL=0" "VB_SPLITS_FFN1=2" "VB_SPLITS_OUT=4" 2>&1 | tail -20) > gpurun_out/fold_sweep11_b64.log 2>&1
(SWEEP_REPS=3 timeout 300 python tools/sweep_decode.py 1 400 "" "VB_DECODE_FOLD=0" "VB_RED_WAIT_FULL=0" 2>&1 | tail -9) > gpurun_out/fold_sweep11_b1.log 2>&1
for f in fold_t8 fold_t8_nowait trace_fold_v8 trace_fold_nowait_v8 fold_sweep11_b64 fold_sweep11_b1; do echo "== $f"; cat gpurun_out/$f.log; done
