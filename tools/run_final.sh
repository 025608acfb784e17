#!/bin/bash
# Final validation of the tree on one box: GPU test-suite, smoke, bench line, device timelines, ncu captures of the
# LayerNorm-folded decode step (DRAM traffic / launch list of one step, --set full of the two projection kernels).
mkdir -p gpurun_out
(timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/final_tests.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5) > gpurun_out/final_smoke.log 2>&1
(timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -3 gpurun_out/final_bench.err) > gpurun_out/final_bench.log 2>&1
T=valle_b200/lib/libvalle_b200_trace.so
(VB_LIB_PATH=$T timeout 200 python tools/trace_ar_step.py 64 430 gpurun_out/final_trace_b64.json 2>&1 | tail -11) > gpurun_out/final_trace_b64.log 2>&1
(VB_LIB_PATH=$T VB_DECODE_FOLD=0 timeout 200 python tools/trace_ar_step.py 64 430 gpurun_out/final_trace_b64_unfolded.json 2>&1 | tail -13) > gpurun_out/final_trace_b64_unfolded.log 2>&1
(VB_LIB_PATH=$T timeout 200 python tools/trace_ar_step.py 1 300 gpurun_out/final_trace_b1.json 2>&1 | tail -12) > gpurun_out/final_trace_b1.log 2>&1
(VB_NO_GRAPH=1 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 27700 -c 240 --csv --log-file gpurun_out/final_ar_step_dram.csv python tools/profile_decode.py 64 380 bf16 ar_only 2>&1 | tail -3) > gpurun_out/final_ncu_dram.log 2>&1
(VB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:gemm_decode -s 300 -c 6 -o gpurun_out/final_gemm_decode -f python tools/profile_decode.py 64 40 bf16 ar_only 2>&1 | tail -3) > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out/final_* | head -30
for f in final_tests final_smoke final_bench final_trace_b64 final_trace_b64_unfolded final_trace_b1 final_ncu_dram final_ncu_full; do echo "== $f"; cat gpurun_out/$f.log | cut -c1-300; done
