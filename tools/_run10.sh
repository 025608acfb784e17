set -x
mkdir -p gpurun_out
timeout 400 python tools/sweep_decode.py 64 753 "" > gpurun_out/r10_sweep.log 2>&1; tail -2 gpurun_out/r10_sweep.log | cut -c1-150
timeout 700 python bench.py > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r10_bench.err
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 64 430 gpurun_out/r10_trace_b64.json > gpurun_out/r10_trace_b64.log 2>&1; head -16 gpurun_out/r10_trace_b64.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r10_pytest.log; grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/r10_pytest.log | head
