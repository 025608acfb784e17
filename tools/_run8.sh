set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r8_pytest.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/r8_pytest.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/r8_smoke.log 2>&1; tail -3 gpurun_out/r8_smoke.log
timeout 400 python tools/sweep_decode.py 64 753 "" > gpurun_out/r8_sweep.log 2>&1; tail -2 gpurun_out/r8_sweep.log | cut -c1-150
