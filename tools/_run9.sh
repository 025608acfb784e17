set -x
mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r9_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r9_bench_ref.json 2> gpurun_out/r9_bench_ref.err; echo "ref rc=$?"
# eager launch list of a few decode steps around context 640 (B=64): 100 kernels per step
VB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 37000 -c 700 --csv --log-file gpurun_out/r9_launches_ar_b64.csv python tools/profile_decode.py 64 380 bf16 > gpurun_out/r9_ncu_launches.log 2>&1; python tools/summarize_launches.py gpurun_out/r9_launches_ar_b64.csv | head -14
# DRAM traffic of one step
VB_NO_GRAPH=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 37000 -c 400 --csv --log-file gpurun_out/r9_ar_step_dram.csv python tools/profile_decode.py 64 380 bf16 ar_only > gpurun_out/r9_ncu_dram.log 2>&1; python tools/ar_step_traffic.py gpurun_out/r9_ar_step_dram.csv gpurun_out/r9_ar_step_traffic.json 64 47 225 372
# NAR launch list (one continual call at the bench shape)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r9_launches_nar_b64.csv python tools/profile_decode.py 64 753 bf16 nar > gpurun_out/r9_ncu_nar.log 2>&1; python tools/summarize_launches.py gpurun_out/r9_launches_nar_b64.csv | head -14
VB_LIB_PATH=valle_b200/lib/libvalle_b200_trace.so timeout 300 python tools/trace_ar_step.py 64 430 gpurun_out/r9_trace_b64.json > gpurun_out/r9_trace_b64.log 2>&1; head -16 gpurun_out/r9_trace_b64.log
