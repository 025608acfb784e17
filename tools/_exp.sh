set +e
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -1 gpurun_out/bench_r1_final.json | cut -c1-400
VB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_ar_b64_eager_v2.csv python tools/profile_decode.py 64 30 bf16 ar_only > gpurun_out/l1.log 2>&1; tail -1 gpurun_out/l1.log | cut -c1-200
VB_NO_GRAPH=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_2phase_pf -s 100 -c 1 -o gpurun_out/attn_decode_pf python tools/profile_decode.py 64 12 bf16 ar_only > gpurun_out/l2.log 2>&1; tail -1 gpurun_out/l2.log | cut -c1-200
