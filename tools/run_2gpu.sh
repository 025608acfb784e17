#!/bin/bash
# two-GPU sanity of the final tree: the second-device test and the weak-scaling bench line at N=2
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_parity_bf16_gpu.py tests/test_parity_gpu.py -x -q -k "second_gpu or tensor_core_path_matches_gemv" 2>&1 | tail -4) > gpurun_out/final_2gpu_tests.log 2>&1
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final_bench_2gpu.json 2> gpurun_out/final_bench_2gpu.err; tail -3 gpurun_out/final_bench_2gpu.err) > gpurun_out/final_bench_2gpu.log 2>&1
cat gpurun_out/final_2gpu_tests.log gpurun_out/final_bench_2gpu.log; head -c 600 gpurun_out/final_bench_2gpu.json
