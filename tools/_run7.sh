set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r7_bench_2gpu.json 2> gpurun_out/r7_bench_2gpu.err; echo "rc=$?"; tail -c 600 gpurun_out/r7_bench_2gpu.err; cat gpurun_out/r7_bench_2gpu.json | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r7_ref_2gpu.json 2> gpurun_out/r7_ref_2gpu.err; echo "rc=$?"; cat gpurun_out/r7_ref_2gpu.json | cut -c1-600
