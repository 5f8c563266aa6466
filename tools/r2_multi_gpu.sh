#!/bin/bash
# round-2 two-GPU check: sharded scoring parity (NCCL) and the bench at N = 2
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r02_check_multi_gpu_n$N.json 2> gpurun_out/r02_check_multi_gpu_n$N.err
tail -n 3 gpurun_out/r02_check_multi_gpu_n$N.err; cut -c1-1500 gpurun_out/r02_check_multi_gpu_n$N.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-producer > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -n 3 gpurun_out/r02_bench_n$N.err; cut -c1-3000 gpurun_out/r02_bench_n$N.json
