#!/bin/bash
# Multi-GPU visit: DDP parity check + bench under torchrun at N GPUs.   gpurun --gpus 2 -- 'bash tools/gpu_check2.sh <tag> 2'
TAG=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py > gpurun_out/${TAG}_ddp_check.log 2>&1
echo "ddp_check exit $?"; tail -4 gpurun_out/${TAG}_ddp_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
