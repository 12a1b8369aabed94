#!/bin/bash
# r02g (2 GPUs): peer-memory gradient exchange: correctness + timing; bench line at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_check_ddp.py > gpurun_out/r02g_ddp_check.json 2> gpurun_out/r02g_ddp_check.err
echo "ddp check rc=$?"; tail -c 1500 gpurun_out/r02g_ddp_check.json; tail -c 1500 gpurun_out/r02g_ddp_check.err
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02g_bench_n2.json
