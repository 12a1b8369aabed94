#!/bin/bash
# r02e: dgrad chain correctness + A/B timings in one process + traces with wait accounting
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_train.py tests/test_gpu_mlp.py tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider -k "first_order or trainer_step or graph_step or psnr_drift or backward or frozen or train_loop" > gpurun_out/r02e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_tests.log
tail -6 gpurun_out/r02e_tests.log
timeout 300 python scripts/ab_chain.py > gpurun_out/r02e_ab.txt 2>&1; tail -12 gpurun_out/r02e_ab.txt
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02e_trace_train.txt 2>&1
BN_CHAIN_ONEPASS=1 timeout 120 python scripts/trace_chain.py train > gpurun_out/r02e_trace_train_onepass.txt 2>&1
timeout 120 python scripts/trace_chain.py > gpurun_out/r02e_trace_sigma.txt 2>&1
