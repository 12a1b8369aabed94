#!/bin/bash
# r03a: weight-gradient GEMM with the B tile multicast inside 4-CTA clusters (BN_NT_MC=1)
mkdir -p gpurun_out
BN_NT_MC=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider -k "nt_" > gpurun_out/r03a_tests_gemm.log 2>&1
echo "gemm tests exit $?" >> gpurun_out/r03a_tests_gemm.log; tail -4 gpurun_out/r03a_tests_gemm.log
BN_NT_MC=1 timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "mlp or train or bf16" > gpurun_out/r03a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r03a_tests.log; tail -4 gpurun_out/r03a_tests.log
BN_NT_MC=1 timeout 200 python scripts/trace_wgrad.py 1024 > gpurun_out/r03a_trace_wgrad.txt 2>&1; tail -4 gpurun_out/r03a_trace_wgrad.txt
timeout 300 python scripts/ab_wgrad.py 1024 > gpurun_out/r03a_ab.txt 2>&1; grep "base\|multicast" gpurun_out/r03a_ab.txt
