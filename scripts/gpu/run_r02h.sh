#!/bin/bash
# r02h: single-fence epilogue; cosine store variants; dgrad chain variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_mlp.py -m gpu -q -x -p no:cacheprovider -k "stored or first_order or chain" > gpurun_out/r02h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_tests.log; tail -3 gpurun_out/r02h_tests.log
BN_CHAIN_CSTG=1 timeout 300 python -m pytest tests/test_gpu_bf16_parity.py -m gpu -q -x -p no:cacheprovider -k "stored" > gpurun_out/r02h_tests_cstg.log 2>&1
echo "pytest cstg rc=$?" >> gpurun_out/r02h_tests_cstg.log; tail -3 gpurun_out/r02h_tests_cstg.log
timeout 400 python scripts/ab_chain.py > gpurun_out/r02h_ab.txt 2>&1; tail -32 gpurun_out/r02h_ab.txt
BN_CHAIN_CBOX2=1 timeout 400 python scripts/ab_chain.py > gpurun_out/r02h_ab_cbox2.txt 2>&1; grep "default" gpurun_out/r02h_ab_cbox2.txt
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02h_trace_train.txt 2>&1
