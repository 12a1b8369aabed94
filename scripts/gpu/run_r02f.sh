#!/bin/bash
# r02f: L2 eviction hints A/B (forward chain, dgrad chain, per-layer GEMMs), traces, PSNR tests
mkdir -p gpurun_out
timeout 400 python scripts/ab_chain.py > gpurun_out/r02f_ab.txt 2>&1; tail -22 gpurun_out/r02f_ab.txt
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02f_trace_train.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_train.py tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "psnr_drift or shared_trunk or golden or training_gradients" > gpurun_out/r02f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_tests.log
tail -8 gpurun_out/r02f_tests.log
