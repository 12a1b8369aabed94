#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_beta.py tests/test_gpu_bf16_parity.py tests/test_gpu_mlp.py tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "beta or psnr_drift or viewdir or backward" > gpurun_out/r02k_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02k_tests.log; tail -8 gpurun_out/r02k_tests.log
