#!/bin/bash
# r02j: beta channel tests, determinism test, goldens, then the whole suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_beta.py tests/test_gpu_bf16_parity.py tests/test_gpu_render.py -m gpu -q -p no:cacheprovider -k "beta or deterministic or golden" > gpurun_out/r02j_tests_new.log 2>&1
echo "pytest new rc=$?" >> gpurun_out/r02j_tests_new.log; tail -15 gpurun_out/r02j_tests_new.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_tests.log; tail -8 gpurun_out/r02j_tests.log
