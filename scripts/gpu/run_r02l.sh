#!/bin/bash
# r02l: full bench line on one GPU (after the bench reorder / constants fix) + beta tests
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02l_bench.err
timeout 600 python -m pytest tests/test_gpu_beta.py -m gpu -q -p no:cacheprovider > gpurun_out/r02l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_tests.log; tail -4 gpurun_out/r02l_tests.log
