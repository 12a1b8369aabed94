#!/bin/bash
# r02d: training chain with the two-pass second-half epilogue: whole GPU suite, traces, full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not psnr_drift_brdf" > gpurun_out/r02d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_tests.log
tail -15 gpurun_out/r02d_tests.log
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02d_trace_train.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02d_bench.err
