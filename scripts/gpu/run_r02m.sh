#!/bin/bash
# r02m: density out of the fused trunk kernel + L2 prefetch in the weight-gradient GEMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02m_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02m_tests.log
timeout 300 python scripts/ab_wgrad.py 1024 > gpurun_out/r02m_ab_wgrad.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
BN_NT_PREFETCH=0 timeout 600 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02m_bench_nopf.json 2> gpurun_out/r02m_bench_nopf.err
tail -3 gpurun_out/r02m_tests.log; cat gpurun_out/r02m_ab_wgrad.txt | tail -16; head -c 600 gpurun_out/r02m_bench.json; echo; head -c 300 gpurun_out/r02m_bench_nopf.json
