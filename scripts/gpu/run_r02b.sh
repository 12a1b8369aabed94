#!/bin/bash
# r02b: chain kernels with the bias on the tensor core: correctness, clock64 trace, timing; bf16 conditioning probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_gemm.py tests/test_gpu_bf16_parity.py -m gpu -q -x -k "chain or gemm or forward or backward or stored or first_order" -p no:cacheprovider > gpurun_out/r02b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_tests.log
tail -4 gpurun_out/r02b_tests.log
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02b_trace_train.txt 2>&1
timeout 120 python scripts/trace_chain.py > gpurun_out/r02b_trace_sigma.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-tile-products --no-composite > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02b_bench.err
timeout 900 python scripts/r02_probe_bf16.py > gpurun_out/r02b_probe.txt 2>&1
echo "probe rc=$?"; tail -5 gpurun_out/r02b_probe.txt
