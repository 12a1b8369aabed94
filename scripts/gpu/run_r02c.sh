#!/bin/bash
# r02c: chain kernels with per-K-block release + per-unit h stores: correctness, traces, timing; low-lr PSNR protocol probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_gemm.py tests/test_gpu_bf16_parity.py tests/test_gpu_render.py -m gpu -q -x -k "chain or gemm or forward or backward or stored or first_order or shared_trunk or full_size" -p no:cacheprovider > gpurun_out/r02c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_tests.log
tail -4 gpurun_out/r02c_tests.log
timeout 120 python scripts/trace_chain.py train > gpurun_out/r02c_trace_train.txt 2>&1
timeout 120 python scripts/trace_chain.py > gpurun_out/r02c_trace_sigma.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-tile-products --no-composite > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02c_bench.err
timeout 900 python scripts/r02_probe_bf16.py 3 > gpurun_out/r02c_probe3.txt 2>&1
echo "probe rc=$?"; tail -5 gpurun_out/r02c_probe3.txt
