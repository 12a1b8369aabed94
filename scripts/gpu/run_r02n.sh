#!/bin/bash
# r02n: same-box A/B of the density-from-the-chain change and of what bounds the weight-gradient GEMM
mkdir -p gpurun_out
timeout 300 python scripts/ab_wgrad.py 1024 > gpurun_out/r02n_ab.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "bf16 or parity or golden or mlp" > gpurun_out/r02n_tests.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
BN_CHAIN_NO_SIG=1 timeout 300 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02n_bench_nosig.json 2> gpurun_out/r02n_bench_nosig.err
timeout 300 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02n_bench2.json 2> gpurun_out/r02n_bench2.err
BN_CHAIN_NO_SIG=1 timeout 300 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02n_bench_nosig2.json 2> gpurun_out/r02n_bench_nosig2.err
cat gpurun_out/r02n_ab.txt | tail -16; tail -3 gpurun_out/r02n_tests.log
for f in r02n_bench r02n_bench_nosig r02n_bench2 r02n_bench_nosig2; do head -c 330 gpurun_out/$f.json | tail -c 130; echo; done
