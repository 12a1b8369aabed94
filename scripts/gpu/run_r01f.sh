#!/bin/bash
# Round-1 (session f) GPU pass: every GPU test file in its own process, the e2e feed experiment, the DSM micro-benchmark,
# smoke(), the default bench line and the reference arm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu/run_gpu_tests.sh tests/test_gpu_dsm.py tests/test_gpu_train.py tests/test_gpu_sampler.py tests/test_gpu_composite.py \
     tests/test_gpu_gemm.py tests/test_gpu_mlp.py tests/test_gpu_render.py
echo "tests rc=$?"
timeout 300 python scripts/exp_e2e.py 60 > gpurun_out/exp_e2e.log 2>&1; echo "exp_e2e rc=$?"; tail -3 gpurun_out/exp_e2e.log
timeout 300 python scripts/bench_dsm.py > gpurun_out/bench_dsm.json 2> gpurun_out/bench_dsm.err; echo "bench_dsm rc=$?"; tail -c 1500 gpurun_out/bench_dsm.json
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench.json
timeout 300 python bench.py --feed inline --no-cpu-baseline --no-composite > gpurun_out/bench_inline.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_inline.json
