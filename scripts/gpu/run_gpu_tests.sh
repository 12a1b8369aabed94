#!/bin/bash
# Runs each GPU test file in its own process under a timeout (a hung kernel must not take the box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${@:-tests/test_gpu_sampler.py tests/test_gpu_composite.py tests/test_gpu_gemm.py tests/test_gpu_mlp.py tests/test_gpu_render.py}; do
  name=$(basename $f .py)
  echo "=== $f"
  timeout 600 python -m pytest $f -q -m gpu -x --tb=short -s > gpurun_out/$name.log 2>&1
  r=$?; [ $r -ne 0 ] && rc=$r
  tail -n 25 gpurun_out/$name.log
done
exit $rc
