#!/bin/bash
# r02a: whole GPU suite (old + new tests), then the bench line with the new legs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.json
