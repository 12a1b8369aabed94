#!/bin/bash
# One-GPU pass exactly as the driver runs it: whole GPU suite in one process, smoke, default bench, reference arm, and
# (after the plain bench exited 0) the ncu launch list of the same command.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-500 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
