#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
