#!/bin/bash
# ncu launch list of one step, then a full capture of the GEMM family (one of each kind)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
# step = 39 GEMM launches: 8 sigma-only + 8 full fwd + bias + heads-sin + 2 skinny + bwd chain
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s ${NCU_SKIP:-53} -c ${NCU_COUNT:-12} -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out | tail -8
