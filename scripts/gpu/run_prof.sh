#!/bin/bash
# ncu launch list of one training step (graph replay), then ONE --set full capture of the dominant kernels:
# the fused trunk forward (train_chain_kernel), one trunk dgrad and one trunk wgrad GEMM.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
# -k matches the kernel's base name (no template arguments)
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:train_chain_kernel -s 4 -c 2 -o gpurun_out/prof_chain -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s ${NCU_SKIP:-108} -c ${NCU_COUNT:-16} -o /tmp/prof_gemm -f $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log | cut -c1-200
# gpurun_out/ is capped at 64 MiB: the GEMM report stays on the box, its raw page travels as CSV
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
