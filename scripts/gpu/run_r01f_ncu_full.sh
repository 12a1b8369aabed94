#!/bin/bash
# ONE --set full capture of the step's dominant kernels (after the plain command exited 0): the fused trunk forward
# (train_chain_kernel, both launches of a step) and a handful of gemm_tc_kernel launches (dgrad + wgrad of the trunk).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_chain_kernel -s 4 -c 2 -o gpurun_out/prof_chain -f $CMD > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s ${NCU_SKIP:-108} -c ${NCU_COUNT:-8} -o /tmp/prof_gemm -f $CMD > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/ncu3.log | cut -c1-200
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py gpurun_out/prof_chain_raw.csv | cut -c1-400
python scripts/summarize_ncu_raw.py gpurun_out/prof_gemm_raw.csv | cut -c1-300
