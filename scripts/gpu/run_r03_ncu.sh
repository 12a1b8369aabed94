#!/bin/bash
# r03 ncu evidence on the final kernels: (1) launch list of the bench step, (2) --set full of the fused trunk forward, the fused
# data-gradient chain and the weight-gradient GEMM.  Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products --no-other-configs --sustain 0 --cooldown 0"
$CMD > gpurun_out/r03_ncu_plain.log 2>&1 || { tail -5 gpurun_out/r03_ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r03_launches_raw.csv $CMD > gpurun_out/r03_ncu1.log 2>&1
python scripts/summarize_launches.py gpurun_out/r03_launches_raw.csv > gpurun_out/r03_launches_step_summary.txt; head -12 gpurun_out/r03_launches_step_summary.txt | cut -c1-140
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"train_chain_kernel|dgrad_chain_kernel|EpiWgradT" -s 8 -c 5 -o gpurun_out/r03_prof_tc -f $CMD > gpurun_out/r03_ncu2.log 2>&1
tail -1 gpurun_out/r03_ncu2.log | cut -c1-200
ncu -i gpurun_out/r03_prof_tc.ncu-rep --page raw --csv > gpurun_out/r03_prof_tc_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py gpurun_out/r03_prof_tc_raw.csv > gpurun_out/r03_ncu_full_tensor_kernels.csv; cut -c1-400 gpurun_out/r03_ncu_full_tensor_kernels.csv
rm -f gpurun_out/r03_prof_tc.ncu-rep
