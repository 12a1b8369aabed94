#!/bin/bash
# r02 ncu evidence: (1) launch list of the bench step, (2) --set full of the fused trunk forward and the fused data-gradient
# chain, (3) --set full of the K-C kernels (compositing / shading / per-sample BRDF / permutation) at 65 536 rays.
# Every ncu run follows a plain run of the same command that exited 0 (all ncu runs of one call count as one).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products --no-other-configs --sustain 0"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 || { tail -5 gpurun_out/r02_ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 420 --csv --log-file gpurun_out/r02_launches_raw.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
tail -1 gpurun_out/r02_ncu1.log | cut -c1-200
python scripts/summarize_launches.py gpurun_out/r02_launches_raw.csv > gpurun_out/r02_launches_step_summary.txt; head -30 gpurun_out/r02_launches_step_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"train_chain_kernel|dgrad_chain_kernel" -s 6 -c 3 -o gpurun_out/r02_prof_chain -f $CMD > gpurun_out/r02_ncu2.log 2>&1
tail -1 gpurun_out/r02_ncu2.log | cut -c1-200
ncu -i gpurun_out/r02_prof_chain.ncu-rep --page raw --csv > gpurun_out/r02_prof_chain_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py gpurun_out/r02_prof_chain_raw.csv > gpurun_out/r02_ncu_full_chain_kernels.csv; cut -c1-300 gpurun_out/r02_ncu_full_chain_kernels.csv
KC="python scripts/prof_kc.py 65536"
$KC > gpurun_out/r02_kc_plain.log 2>&1 || { tail -5 gpurun_out/r02_kc_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none -k regex:"composite_|shade_rays|brdf_points|permute_rows" -c 40 -o gpurun_out/r02_prof_kc -f $KC > gpurun_out/r02_ncu3.log 2>&1
tail -1 gpurun_out/r02_ncu3.log | cut -c1-200
ncu -i gpurun_out/r02_prof_kc.ncu-rep --page raw --csv > gpurun_out/r02_prof_kc_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py gpurun_out/r02_prof_kc_raw.csv > gpurun_out/r02_ncu_full_kc_kernels.csv; cut -c1-220 gpurun_out/r02_ncu_full_kc_kernels.csv
rm -f gpurun_out/r02_prof_kc.ncu-rep
