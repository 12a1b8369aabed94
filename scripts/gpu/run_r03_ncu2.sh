#!/bin/bash
# r03 ncu --set full of the remaining kernels of the step: weight-gradient GEMM, fused per-ray kernels, feature-layer GEMMs
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products --no-other-configs --sustain 0 --cooldown 0"
$CMD > gpurun_out/r03_ncu_plain2.log 2>&1 || { tail -5 gpurun_out/r03_ncu_plain2.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"EpiWgradT|lambertian_render_loss|coarse_to_fine|EpiBiasT|EpiSinT|EpiHeadsOut" -s 30 -c 16 -o gpurun_out/r03_prof_rest -f $CMD > gpurun_out/r03_ncu3.log 2>&1
tail -1 gpurun_out/r03_ncu3.log | cut -c1-200
ncu -i gpurun_out/r03_prof_rest.ncu-rep --page raw --csv > gpurun_out/r03_prof_rest_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py gpurun_out/r03_prof_rest_raw.csv > gpurun_out/r03_ncu_full_other_kernels.csv; cut -c1-330 gpurun_out/r03_ncu_full_other_kernels.csv | head -20
rm -f gpurun_out/r03_prof_rest.ncu-rep
