#!/bin/bash
# DSM kernels: parity tests, micro-benchmark, then (after the plain run exited 0) the ncu launch list and one --set full capture.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dsm.py -q -m gpu -x --tb=short 2>&1 | tail -5
timeout 300 python scripts/bench_dsm.py > gpurun_out/bench_dsm.json 2> gpurun_out/bench_dsm.err || exit 1
cut -c1-1200 gpurun_out/bench_dsm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dsm_ -c 60 --csv --log-file gpurun_out/dsm_launches.csv python scripts/bench_dsm.py 64 64 > gpurun_out/ncu_dsm1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_ -s 20 -c 8 -o gpurun_out/prof_dsm -f python scripts/bench_dsm.py > gpurun_out/ncu_dsm2.log 2>&1
ncu -i gpurun_out/prof_dsm.ncu-rep --page raw --csv > gpurun_out/prof_dsm_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
