#!/bin/bash
# bench (eager + CUDA graph), then the ncu launch list and one full capture of the GEMM kernel
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; tail -2 gpurun_out/bench_eager.err
python bench.py --steps 20 --warmup 3 --graph 1 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -2 gpurun_out/bench_graph.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -20
