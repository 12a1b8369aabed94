#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_wgrad.py 1024 > gpurun_out/r02p_trace_wgrad.txt 2>&1
cat gpurun_out/r02p_trace_wgrad.txt | tail -12
