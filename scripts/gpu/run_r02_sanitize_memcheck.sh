#!/bin/bash
# compute-sanitizer memcheck over one small pass of every kernel family (after the plain run exited 0)
mkdir -p gpurun_out
python scripts/sanitize_smoke.py > gpurun_out/r02_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/r02_sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r02_sanitizer_memcheck.txt
