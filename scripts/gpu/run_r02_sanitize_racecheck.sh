#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the same pass
mkdir -p gpurun_out
python scripts/sanitize_smoke.py 64 > gpurun_out/r02_sanitize_plain2.log 2>&1 || { tail -5 gpurun_out/r02_sanitize_plain2.log; exit 1; }
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_smoke.py 64 > gpurun_out/r02_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -25 gpurun_out/r02_sanitizer_racecheck.txt
