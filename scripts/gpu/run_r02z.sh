#!/bin/bash
# r02z: lean tcgen05 issue loop (whole warp, elected lane, descriptor increments) in every tcgen05 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02z_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02z_tests.log
tail -5 gpurun_out/r02z_tests.log
timeout 300 python scripts/trace_wgrad.py 1024 > gpurun_out/r02z_trace_wgrad.txt 2>&1; tail -4 gpurun_out/r02z_trace_wgrad.txt
timeout 300 python scripts/trace_chain.py train > gpurun_out/r02z_trace_train.txt 2>&1; tail -6 gpurun_out/r02z_trace_train.txt | cut -c1-230
timeout 300 python scripts/ab_wgrad.py 1024 > gpurun_out/r02z_ab.txt 2>&1; grep "base" gpurun_out/r02z_ab.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "chain us", round(d["roofline"]["us_per_launch"],1), "frac", round(d["roofline"]["frac"],3), "sust", (d.get("sustained") or {}).get("value"))
print(d["roofline"]["all_tcgen05"]["frac"], d["roofline"]["all_tcgen05"]["by_kind"])
PY
