#!/bin/bash
# r02o: what bounds the weight-gradient GEMM (timing experiments), roofline leg with its own warm-up
mkdir -p gpurun_out
timeout 400 python scripts/ab_wgrad.py 1024 > gpurun_out/r02o_ab.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --sustain 0 --no-cpu-baseline --no-other-configs --no-composite --no-tile-products > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
cat gpurun_out/r02o_ab.txt | tail -30
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02o_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"]), d["gpu_launches"], round(d["roofline"]["us_per_launch"],1), d["roofline"]["frac"])
PY
