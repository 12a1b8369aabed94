#!/bin/bash
# r02s: two MMA issuer threads in every tcgen05 kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "gemm or mlp or chain or bf16" > gpurun_out/r02s_tests_quick.log 2>&1
echo "quick tests exit $?" >> gpurun_out/r02s_tests_quick.log
tail -5 gpurun_out/r02s_tests_quick.log
timeout 300 python scripts/trace_wgrad.py 1024 > gpurun_out/r02s_trace_wgrad.txt 2>&1; tail -4 gpurun_out/r02s_trace_wgrad.txt
timeout 300 python scripts/ab_wgrad.py 1024 > gpurun_out/r02s_ab.txt 2>&1; grep "base" gpurun_out/r02s_ab.txt
timeout 300 python scripts/trace_chain.py train > gpurun_out/r02s_trace_train.txt 2>&1; tail -12 gpurun_out/r02s_trace_train.txt | cut -c1-220
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-tile-products > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02s_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "chain us", round(d["roofline"]["us_per_launch"],1), "frac", round(d["roofline"]["frac"],3), "sust", d.get("sustained",{}).get("value"))
print(d["roofline"]["all_tcgen05"]["by_kind"])
PY
