#!/bin/bash
# r03 final: whole GPU suite, smoke, the driver's own bench commands (ours + reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r03_final_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r03_final_tests.log; tail -4 gpurun_out/r03_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r03_final_smoke.log 2>&1; tail -2 gpurun_out/r03_final_smoke.log
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03_final_bench.json 2> gpurun_out/r03_final_bench.err; echo "bench exit $?"
timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03_final_bench_reference.json 2> gpurun_out/r03_final_bench_reference.err; echo "reference arm exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03_final_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "chain us", round(d["roofline"]["us_per_launch"],1), "frac", round(d["roofline"]["frac"],3), "all", round(d["roofline"]["all_tcgen05"]["frac"],3), "sust", (d.get("sustained") or {}).get("value"))
print("cpu_baseline", d.get("cpu_baseline"))
print("composite", {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get("roofline_composite") or {}).items() if k in ("frac","achieved","peak","traffic")})
r=json.loads(open("gpurun_out/r03_final_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("unit"), r.get("cpu_baseline"))
PY
