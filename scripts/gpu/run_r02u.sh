#!/bin/bash
# r02u: final tree: whole GPU suite, smoke, bench (all legs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02u_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02u_tests.log
tail -4 gpurun_out/r02u_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02u_smoke.log 2>&1; tail -2 gpurun_out/r02u_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
echo "bench exit $?"
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-other-configs --no-tile-products --no-composite > gpurun_out/r02u_bench_k40.json 2> gpurun_out/r02u_bench_k40.err
python - <<'PY'
import json
for f in ("r02u_bench","r02u_bench_k40"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "chain us", round(d["roofline"]["us_per_launch"],1), "frac", round(d["roofline"]["frac"],3), "sust", (d.get("sustained") or {}).get("value"))
PY
