#!/bin/bash
# r02y: Lambertian K-C in one launch (compositing + colour + loss + backward): tests, bench, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02y_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02y_tests.log
tail -6 gpurun_out/r02y_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err
echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-composite --no-tile-products --no-other-configs --sustain 0 --cooldown 0"
$CMD > gpurun_out/r02y_ncu_plain.log 2>&1 || { tail -5 gpurun_out/r02y_ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r02y_launches_raw.csv $CMD > gpurun_out/r02y_ncu1.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02y_launches_raw.csv > gpurun_out/r02y_launches_step_summary.txt; head -44 gpurun_out/r02y_launches_step_summary.txt | cut -c1-150
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02y_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "chain us", round(d["roofline"]["us_per_launch"],1), "frac", round(d["roofline"]["frac"],3), "sust", (d.get("sustained") or {}).get("value"))
PY
