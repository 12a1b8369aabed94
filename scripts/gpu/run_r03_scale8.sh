#!/bin/bash
# r03 scaling on ONE 8-GPU box with the final kernels: bench line at N = 1, 2, 4, 8 back to back + the 8-rank exchange check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
FAST="--steps 20 --warmup 5 --no-tile-products --no-other-configs --no-cpu-baseline --no-composite --sustain 1"
timeout 200 python bench.py --gpus 1 $FAST > gpurun_out/r03_scale_n1.json 2> gpurun_out/r03_scale_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port 2954$N bench.py --gpus $N $FAST > gpurun_out/r03_scale_n$N.json 2> gpurun_out/r03_scale_n$N.err
  echo "n$N rc=$?"
done
timeout 300 $TR --nproc-per-node 8 --master-port 29549 tests/mgpu_check_ddp.py > gpurun_out/r03_ddp_check_n8.json 2> gpurun_out/r03_ddp_check_n8.err; echo "ddp rc=$?"; tail -c 600 gpurun_out/r03_ddp_check_n8.json
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r03_scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), (d.get("sustained") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
