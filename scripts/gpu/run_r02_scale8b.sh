#!/bin/bash
# r02 scaling on ONE 8-GPU box: the bench line at N = 1, 2, 4, 8 back to back (N = 8 with the other configurations)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-tile-products --no-other-configs --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
echo "n1 rc=$?"
for N in 2 4; do
  timeout 300 $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --no-tile-products --no-other-configs > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  echo "n$N rc=$?"
done
NCCL_DEBUG=WARN timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-tile-products > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
echo "n8 rc=$?"; tail -c 300 gpurun_out/r02_scale_n8.err
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("sustained", {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
