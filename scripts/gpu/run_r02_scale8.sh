#!/bin/bash
# r02 scaling on ONE 8-GPU box: exchange check at 8 ranks, then the bench line at N = 8, 4, 2, 1 (same box, back to back)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29521 tests/mgpu_check_ddp.py > gpurun_out/r02_ddp_check_n8.json 2> gpurun_out/r02_ddp_check_n8.err
echo "ddp check n8 rc=$?"; tail -c 1200 gpurun_out/r02_ddp_check_n8.json
NCCL_DEBUG=WARN timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-tile-products > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err
echo "n8 rc=$?"; tail -c 300 gpurun_out/r02_scale_n8.err
for N in 4 2; do
  timeout 300 $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --no-tile-products --no-other-configs > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  echo "n$N rc=$?"
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-tile-products --no-other-configs --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
echo "n1 rc=$?"
BN_NO_P2P=1 timeout 300 $TR --nproc-per-node 8 --master-port 29529 bench.py --gpus 8 --steps 20 --warmup 5 --no-tile-products --no-other-configs > gpurun_out/r02_scale_n8_nccl.json 2> gpurun_out/r02_scale_n8_nccl.err
echo "n8 nccl rc=$?"
python - <<'PY'
import json
for n in (1, 2, 4, 8, "8_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("sustained", {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
