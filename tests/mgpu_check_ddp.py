"""Multi-GPU check (test infrastructure; not collected by pytest, which the driver runs on one GPU).
Gradient exchange over NVLink peer memory (brdf_nerf_b200.ddp.PeerExchange / csrc/ddp.cu):
  1. bn_allreduce_p2p == the rank-ordered sum of the ranks' buckets, bit for bit, and == NCCL all_reduce to fp32 rounding;
     repeated calls (epoch counters), every rank ends with the same bucket;
  2. a graph-captured data-parallel Trainer step (exchange + Adam inside the graph) keeps the replicas bit-identical and
     follows the torch.distributed.all_reduce path (BN_NO_P2P=1) to rounding;
  3. step time of both paths.
Rank 0 prints one JSON line.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_check_ddp.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brdf_nerf_b200 import ddp  # noqa: E402
from brdf_nerf_b200.config import named_config  # noqa: E402
from brdf_nerf_b200.models import load_model  # noqa: E402
from brdf_nerf_b200.synth import make_rays  # noqa: E402
from brdf_nerf_b200.train import Trainer  # noqa: E402


def timed(fn, steps, warm, dev):
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    # ---- 1. the exchange kernel alone
    n = 2_297_000 // 8 * 8
    ex = ddp.PeerExchange(n, dev)
    assert ex.ok, f"peer mapping failed: {ex.error}"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    worst = 0.0
    for it in range(5):
        mine = torch.randn(n, device=dev, generator=g)
        ex.bucket.copy_(mine)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        want = parts[0].clone()
        for p in range(1, world):
            want += parts[p]
        nccl = mine.clone()
        dist.all_reduce(nccl)
        torch.cuda.synchronize()
        ex.all_reduce_()
        torch.cuda.synchronize()
        assert torch.equal(ex.bucket, want), f"rank {rank} call {it}: p2p sum differs from the rank-ordered sum by {(ex.bucket - want).abs().max().item()}"
        worst = max(worst, (ex.bucket - nccl).abs().max().item())
    out["p2p_vs_rank_ordered_sum"] = "bit-exact (5 calls)"
    out["p2p_vs_nccl_max_abs"] = worst
    ms_p2p = timed(ex.all_reduce_, 50, 5, dev)
    buf = torch.randn(n, device=dev)
    ms_nccl = timed(lambda: dist.all_reduce(buf), 50, 5, dev)
    out["exchange_us"] = {"p2p_kernel": ms_p2p * 1e3, "nccl_all_reduce": ms_nccl * 1e3, "bucket_MB": n * 4 / 1e6}
    ex.close()
    # ---- 2. + 3. data-parallel training step: peer exchange inside the graph vs NCCL outside
    args = named_config("lambertian_ds")
    res = {}
    for tag, env in (("p2p_in_graph", None), ("nccl_eager", "1")):
        if env:
            os.environ["BN_NO_P2P"] = env
        else:
            os.environ.pop("BN_NO_P2P", None)
        torch.manual_seed(0)
        model = load_model(args, precision="bf16").to(dev)
        tr = Trainer(model, args, world_size=world, use_graph=True)
        assert (tr._exchange is not None) == (env is None)
        batch = make_rays(1024, seed=20240912 + rank, depth_supervision=True).to(dev)
        losses = [float(tr.step(batch)) for _ in range(4)]
        flat = model.flat_params.clone()
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        same = all(torch.equal(parts[0], p) for p in parts[1:])
        ms = timed(lambda: tr.step(batch), 40, 5, dev)
        res[tag] = {"losses": losses, "replicas_bit_identical": bool(same), "ms_per_step": ms,
                    "graph_holds_whole_step": bool(tr.whole_step_graph), "params_after_4": flat}
        del tr, model
        torch.cuda.empty_cache()
    a, b = res["p2p_in_graph"].pop("params_after_4"), res["nccl_eager"].pop("params_after_4")
    out["train_step"] = res
    out["params_p2p_vs_nccl_max_abs_after_4_steps"] = (a - b).abs().max().item()
    assert res["p2p_in_graph"]["replicas_bit_identical"], "replicas diverged"
    assert all(abs(x - y) < 5e-3 for x, y in zip(res["p2p_in_graph"]["losses"], res["nccl_eager"]["losses"])), res
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
