"""CPU, world_size 2 over gloo: the data-parallel rule of the training driver —
rays sharded by rank, ONE all-reduce of the flat gradient bucket, scale 1/world — reproduces the
single-process gradient of the concatenated batch.  The per-shard compute is the oracle here (the
CUDA path needs a GPU); what is under test is the host-side sharding / reduction logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.synth import make_rays
    from brdf_nerf_b200.train import allreduce_grads_
    from oracle import losses_torch as LT
    from oracle import render_torch as RT
    args = named_config("lambertian_ds")
    torch.manual_seed(0)
    model = load_model(args)                      # CPU module: only its flat storage is used here
    full = make_rays(n, depth_supervision=True)
    draws = RT.Draws.make(n, 64, 64, 128, seed=5, with_gt=True)
    shard = full.shard(rank, world)
    per = n // world
    sl = slice(rank * per, (rank + 1) * per)
    d = RT.Draws(u_strat=draws.u_strat[sl], noise1=draws.noise1[sl], u_pred=draws.u_pred[sl], noise2=draws.noise2[sl],
                 u_gt=draws.u_gt[sl])
    om = RT.OracleModel(model.state_dict(), args, requires_grad=True)
    # chunk-first-ray scalars: shard 1 must clamp with ITS first ray, like a per-rank reference call
    res, _, _ = RT.render_rays(om, args, shard.rays, d, mode="train", valid_depth=shard.valid_depth,
                               target_depths=shard.target_depths, target_std=shard.target_std)
    LT.train_loss(res, shard, args).backward()
    flat = model.flat_grads
    for name, p in model.named_parameters():
        p.grad.copy_(om.p[name].grad)             # p.grad is a view into the flat bucket
    scale = allreduce_grads_(flat, world)
    if rank == 0:
        out_q.put(torch.cat([p.grad.reshape(-1) for p in model.parameters()]) * scale)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradients_equal_full_batch():
    world, n = 2, 32
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single process, both shards evaluated separately (same per-call semantics), mean of the two losses
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.synth import make_rays
    from oracle import losses_torch as LT
    from oracle import render_torch as RT
    args = named_config("lambertian_ds")
    torch.manual_seed(0)
    model = load_model(args)
    full = make_rays(n, depth_supervision=True)
    draws = RT.Draws.make(n, 64, 64, 128, seed=5, with_gt=True)
    om = RT.OracleModel(model.state_dict(), args, requires_grad=True)
    total = 0
    per = n // world
    for r in range(world):
        sl = slice(r * per, (r + 1) * per)
        sh = full.shard(r, world)
        d = RT.Draws(u_strat=draws.u_strat[sl], noise1=draws.noise1[sl], u_pred=draws.u_pred[sl], noise2=draws.noise2[sl],
                     u_gt=draws.u_gt[sl])
        res, _, _ = RT.render_rays(om, args, sh.rays, d, mode="train", valid_depth=sh.valid_depth,
                                   target_depths=sh.target_depths, target_std=sh.target_std)
        total = total + LT.train_loss(res, sh, args) / world
    total.backward()
    want = torch.cat([om.p[name].grad.reshape(-1) for name, _ in model.named_parameters()])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-8)
