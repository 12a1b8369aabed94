#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: train rays/s of the SpS-BRDF-NeRF hot path.

Workload (config.workload): BASELINE.json configs[1] — Lambertian pre-train step with depth
supervision (ds_lambda=10, --mapping), 1024 rays/GPU x (64 stratified + 64 guided) samples,
synthetic 3-view satellite rays, random-init weights (seed 0).  One step = render_rays forward
(both MLP passes, sampler, compositing) + colour/depth loss + backward + [NCCL all-reduce of the flat
gradient bucket] + fused Adam.  Weak scaling: every rank processes its own 1024-ray batch.

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA, bf16 tcgen05 MLP)
    python bench.py --impl reference ...                     # the reference algorithm on host cores
Under torchrun (N > 1) one process per GPU; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAYS_PER_GPU = 1024
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel of the step, chain::train_chain_kernel at
# P = 65536 points (two launches per step: stratified points, guided points), from the ncu --set full capture under profiles/
NCU_CHAIN_DRAM_BYTES_PER_LAUNCH = 4.7e6 + 1026.3e6
NCU_TRAFFIC_NOTE = ("ncu --set full (profiles/r03_ncu_full_tensor_kernels.csv, P = 65536): 4.7 MB read + 1026.3 MB written per launch "
                    "= the algorithmic 1.07 GB (h_l, c_l of 8 layers + encoding, bf16; nothing is read back; the weights stay in L2). "
                    "fused data-gradient chain (dgrad_chain_kernel, P = 131072, same capture): 1079 MB read + 889 MB written vs "
                    "1073 + 939 MB algorithmic (c_l of 7 layers + dZ_7 in, dZ_l of 7 layers out; the tail of the writes is still in L2); "
                    "trunk wgrad GEMM (profiles/r01e_ncu_full_gemm.csv): 269.5 MB read vs 256 MiB algorithmic")
# K-C kernels at 65 536 rays x 128 samples, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r02_ncu_full_kc_kernels.csv)
NCU_KC_DRAM_BYTES = {"composite_fwd C=4": 167.8e6 + 69.6e6, "composite_bwd C=4": 270.0e6 + 105.9e6,
                     "composite_fwd C=16": 570.4e6 + 94.8e6, "composite_bwd C=16": 676.0e6 + 490.3e6}
METRIC = "train rays/s (SpS-BRDF-NeRF, 1/2/4/8 B200); MLP tensor-pipe %; composite GB/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def mlp_flops_per_ray(args, n_coarse, n_full, train=True, shared_trunk=False):
    """Algorithmic MLP FLOPs of one ray: sigma-only pass over n_coarse points, full pass
    over n_full points, backward (2x) through the full pass only.  shared_trunk: what this library executes — the
    stratified points' trunk is evaluated once and shared by both passes (the reference evaluates it twice)."""
    F, L = args.fc_feat, args.fc_layers
    enc = 60 if args.mapping else 3
    trunk = enc * F + (L - 2) * F * F + (F + enc) * F
    sig, feat, col = F, F * F, F * (F // 2) + (F // 2) * 3
    fwd = n_coarse * ((0 if shared_trunk else trunk) + sig) + n_full * (trunk + sig + feat + col)
    bwd = 2 * n_full * (trunk + sig + feat + col) if train else 0
    return 2.0 * (fwd + bwd)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """Number of samples written so far (brackets the timed region inside the continuous log)."""
        try:
            self.f.flush()
            return sum(1 for _ in open(self.f.name))
        except Exception:
            return 0

    def stop(self, first=0, last=None):
        if self.p is None:
            self.rows = None
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        return self.summary(first, last)

    def summary(self, first=0, last=None):
        """Clocks / power / throttle reasons of the samples [first, last) of the log (after stop())."""
        if getattr(self, "rows", None) is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        # one sample of slack on either side: nvidia-smi stamps are ~20 ms apart
        rows = self.rows[max(0, first - 1):(None if last is None else last + 1)] or self.rows
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "power_w_median": statistics.median(pw) if pw else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def reference_available():
    """The unmodified reference (live tree in the build container, oracle/_ref staged by oracle/stage_ref.py on the GPU box)."""
    try:
        from oracle import ref_harness as RH
        return RH.kind() != "absent"
    except Exception:                                                    # noqa: BLE001
        return False


def cpu_reference_rate(args, rays_per_step, steps, warmup, seed=0, device="cpu"):
    """rays/s of a full training step (render_rays + losses + backward + Adam) of the reference algorithm on `device`.
    kind "reference": the UNMODIFIED reference files (oracle/ref_step.py drives rendering.render_rays, metrics.SNerfLoss /
    DepthLoss and torch.optim.Adam exactly as main.py:194-268 does); kind "port": the oracle restatement, only when the
    reference files are not available.  Returns (rays/s, ms/step, threads, kind)."""
    from brdf_nerf_b200.synth import make_rays
    torch.set_num_threads(os.cpu_count() or 1)
    batch = make_rays(rays_per_step, depth_supervision=True)
    if reference_available():
        from oracle import ref_step
        v, ms = ref_step.rate(args, batch, steps, warmup, device=device)
        return v, ms, torch.get_num_threads(), "reference"
    if device != "cpu":
        raise RuntimeError("the oracle port runs on the CPU only")
    from brdf_nerf_b200.models import load_model
    from oracle import losses_torch as LT
    from oracle import render_torch as RT
    torch.manual_seed(seed)
    state = load_model(args).state_dict()
    om = RT.OracleModel(state, args, requires_grad=True)
    opt = torch.optim.Adam(om.parameters(), lr=args.lr)
    S1, G = args.n_samples, args.guided_samples
    times = []
    for i in range(warmup + steps):
        draws = RT.Draws.make(rays_per_step, S1, G, S1 + G, seed=1234 + i, with_gt=True)
        t0 = time.perf_counter()
        res, _, _ = RT.render_rays(om, args, batch.rays, draws, mode="train", valid_depth=batch.valid_depth,
                                   target_depths=batch.target_depths, target_std=batch.target_std)
        loss = LT.train_loss(res, batch, args)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return rays_per_step / (ms / 1e3), ms, torch.get_num_threads(), "port"


def tile_products_leg(cpu=True):
    """The callers either side of the hot path at BASELINE config 5 size, one 2048 x 2048 tile: RPC camera ->
    ray records and rendered depth -> DSM, per-kernel time and algorithmic GB/s (scripts/bench_georays.py, bench_dsm.py),
    with the reference's CPU path (oracle port: numpy + the C restatement of plyflatten, one thread like the reference)
    timed beside them on a bounded sample.  Never fails the bench line: an error is reported as a string."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_dsm
        import bench_georays
        out = {"dsm": bench_dsm.run(2048, 2048), "georays": bench_georays.run(2048, 2048)}
        if cpu:
            import numpy as np
            from brdf_nerf_b200.synth import SCENE_CENTER, SCENE_RANGE, make_tile_rays, tile_surface_depth
            from oracle import dsm_np as D
            from oracle import georays_np as G
            h = w = 1024                                           # a quarter tile: ~0.2 s of CPU work
            rays = make_tile_rays(h, w, view=0)
            depth = tile_surface_depth(rays)
            t0 = time.perf_counter()
            D.dsm_from_nerf_prediction(rays.numpy(), depth.numpy(), SCENE_RANGE * w / 2048, SCENE_CENTER)
            dt = time.perf_counter() - t0
            out["dsm"]["cpu_baseline"] = {"value": h * w / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                                          "sample": "1024 x 1024 tile: numpy float64 cloud + grid + C restatement of plyflatten"}
            n = 200_000
            idx = np.arange(0, 2048 * 2048, (2048 * 2048) // n)[:n]
            o = G.synthetic_rpc(0)
            t0 = time.perf_counter()
            G.get_rays((idx % 2048).astype(np.float64), (idx // 2048).astype(np.float64), o, -25.0, 95.0, cs="utm")
            dt = time.perf_counter() - t0
            out["georays"]["cpu_baseline"] = {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                                              "sample": f"{n} strided pixels of the 2048 x 2048 image, numpy float64 restatement of rpcm + get_rays"}
        return out
    except Exception as e:                                          # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


WORKLOAD = ("spsbrdf-nerf Lambertian pretrain + depth supervision (ds_lambda=10, --mapping), "
            "1024 rays x (64+64) samples per GPU per step, fc 8x512, random init seed 0")


def run_reference(opts):
    """The reference arm: the reference's own implementation of the path on this box's host cores, all threads, same
    workload (1024 rays per step, so that `config` equals our arm's)."""
    rank, _, world = _dist_env()
    if rank != 0:
        return
    from brdf_nerf_b200.config import named_config
    args = named_config("lambertian_ds")
    value, ms, cores, kind = cpu_reference_rate(args, RAYS_PER_GPU, opts.steps, opts.warmup)
    what = ("the UNMODIFIED reference files (oracle/_ref: rendering.render_rays + metrics.SNerfLoss / DepthLoss + torch.optim.Adam)"
            if kind == "reference" else "torch CPU fp32 oracle port (reference files not staged)")
    sample = f"{opts.steps} steps of {RAYS_PER_GPU} rays after {opts.warmup} warm-up, {what}, torch CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": opts.gpus, "steps": opts.steps,
            "warmup": opts.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "global_rays": RAYS_PER_GPU,
                       "sample": "every timed step is the full 1024-ray step on the host cores (rank 0 only)"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def other_configs_leg(opts, dev, rank, world, barrier):
    """The other BASELINE.json configurations on the same launch (every rank takes part: the training ones all-reduce):
      configs[2]  BRDF stage RPV111 + analytic normals (second-order backward) + cos_irra_on, 1024 rays per GPU (weak)
      configs[3]  Hapke b,c,theta / microfacet, 8192 rays per step ray-sharded over the ranks (strong: 8192 / N per GPU)
      configs[4]  inference RGB + depth + normals + albedo in chunks of 5120 rays (eval.py:56-76), rays/s per job (weak)
    A few graph-replayed steps each, device-timed, max over ranks."""
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.rendering import render_rays
    from brdf_nerf_b200.synth import make_rays
    from brdf_nerf_b200.train import Trainer
    out = {}

    def timed(fn, steps, warmup=3):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t[0])

    def train(key, cfg, rays_per_gpu, steps, scaling, **kw):
        args = named_config(cfg)
        torch.manual_seed(0)
        model = load_model(args, precision=opts.precision).to(dev)
        tr = Trainer(model, args, world_size=world, use_graph=bool(opts.graph))
        batch = make_rays(rays_per_gpu, seed=20240912 + rank).to(dev)
        ms = timed(lambda: tr.step(batch, **kw), steps)
        total = rays_per_gpu * world
        out[key] = {"model": cfg, "kwargs": kw, "rays_per_gpu": rays_per_gpu, "global_rays": total, "n_gpus": world,
                    "scaling": scaling, "steps": steps, "ms_per_step": ms, "rays_per_s": total / ms * 1e3,
                    "gpu_launches_per_step": int(getattr(tr, "graph_launches", 0))}
        del tr, model, batch
        torch.cuda.empty_cache()

    train("configs[2] BRDF stage RPV111 funcM/F/H=1, normal=analystic, cos_irra_on", "rpv111", RAYS_PER_GPU, 10, "weak",
          apply_brdf=True, cos_irra_on=True)
    per = max(128, (8192 // world) // 128 * 128)
    train("configs[3] Hapke b,c,theta, 8192 rays ray-sharded", "hapke_bct", per, 4, "strong",
          apply_brdf=True, apply_theta=True, cos_irra_on=True)
    train("configs[3] microfacet, 8192 rays ray-sharded", "microfacet", per, 4, "strong", apply_brdf=True, cos_irra_on=True)
    # inference: independent chunks, no collective
    args = named_config("rpv111")
    torch.manual_seed(0)
    model = load_model(args, precision=opts.precision).to(dev)
    chunk, n_chunks = int(args.chunk), 4
    rays = make_rays(chunk * n_chunks, seed=rank).rays.to(dev)

    def infer():
        with torch.no_grad():
            for c in range(n_chunks):
                render_rays({"coarse": model}, args, rays[c * chunk:(c + 1) * chunk], None, mode="test", apply_brdf=True,
                            cos_irra_on=True)

    ms = timed(infer, 3, warmup=2)
    rps = chunk * n_chunks * world / ms * 1e3
    out["configs[4] inference RGB+depth+normals+albedo, 5120-ray chunks, ray-sharded"] = {
        "model": "rpv111", "chunk_rays": chunk, "chunks_per_gpu": n_chunks, "n_gpus": world, "scaling": "weak",
        "ms_per_chunk": ms / n_chunks, "rays_per_s": rps, "tile_2048x2048_seconds": 2048 * 2048 / rps}
    del model
    torch.cuda.empty_cache()
    return out


def run_ours(opts):
    from brdf_nerf_b200 import _lib as L
    from brdf_nerf_b200.config import named_config
    from brdf_nerf_b200.models import load_model
    from brdf_nerf_b200.synth import make_rays
    from brdf_nerf_b200.train import Trainer
    rank, local, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    lib = L.load()
    clocks = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi needs ~0.5 s to produce its first line
    args = named_config("lambertian_ds")
    torch.manual_seed(0)
    model = load_model(args, precision=opts.precision).to(dev)
    trainer = Trainer(model, args, world_size=world, use_graph=bool(opts.graph))
    host_batch = make_rays(RAYS_PER_GPU, seed=20240912 + rank, depth_supervision=True).packed(pin=True)   # one pinned buffer
    batch = host_batch.to(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, opts.warmup)):
        loss = trainer.step(batch)
    if opts.graph and trainer.static_batch() is not None:
        batch = trainer.static_batch()        # inputs resident in HBM: the graph's own input buffers
    # ---- roofline leg, FIRST (right after the warm-up, on a GPU that has not yet run into its power cap: the per-launch times are
    # compared with the BURST peak, "a kernel timed alone"): per-launch CUDA-event timing of the GEMM family on the launching stream,
    # three eager steps of the same batch after three eager warm-up steps (the eager path allocates its own workspace) ----
    prof = None
    if rank == 0:
        lib.bn_profile_enable.restype = C.c_int
        eager = Trainer(model, args, world_size=1, use_graph=False)
        eager.m, eager.v, eager.step_count = trainer.m, trainer.v, trainer.step_count
        eager.step(batch)                  # allocates the eager path's workspace
        torch.cuda.synchronize()
        time.sleep(opts.cooldown)
        for _ in range(3):
            eager.step(batch)
        torch.cuda.synchronize()
        lib.bn_profile_enable(1)
        nprof = 3
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(nprof):
            eager.step(batch)
        pe1.record()
        torch.cuda.synchronize()
        ms_eager = pe0.elapsed_time(pe1) / nprof      # the profiled steps themselves (eager launches + per-launch events)
        cnt = (C.c_longlong * 4)(); tms = (C.c_double * 4)(); work = (C.c_double * 4)()
        lib.bn_profile_collect(4, cnt, tms, work)
        lib.bn_profile_enable(0)
        prof = (nprof, ms_eager, cnt, tms, work)
        del eager
    barrier()
    time.sleep(opts.cooldown)
    for _ in range(3):                     # every rank: back on the graph path before the timed legs
        loss = trainer.step(batch)
    barrier()
    # ---- device-resident timing -------------------------------------------------------------
    if clocks is not None:
        t_wait = time.time()
        while clocks.mark() == 0 and time.time() - t_wait < 3.0:
            time.sleep(0.05)
    mark0 = clocks.mark() if clocks else 0
    lc0 = lib.bn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(opts.steps):
        loss = trainer.step(batch)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / opts.steps
    # kernels of this library inside the timed region: eager launches counted by the library itself, plus the
    # kernel nodes every CUDA-graph replay executes (counted once, at capture time)
    launches = (lib.bn_launch_count() - lc0) + opts.steps * getattr(trainer, "graph_launches", 0)
    mark_main = clocks.mark() if clocks else 0
    total_rays = RAYS_PER_GPU * world
    # ---- end to end through the public API, right after the resident leg (same steps, same thermal state): every step's
    # batch starts in pinned host memory and is copied into the device (H2D), every step's loss is read back into host memory
    # (D2H), both inside the timed region.  --feed inline (default):
    # both copies are issued on the compute stream between two graph replays; --feed prefetch: Trainer.prefetch() copies batch
    # k+1 on the trainer's copy stream while step k computes and read_loss_async() moves the loss on that same stream
    # (measured on B200, profiles/r01f_exp_e2e.log: no gain, the two copies cost ~5 us each; the e2e leg differs from the
    # resident leg mainly by running later, on a GPU that has reached its power cap).
    # The host awaits the loss of step k after step k+1 has been enqueued, the way a training loop logs without stalling.
    pin = torch.empty(4, dtype=torch.float32).pin_memory()
    evs = [None] * 4
    barrier()
    time.sleep(opts.cooldown)
    for _ in range(3):                       # warm-up of THIS path (first host-fed step: staging copies, pinned read-back)
        loss = trainer.step(trainer.prefetch(host_batch) if opts.feed == "prefetch" else
                            (host_batch if opts.graph else host_batch.to(dev, non_blocking=True)))
        pin[0:1].copy_(loss.reshape(1), non_blocking=True)
    barrier()
    e0.record()
    prev = None
    loss_host = float("nan")
    if opts.feed == "prefetch":
        staged = trainer.prefetch(host_batch)
        for i in range(opts.steps):
            loss = trainer.step(staged)
            if i + 1 < opts.steps:
                staged = trainer.prefetch(host_batch)
            evs[i % 4] = trainer.read_loss_async(loss, pin[i % 4:i % 4 + 1])
            if prev is not None:
                evs[prev].synchronize()
                loss_host = float(pin[prev])
            prev = i % 4
    else:
        for i in range(opts.steps):
            loss = trainer.step(host_batch if opts.graph else host_batch.to(dev, non_blocking=True))
            pin[i % 4:i % 4 + 1].copy_(loss.reshape(1), non_blocking=True)
            evs[i % 4] = torch.cuda.Event()
            evs[i % 4].record()
            if prev is not None:
                evs[prev].synchronize()
                loss_host = float(pin[prev])
            prev = i % 4
    evs[prev].synchronize()
    loss_host = float(pin[prev])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / opts.steps
    # ---- roofline numbers from the profiled eager steps taken before the timed legs ----
    roof = None
    if rank == 0:
        peaks = _peaks()
        nprof, ms_eager, cnt, tms, work = prof
        gemm_ms = sum(tms) / nprof
        alg = mlp_flops_per_ray(args, args.n_samples, args.n_samples + args.guided_samples) * RAYS_PER_GPU
        alg_shared = mlp_flops_per_ray(args, args.n_samples, args.n_samples + args.guided_samples, shared_trunk=True) * RAYS_PER_GPU
        gemm_flops = sum(work) / nprof
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        kind = lambda i: {"launches": cnt[i] / nprof, "ms": tms[i] / nprof, "tflops": work[i] / max(tms[i], 1e-9) / 1e9}
        chain_tf = work[2] / max(tms[2], 1e-9) / 1e9
        # the dominant kernel of the step: the fused PE + 8-layer SIREN trunk forward (chain::train_chain_kernel)
        # denominator: the BURST cuBLAS figure — the profiled region is three eager steps (~10 ms), nowhere near the seconds it
        # takes to reach the power cap; the fraction against the sustained figure is given beside it
        roof = {"bound": "tensor", "kernel": "bn::chain::train_chain_kernel (fused PE + 8 SIREN layers, forward, writes h_l / cos_l for the backward)",
                "achieved": chain_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": chain_tf / peaks["tf_burst"],
                "frac_of_sustained_peak": chain_tf / peaks["tf_sust"],
                "peak_kind": f"{peaks['src']} burst bf16 cuBLAS (kernel timed per launch inside a ~10 ms region; sustained peak {peaks['tf_sust']})",
                "traffic": NCU_CHAIN_DRAM_BYTES_PER_LAUNCH, "traffic_note": NCU_TRAFFIC_NOTE,
                "launches_per_step": cnt[2] / nprof, "us_per_launch": 1e3 * tms[2] / max(cnt[2], 1),
                "share_of_step": tms[2] / nprof / ms,      # of the timed (graph-replayed) step; compare with the ncu launch list

                "all_tcgen05": {"kernel": "every tcgen05 launch of a step (chain kernels + gemm_tc_kernel fwd/dgrad/wgrad)",
                                "achieved": achieved, "frac": achieved / peaks["tf_burst"], "frac_of_sustained_peak": achieved / peaks["tf_sust"],
                                "launches_per_step": sum(cnt) / nprof, "ms_per_step": gemm_ms, "share_of_step": gemm_ms / ms,
                                "share_note": "per-launch times are taken with the backward serialised (no concurrent weight-gradient stream), "
                                              "the timed step overlaps them: the sum may exceed the step",
                                "profiled_step_ms": ms_eager,
                                "executed_gflop_per_step": gemm_flops / 1e9, "algorithmic_gflop_per_step_shared_trunk": alg_shared / 1e9,
                                "reference_gflop_per_step": alg / 1e9,
                                "by_kind": {"chain_fwd": kind(2), "chain_dgrad": kind(3), "tn_fwd_dgrad": kind(0), "nt_wgrad": kind(1)}}}
    mark1 = clocks.mark() if clocks else 0
    # ---- sustained leg: the same resident step for >= opts.sustain seconds back to back.  The K timed steps above last a few
    # tens of milliseconds (a burst: the GPU has not reached its 1 kW power cap); a training run lives here instead.
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    ms_sus, n_sus = None, 0
    if opts.sustain > 0:
        n_sus = max(opts.steps, int(opts.sustain * 1e3 / max(ms, 1e-3)))      # from the max-over-ranks time: same count everywhere
        barrier()
        e0.record()
        for _ in range(n_sus):
            loss = trainer.step(batch)
        e1.record()
        barrier()
        ms_sus = e0.elapsed_time(e1) / n_sus
    mark2 = clocks.mark() if clocks else 0
    clk_sus = None
    if clocks:
        clocks.stop()
        clk = clocks.summary(mark0, mark_main)
        clk_sus = clocks.summary(mark1, mark2) if ms_sus is not None else None
    else:
        clk = None
    if ms_sus is not None:
        t = torch.tensor([ms_sus], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_sus = float(t[0])

    # ---- HBM leg: the compositing kernels (K-C) at inference-chunk size, GB/s against the measured copy peak ----
    roof_hbm = None
    if rank == 0 and not opts.no_composite:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_composite
        n_c = 65536
        r = bench_composite.run(n_c, 128, report=lambda *_: None)
        # lead with the kernel of the HEADLINE configuration (Lambertian forward, C = 4); byte model = SURVEY 8d: read z + packed
        # row, write alpha / T / w (32 B per sample at C = 4) + the per-ray outputs.  The backward's byte model (re-read z, row,
        # alpha, T, w; write the row gradient) is this library's own count; SURVEY's coarser 92 B/sample figure is given beside it.
        dom = r["composite_fwd C=4"]
        surv = {"composite_fwd C=4": 32, "composite_fwd C=16": 56, "composite_bwd C=4": 92, "composite_bwd C=16": 92}
        roof_hbm = {"bound": "hbm", "kernel": "bn::composite_fwd128_kernel<4> (Lambertian compositing forward: the K-C kernel of the headline configuration)",
                    "achieved": dom["gbs"], "peak": _peaks()["hbm"], "unit": "GB/s", "frac": dom["gbs"] / _peaks()["hbm"],
                    "peak_kind": f"{_peaks()['src']} HBM copy bandwidth (kernel timed alone, inputs rotated over 3 sets > L2)",
                    "traffic": NCU_KC_DRAM_BYTES["composite_fwd C=4"],
                    "traffic_note": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r02_ncu_full_kc_kernels.csv); "
                                    "below the algorithmic bytes because the tail of the writes is still in L2 when the kernel ends",
                    "algorithmic_bytes_per_launch": dom["bytes"], "us_per_launch": dom["us"],
                    "workload": f"{n_c} rays x 128 samples (one inference chunk); the 1024-ray training batch moves 7 MB and is latency bound",
                    "all": {k: {"us": v["us"], "GB/s": v["gbs"], "frac": v["frac"], "algorithmic_bytes": v["bytes"],
                                "ncu_dram_bytes": NCU_KC_DRAM_BYTES.get(k),
                                "frac_at_survey_bytes_per_sample": surv[k] * n_c * 128 / (v["us"] * 1e-6) / 1e9 / _peaks()["hbm"]}
                            for k, v in r.items()}}
    other = None
    if not opts.no_other_configs:
        del trainer
        torch.cuda.empty_cache()
        other = other_configs_leg(opts, dev, rank, world, barrier)
    if rank == 0:
        cpu = cuda_eager = None
        if world == 1 and not opts.no_cpu_baseline:
            v, cms, cores, kind = cpu_reference_rate(args, RAYS_PER_GPU, 4, 1)
            cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "ms_per_step": cms,
                   "sample": "4 steps of 1024 rays after 1 warm-up (full step: render + loss + backward + Adam), torch CPU fp32, "
                             + ("the unmodified reference files staged in oracle/_ref" if kind == "reference" else "oracle port")}
            if kind == "reference":
                # the same unmodified reference code on THIS B200 through torch's eager CUDA kernels (strict fp32, allow_tf32 off):
                # the same-box comparator of SURVEY 8d.  A reported baseline like cpu_baseline, not the target.
                try:
                    v2, ms2, _, _ = cpu_reference_rate(args, RAYS_PER_GPU, 5, 2, device="cuda")
                    cuda_eager = {"value": v2, "unit": "rays/s", "ms_per_step": ms2, "kind": "reference", "device": "cuda:0 (torch eager, fp32, allow_tf32=False)",
                                  "sample": "5 steps of 1024 rays after 2 warm-up, the unmodified reference files on the B200"}
                except Exception as e:                                   # noqa: BLE001
                    cuda_eager = {"error": f"{type(e).__name__}: {e}"}
        line = {"metric": METRIC, "value": total_rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": opts.steps,
                "warmup": max(3, opts.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if opts.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "rays_per_gpu": RAYS_PER_GPU, "global_rays": total_rays, "parallelism": f"ray-sharded dp{world}",
                           "cuda_graph": bool(opts.graph), "e2e_feed": opts.feed,
                           "l2": "per-step working set (~3.4 GB of activations) >> 126 MB L2; no explicit flush",
                           "legs": f"roofline (per-launch events, 3 eager steps), resident (value), e2e: each after {opts.cooldown:g} s of idle "
                                   "GPU and its own warm-up steps, so all three start from the same power state (a 1 kW part: ~50 ms of "
                                   "burst clocks, then sw_power_cap); `sustained` is the steady state of the same step"},
                "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(host_batch.flat.numel()), "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "roofline_composite": roof_hbm,
                "sustained": None if ms_sus is None else {
                    "value": total_rays / (ms_sus * 1e-3), "unit": "rays/s", "ms_per_step": ms_sus, "steps": n_sus,
                    "seconds": ms_sus * n_sus * 1e-3, "clocks": clk_sus,
                    "note": "same resident step back to back for >= --sustain seconds (power-capped steady state); `value` above is the K-step burst"},
                "cpu_baseline": cpu, "reference_cuda_eager": cuda_eager, "other_configs": other,
                "tile_products": tile_products_leg(cpu=not opts.no_cpu_baseline) if (world == 1 and not opts.no_tile_products) else None,
                "loss": float(loss_host)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--feed", default="inline", choices=["prefetch", "inline"],
                    help="e2e leg: host batches through Trainer.prefetch() on a copy stream, or copied on the compute stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs[2..4] leg")
    ap.add_argument("--cooldown", type=float, default=1.0,
                    help="seconds of idle GPU before each timed leg (roofline, resident, e2e), so that every leg starts from the same "
                         "power state; its own warm-up steps follow.  The sustained leg reports the power-capped steady state")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure (0 = off)")
    ap.add_argument("--no-composite", action="store_true", help="skip the compositing (HBM) roofline leg")
    ap.add_argument("--no-tile-products", action="store_true", help="skip the ray-feed / DSM leg (SURVEY 8f-3/4 kernels)")
    opts = ap.parse_args()
    if opts.impl == "reference":
        run_reference(opts)
    else:
        run_ours(opts)


if __name__ == "__main__":
    main()
