"""Minimal training-step driver for the hot path (replaces the Lightning glue of the reference,
main.py:194-353 + 147-168): forward through the CUDA pipeline, loss gradients,
backward kernels writing straight into the flat fp32 gradient bucket, ONE all-reduce of that bucket
over NCCL when world_size > 1, fused Adam on the flat buffers.

No autograd tape is built: the step calls the forward / backward kernel chains directly
(`rendering._forward` / `rendering._backward`).  Nothing in a step synchronises with the host, so it
can be captured in a CUDA graph (`use_graph=True`).

Losses (reference metrics.py:39-61 SNerfLoss with lambda_sc = 0, metrics.py:82-161 DepthLoss with
subset=True, GNLL=False):
    L = lambda_rgb * mean((rgb - target)^2)
      + [step < ds_drop] (lambda_ds / 3) * mean_sel( n_sel/N * w_i (d_i - d*_i)^2 )
The depth term's normaliser n_sel/N * 1/n_sel collapses to 1/N, so the selection mask is applied as
a 0/1 weight without any host round trip (the reference syncs three times to build index lists).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import ops
from . import rendering as R
from .synth import RayBatch


def loss_and_grads(args, outs, st, batch: RayBatch, use_depth: bool):
    """Returns (loss (1,) tensor, g_rgb (N,3), g_depth (N) or None): one fused launch (csrc/loss.cu)."""
    use_ds = use_depth and batch.valid_depth is not None
    return ops.loss_color_depth(outs["rgb"], batch.rgbs, float(args.lambda_rgb), depth=outs["depth"], z=outs["z"],
                                weights=outs["weights"], valid_depth=batch.valid_depth if use_ds else None,
                                target_depths=batch.target_depths, target_std=batch.target_std,
                                lambda_ds=float(args.ds_lambda) if use_ds else 0.0,
                                use_all_depth=bool(getattr(args, "usealldepth", False)),
                                no_weights=bool(getattr(args, "ds_noweights", False)))


def allreduce_grads_(flat_grads: torch.Tensor, world_size: int, group=None) -> float:
    """Data-parallel gradient exchange: ONE all-reduce (sum) of the flat fp32 bucket over
    NCCL/NVLink (gloo in the CPU tests).  Returns the scale (1/world) the optimizer must apply — the
    averaging is folded into the fused Adam kernel instead of a separate pass over the bucket.
    Matches what DDP does in the reference (Lightning `gpus > 1`, main.py:720-731): per-rank mean loss,
    gradients averaged over ranks."""
    if world_size > 1:
        torch.distributed.all_reduce(flat_grads, op=torch.distributed.ReduceOp.SUM, group=group)
    return 1.0 / world_size


class Trainer:
    def __init__(self, model, args, lr: Optional[float] = None, world_size: int = 1, process_group=None,
                 use_graph: bool = False):
        self.model, self.args = model, args
        self.lr = float(args.lr if lr is None else lr)
        self.world = int(world_size)
        self.pg = process_group
        self.step_count = 0
        self.use_graph = use_graph
        flat = model.flat_params
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self._graph = None
        self._static: Optional[RayBatch] = None
        self._loss = None
        self._kw = None
        self.graph_launches = 0
        self.use_depth_loss = True          # switched off by the schedule after ds_drop (main.py:248)
        self.use_normal_reg = True          # NormalRegLoss once train_steps > nrrg_on (main.py:272,280); needs nr_reg_*_lambda > 0
        self.use_hard_surface = False       # HardSurfaceLoss once epoch > 2 (main.py:292); needs hs_lambda > 0
        # the captured graph includes Adam (one host launch per step).  A torch.distributed (NCCL) all-reduce stays outside
        # the graph: captured, it ran slower on 2 B200s (2.80 vs 2.53 ms/step) and stalled process-group teardown.
        # With world_size > 1 on one NVLink node the bucket lives in peer-mapped memory and the exchange is a kernel of this
        # library (brdf_nerf_b200.ddp.PeerExchange, csrc/ddp.cu): capturable, so the graph again holds the whole step.
        self._exchange = None
        if self.world > 1:
            from . import ddp
            self._exchange = ddp.make_exchange(model, process_group)
        self.whole_step_graph = self.world == 1 or self._exchange is not None
        self._opt_state = None              # device [lr, step, bc1, sqrt(bc2)] of the graph-captured Adam
        self._dev_lr, self._dev_step, self._graph_updates = None, 0, False
        # host feed (prefetch / read_loss_async): a copy stream next to the compute stream, two staging batches
        self._copy_stream = None
        self._stage, self._stage_ready, self._stage_free, self._stage_i = [None, None], [None, None], [None, None], 0
        self._loss_ring, self._loss_i = None, 0
        # NormalLoss (nr_spv_lambda / nr_spv_type, main.py:301-327: supervision by normals derived from the depth maps) is
        # not part of this driver: refuse instead of silently training without it
        if abs(float(getattr(args, "nr_spv_lambda", 0.0) or 0.0)) > 1e-5:
            raise NotImplementedError("nr_spv_lambda != 0 (NormalLoss, main.py:301-327) is not implemented by Trainer")
        if getattr(model, "beta", False) == True:           # noqa: E712
            raise NotImplementedError("Trainer drives the spsbrdf-nerf recipe, which never trains with --beta (no loss reads the "
                                      "channel, metrics.py:170-171); render beta models through render_rays + autograd")
        self._frozen = None                 # [(offset, length)] of the parameters with requires_grad == False

    # one optimisation step; `batch` tensors must already live on the model's device
    def _step_impl(self, batch: RayBatch, draws, kw):
        model, args = self.model, self.args
        # nothing below reads the depth-ordered per-sample rows unless a regulariser is on (normals: decided inside _forward)
        hs_on = self.use_hard_surface and float(getattr(args, "hs_lambda", 0.0)) != 0.0
        outs, st = R._forward(model, args, batch.rays, draws, train=True, mode="train",
                              valid_depth=batch.valid_depth, target_depths=batch.target_depths,
                              target_std=batch.target_std, lazy_packed=not hs_on, defer_kc=not hs_on, **kw)
        if outs.get("deferred"):          # plain Lambertian stage: compositing, colour, loss and their backward in one launch
            grads = model.flat_grads
            grads.zero_()
            loss = R._lambertian_loss_backward(model, st, args, batch.rgbs, batch.valid_depth, batch.target_depths,
                                               batch.target_std, float(args.ds_lambda) > 0 and self.use_depth_loss, grads)
            self._mask_frozen(grads)
            return loss
        use_depth = float(args.ds_lambda) > 0 and self.use_depth_loss
        loss, g_rgb, g_depth = loss_and_grads(args, outs, st, batch, use_depth)
        g_weights = g_packed = None
        # optional regularisers (main.py:269-299): normals facing away from the camera, hard surface
        lam_an = float(getattr(args, "nr_reg_an_lambda", 0.0)) if (outs["nr_an"] and self.use_normal_reg) else 0.0
        lam_lr = float(getattr(args, "nr_reg_lr_lambda", 0.0)) if (outs["nr_lr"] and self.use_normal_reg) else 0.0
        lam_hs = float(getattr(args, "hs_lambda", 0.0)) if self.use_hard_surface else 0.0
        if lam_an != 0.0 or lam_lr != 0.0 or lam_hs != 0.0:
            ch_an = 4 if outs["nr_an"] else -1
            ch_lr = (7 if outs["nr_an"] else 4) if outs["nr_lr"] else -1
            g_weights, g_packed, g_depth, _ = ops.loss_regularizers(
                loss, outs["weights"], outs["z"], outs["depth"], outs["packed"], st.rays, ch_an, lam_an, ch_lr, lam_lr,
                lam_hs, g_depth=g_depth)
        grads = model.flat_grads
        grads.zero_()
        R._backward(model, st, g_rgb, g_depth, g_weights, g_packed, grads)
        self._mask_frozen(grads)
        return loss

    def frozen_ranges(self):
        """Merged [offset, offset+length) runs of the flat buffer whose parameters have requires_grad == False
        (model.freeze / freeze_rest).  The reference hands only `filter(requires_grad)` to Adam (main.py:148-150)."""
        runs = []
        for _, p, off, n in self._param_slices():
            if not p.requires_grad:
                if runs and runs[-1][0] + runs[-1][1] >= off:
                    runs[-1][1] = off + n - runs[-1][0]
                else:
                    runs.append([off, n])
        return [tuple(r) for r in runs]

    def _mask_frozen(self, grads):
        """Frozen parameters take no optimizer step: their gradient and Adam moments are zeroed before the fused Adam, whose
        update m / (sqrt(v) + eps) is then exactly 0 for them.  The set is read once per graph capture / eager step."""
        runs = self.frozen_ranges()
        for off, n in runs:
            grads[off:off + n].zero_()
            self.m[off:off + n].zero_()
            self.v[off:off + n].zero_()
        self._frozen = runs

    def _exchange_grads(self) -> float:
        """Sum the gradient bucket over the ranks; returns the scale (1 / world) the optimizer applies."""
        if self._exchange is not None:
            self._exchange.all_reduce_()
            return 1.0 / self.world
        return allreduce_grads_(self.model.flat_grads, self.world, self.pg)

    def _reduce_and_update(self):
        model = self.model
        scale = self._exchange_grads()
        self.step_count += 1
        ops.adam_step(model.flat_params, model.flat_grads, self.m, self.v, self.lr, self.step_count,
                      grad_scale=scale)
        model._synced_version = -1            # the packed bf16 copies are stale now

    def step(self, batch: RayBatch, draws=None, apply_brdf=False, apply_theta=False, cos_irra_on=False,
             gsam_only=False):
        """One optimisation step.  With `use_graph` the batch may live in (pinned) host memory: its tensors are copied
        straight into the graph's static input buffers (one async H2D copy per tensor); passing `static_batch()` itself
        skips the copies."""
        kw = dict(apply_brdf=apply_brdf, bTestNormal=False, bTestSun_v=False, gsam_only=gsam_only,
                  apply_theta=apply_theta, cos_irra_on=cos_irra_on)
        if not self.use_graph or draws is not None:
            ready = getattr(batch, "_ready", None)
            if ready is not None:                         # prefetch()ed staging batch, consumed in place by the eager step
                torch.cuda.current_stream().wait_event(ready)
            loss = self._step_impl(batch, draws, kw)
            self._reduce_and_update()
            if ready is not None:
                batch._free.record()
            return loss
        return self._graph_step(batch, kw)

    # ---- checkpoint / resume (reference: Lightning ModelCheckpoint + resume_from_checkpoint, main.py:709-723).  The layout
    # is the Lightning checkpoint's: "state_dict" with the module under its LightningModule attribute name
    # (`nerf_coarse.<param>`, the prefix eval.py:26-54 / main.py:97-104 cut), "optimizer_states" = [torch.optim.Adam
    # state_dict] over the parameters in registration order (main.py:147-150), "global_step".  A checkpoint written by the
    # reference therefore resumes here and vice versa.
    CKPT_PREFIX = "nerf_coarse"

    def _param_slices(self):
        flat = self.model.flat_params
        out = []
        for name, p in self.model.named_parameters():
            off = (p.data_ptr() - flat.data_ptr()) // flat.element_size()
            out.append((name, p, int(off), p.numel()))
        return out

    def state_dict(self) -> dict:
        slices = self._param_slices()
        state = {}
        if self.step_count > 0:                            # torch.optim.Adam creates its state lazily, at the first step
            for i, (_, p, off, n) in enumerate(slices):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[off:off + n].view_as(p).detach().clone(),
                            "exp_avg_sq": self.v[off:off + n].view_as(p).detach().clone()}
        group = {"lr": self.lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                 "params": list(range(len(slices)))}
        return {"state_dict": {f"{self.CKPT_PREFIX}.{k}": v.detach().clone() for k, v in self.model.state_dict().items()},
                "optimizer_states": [{"state": state, "param_groups": [group]}], "global_step": self.step_count}

    def load_state_dict(self, ckpt: dict, load_model: bool = True, strict: bool = True):
        """Resume from `state_dict()` or from a Lightning checkpoint of the reference (`torch.load(path)`)."""
        if load_model:
            prefix = self.CKPT_PREFIX + "."
            sd = {k[len(prefix):]: v for k, v in ckpt["state_dict"].items() if k.startswith(prefix)}
            self.model.load_state_dict(sd, strict=strict)   # copies in place: parameters stay views of the flat buffer
        opt = ckpt["optimizer_states"][0]
        slices = self._param_slices()
        ids = opt["param_groups"][0]["params"]
        if len(ids) != len(slices):
            raise ValueError(f"optimizer state covers {len(ids)} parameters, the model has {len(slices)}")
        self.m.zero_()
        self.v.zero_()
        step = 0
        for pid, (name, p, off, n) in zip(ids, slices):
            st = opt["state"].get(pid)
            if st is None:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state of {name}: shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
            self.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        self.lr = float(opt["param_groups"][0]["lr"])
        self.step_count = step
        self.model._synced_version = -1                    # packed bf16 copies are stale
        # the graph-captured Adam keeps lr / step on the device: step() refreshes them when they differ from the host's
        self._dev_lr, self._dev_step = None, -1

    # ---- host feed: the H2D copy of step k+1's batch and the D2H read of step k's loss run on a copy stream, so the compute
    # stream sees kernels only (no copy-engine hop between two graph replays).  Event order, per staging buffer b:
    #   copy stream:    wait free[b] -> H2D host batch -> record ready[b]
    #   compute stream: wait ready[b] -> static <- staging[b] (device copy) -> record free[b] -> graph replay
    def prefetch(self, host_batch: RayBatch) -> RayBatch:
        """Starts the host->device copy of a packed (ideally pinned) host batch on the trainer's copy stream and returns
        the device-side staging batch to hand to `step()`; call it for batch k+1 right after enqueuing step k.  At most
        two prefetched batches may be outstanding (two staging buffers)."""
        if host_batch.flat is None:
            raise ValueError("prefetch() needs a packed host batch: RayBatch.packed(pin=True)")
        dev = self.model.flat_params.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        b = self._stage_i
        self._stage_i ^= 1
        st = self._stage[b]
        if st is None or st.flat.numel() != host_batch.flat.numel():
            st = self._stage[b] = host_batch.packed(device=dev)
            self._stage_ready[b], self._stage_free[b] = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free[b].record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[b])
            st.flat.copy_(host_batch.flat, non_blocking=True)
            self._stage_ready[b].record()
        st._ready, st._free = self._stage_ready[b], self._stage_free[b]
        return st

    def read_loss_async(self, loss: torch.Tensor, pinned_out: torch.Tensor) -> "torch.cuda.Event":
        """Copies a step's loss into pinned host memory on the copy stream; returns the event that marks the value valid.
        The value is first parked in a 4-slot device ring (the graph's loss tensor is rewritten by the next replay), so
        the caller may run up to three steps ahead of the read."""
        dev = self.model.flat_params.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if self._loss_ring is None:
            self._loss_ring = torch.zeros(4, 1, dtype=torch.float32, device=dev)
        slot = self._loss_ring[self._loss_i % 4]
        self._loss_i += 1
        torch.add(loss.reshape(1), 0.0, out=slot)         # an elementwise kernel on the compute stream, not a memcpy
        parked = torch.cuda.Event()
        parked.record()
        done = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(parked)
            pinned_out.copy_(slot, non_blocking=True)
            done.record()
        return done

    def static_batch(self) -> Optional[RayBatch]:
        """The captured graph's input buffers (None before the first graph step): producers may write the next batch
        into them directly and pass this object to `step()`."""
        return self._static

    # ---- CUDA-graph path: static input buffers; forward + losses + backward + gradient all-reduce + Adam are captured
    # once and replayed per step (ONE host launch per step; lr and the Adam step counter live in device memory)
    def _capture(self, rkw, whole_step: bool):
        model = self.model
        g = torch.cuda.CUDAGraph()
        lib = L.load()
        lc0 = lib.bn_launch_count()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            model.sync_weights(force=True)
            self._loss = self._step_impl(self._static, None, rkw)
            if whole_step:
                scale = self._exchange_grads()
                ops.adam_step_graph(model.flat_params, model.flat_grads, self.m, self.v, self._opt_state, grad_scale=scale)
        self.graph_launches = int(lib.bn_launch_count() - lc0)     # library kernels replayed by every graph launch
        return g

    def _graph_step(self, batch: RayBatch, kw):
        kw = dict(kw, _use_depth=self.use_depth_loss, _use_nr=self.use_normal_reg, _use_hs=self.use_hard_surface,
                  _frozen=tuple(self.frozen_ranges()))        # the frozen set is baked into the captured graph
        ready = getattr(batch, "_ready", None)
        if ready is not None:                             # a prefetch()ed staging batch: its H2D copy runs on the copy stream
            torch.cuda.current_stream().wait_event(ready)
        if self._graph is None or self._kw != kw:
            dev = self.model.flat_params.device
            self._static = batch.packed(device=dev)       # one buffer: a packed host batch arrives with a single copy
            rkw = {k: v for k, v in kw.items() if not k.startswith("_")}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up: fills table / workspace caches
                for _ in range(2):
                    self.model.sync_weights(force=True)
                    self._step_impl(self._static, None, rkw)
            torch.cuda.current_stream().wait_stream(side)
            if self._opt_state is None:
                self._opt_state = torch.zeros(4, dtype=torch.float32, device=self.model.flat_params.device)
            self._graph_updates = self.whole_step_graph
            self._graph = self._capture(rkw, self._graph_updates)
            self._kw = dict(kw)
        if batch is self._static:
            pass
        elif batch.flat is not None and batch.flat.numel() == self._static.flat.numel():
            self._static.flat.copy_(batch.flat, non_blocking=True)
        else:
            for dst, src in zip((self._static.rays, self._static.rgbs, self._static.valid_depth,
                                 self._static.target_depths, self._static.target_std),
                                (batch.rays, batch.rgbs, batch.valid_depth, batch.target_depths, batch.target_std)):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
        if ready is not None:
            batch._free.record()                          # the staging buffer may be refilled from here on
        if not self._graph_updates:
            self._graph.replay()
            self._reduce_and_update()
            return self._loss
        if self._dev_lr != self.lr:                       # StepLR: changes once per epoch
            self._opt_state[0:1].fill_(self.lr)
            self._dev_lr = self.lr
        if self._dev_step != self.step_count:             # eager steps were taken in between
            self._opt_state[1:2].fill_(float(self.step_count))
        self._graph.replay()
        self.step_count += 1
        self._dev_step = self.step_count
        self.model._synced_version = -1
        return self._loss


class TrainLoop:
    """The reference's training loop around the hot path (NeRF_pl.training_step + DataLoader + StepLR,
    main.py:147-246) without Lightning: device-resident ray pool -> schedule flags -> Trainer.step.
    One call of `step()` = one optimisation step; nothing synchronises with the host."""

    def __init__(self, model, args, pool: RayBatch, world_size: int = 1, rank: int = 0, process_group=None,
                 use_graph: bool = True, seed: int = 0):
        from .schedule import DeviceRayPool, Schedule
        self.args = args
        self.trainer = Trainer(model, args, world_size=world_size, process_group=process_group, use_graph=use_graph)
        self.feed = DeviceRayPool(pool, int(args.batch_size), rank=rank, world_size=world_size, seed=seed)
        self.schedule = Schedule(args, dataset_len=pool.rays.shape[0], world_size=world_size)
        if float(args.noise_std) != 0.0 and use_graph:
            raise ValueError("noise_std decays every step (main.py:246): run with use_graph=False when it is non-zero")

    def state_dict(self) -> dict:
        """Trainer checkpoint + the loop's own position: schedule step, noise level, ray-pool epoch / cursor / generator."""
        sd = self.trainer.state_dict()
        feed = self.feed
        sd["loop"] = {"train_steps": self.schedule.train_steps, "noise_std": self.schedule.noise_std, "epoch": feed.epoch,
                      "pos": feed._pos, "perm": None if feed._perm is None else feed._perm.clone(),
                      "generator": feed.gen.get_state()}
        sd["epoch"] = self.schedule.epoch()
        return sd

    def load_state_dict(self, ckpt: dict, load_model: bool = True):
        self.trainer.load_state_dict(ckpt, load_model=load_model)
        loop = ckpt.get("loop")
        if loop is not None:
            self.schedule.train_steps, self.schedule.noise_std = int(loop["train_steps"]), float(loop["noise_std"])
            feed = self.feed
            feed.epoch, feed._pos = int(loop["epoch"]), int(loop["pos"])
            feed._perm = None if loop["perm"] is None else loop["perm"].to(feed.pool.rays.device)
            feed.gen.set_state(loop["generator"])
        else:                                              # a reference checkpoint: only the global step is known
            self.schedule.train_steps = int(ckpt.get("global_step", self.trainer.step_count)) * self.schedule.world

    def step(self):
        f = self.schedule.next()
        self.trainer.lr = f.lr
        self.trainer.use_depth_loss = f.use_depth_loss
        self.trainer.use_normal_reg = f.use_normal_reg
        self.trainer.use_hard_surface = self.schedule.epoch() > 2
        self.args.noise_std = f.noise_std
        batch = self.feed.next_batch()
        return self.trainer.step(batch, apply_brdf=f.apply_brdf, apply_theta=f.apply_theta, cos_irra_on=f.cos_irra_on,
                                 gsam_only=f.gsam_only)
