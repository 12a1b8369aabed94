"""Data-parallel gradient exchange: ray-sharded training keeps the weights replicated and sums the flat fp32
gradient bucket over the ranks once per step — what DDP does for the reference under Lightning (main.py:720-731).

Two implementations behind one call:
  * `PeerExchange` (NCCL process groups on one NVLink node): the bucket of every rank lives in peer-mapped device memory
    (CUDA IPC) and ONE kernel of this library (`bn_allreduce_p2p`, csrc/ddp.cu) reads the peers' buckets and writes the sums
    back over NVLink.  No host-side state: it is captured into the step's CUDA graph together with the fused Adam, so a
    multi-GPU step is one graph launch per rank.
  * `torch.distributed.all_reduce` on the bucket (gloo in the CPU tests; NCCL when peer mapping is unavailable, e.g. ranks
    on different nodes).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


class _RawCuda:
    """__cuda_array_interface__ view of a device allocation owned by the library."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerExchange:
    """In-place sum of `bucket` (n fp32) over the ranks of `group` through NVLink peer memory."""

    N_BLOCKS = 128

    def __init__(self, n_floats: int, device: torch.device, group=None):
        import torch.distributed as dist
        lib = L.load()
        self.lib, self.group = lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if n_floats % 4:
            raise ValueError("bucket length must be a multiple of 4 floats")
        self.n = int(n_floats)
        flag_words = int(lib.bn_allreduce_p2p_flag_words())
        self.flag_off = (self.n * 4 + 255) // 256 * 256
        nbytes = self.flag_off + flag_words * 4
        with torch.cuda.device(device):
            base = C.c_void_p()
            L.check(lib.bn_peer_alloc(nbytes, C.byref(base)))
            self._base = base
            handle = (C.c_ubyte * 64)()
            L.check(lib.bn_peer_export(base, handle))
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
            all_h = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(all_h, mine, group=group)
            self._peer_base = []
            for p in range(self.world):
                if p == self.rank:
                    self._peer_base.append(base.value)
                    continue
                hb = (C.c_ubyte * 64)(*all_h[p].cpu().tolist())
                ptr = C.c_void_p()
                L.check(lib.bn_peer_open(hb, C.byref(ptr)))
                self._peer_base.append(ptr.value)
        self._bufs = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in self._peer_base])
        self._flags = (C.c_void_p * self.world)(*[C.c_void_p(b + self.flag_off) for b in self._peer_base])
        self.bucket = torch.as_tensor(_RawCuda(base.value, self.n), device=device)      # this rank's bucket as a tensor
        self._epoch = torch.zeros(128, dtype=torch.int32, device=device)
        dist.barrier(group=group)                  # every rank has opened every handle before the first exchange

    def all_reduce_(self):
        """Enqueue the exchange on the current stream (graph-capturable)."""
        L.check(self.lib.bn_allreduce_p2p(self._bufs, self._flags, C.c_void_p(self._epoch.data_ptr()), self.n, self.rank,
                                          self.world, self.N_BLOCKS, L.stream_ptr()))

    def close(self):
        for p, b in enumerate(self._peer_base):
            if p != self.rank and b:
                self.lib.bn_peer_close(C.c_void_p(b))
        self._peer_base = []
        if self._base is not None:
            self.lib.bn_peer_free(self._base)
            self._base = None


def make_exchange(model, group=None) -> Optional[PeerExchange]:
    """PeerExchange for the model's gradient bucket when the ranks can map each other's memory (NCCL backend, CUDA
    devices of one node); None otherwise (the caller all-reduces with torch.distributed).  On success the model's
    parameter gradients are re-bound to the peer-mapped bucket."""
    import os
    import torch.distributed as dist
    if os.environ.get("BN_NO_P2P") or not dist.is_initialized() or dist.get_backend(group) != "nccl":
        return None
    if dist.get_world_size(group) > 8:                      # the kernel addresses one NVLink / NVSwitch node (<= 8 peers)
        return None
    flat = model.flat_params
    if not flat.is_cuda:
        return None
    ok = torch.ones(1, device=flat.device)
    ex = None
    try:
        ex = PeerExchange(flat.numel(), flat.device, group)
    except Exception as e:                                  # noqa: BLE001  (peer mapping refused: fall back on every rank)
        print(f"[brdf_nerf_b200.ddp] peer mapping unavailable ({type(e).__name__}: {e}); using torch.distributed.all_reduce")
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if ok.item() < 1:
        if ex is not None:
            ex.close()
        return None
    model.rebind_grads(ex.bucket)
    return ex
