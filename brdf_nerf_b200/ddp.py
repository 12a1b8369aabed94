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
    """In-place sum of `bucket` (n fp32) over the ranks of `group` through NVLink peer memory.

    Construction is collective and cannot deadlock on a local failure: every rank always takes part in the same two
    collectives (the all_gather of handle + status byte, the all_reduce of the final status); if any rank failed to allocate,
    export or map, `ok` is False on EVERY rank and the caller falls back to torch.distributed."""

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
        self._base, self._peer_base, self.error = None, [], None
        payload = torch.zeros(65, dtype=torch.uint8, device=device)          # 64-byte IPC handle + status (1 = ok)
        with torch.cuda.device(device):
            try:
                base = C.c_void_p()
                L.check(lib.bn_peer_alloc(nbytes, C.byref(base)))
                self._base = base
                handle = (C.c_ubyte * 64)()
                L.check(lib.bn_peer_export(base, handle))
                payload = torch.tensor(list(bytes(handle)) + [1], dtype=torch.uint8, device=device)
            except Exception as e:                                           # noqa: BLE001
                self.error = f"{type(e).__name__}: {e}"
            all_h = [torch.empty_like(payload) for _ in range(self.world)]
            dist.all_gather(all_h, payload, group=group)
            all_h = [h.cpu().tolist() for h in all_h]
            ok = all(h[64] == 1 for h in all_h)
            if ok:
                try:
                    for p in range(self.world):
                        if p == self.rank:
                            self._peer_base.append(self._base.value)
                            continue
                        hb = (C.c_ubyte * 64)(*all_h[p][:64])
                        ptr = C.c_void_p()
                        L.check(lib.bn_peer_open(hb, C.byref(ptr)))
                        self._peer_base.append(ptr.value)
                except Exception as e:                                       # noqa: BLE001
                    self.error = f"{type(e).__name__}: {e}"
                    ok = False
            flag = torch.tensor([1.0 if ok else 0.0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)         # also: every rank has opened every handle
            self.ok = bool(flag.item() >= 1.0)
        if not self.ok:
            self.close()
            return
        self._bufs = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in self._peer_base])
        self._flags = (C.c_void_p * self.world)(*[C.c_void_p(b + self.flag_off) for b in self._peer_base])
        self.bucket = torch.as_tensor(_RawCuda(self._base.value, self.n), device=device)      # this rank's bucket as a tensor
        self._epoch = torch.zeros(128, dtype=torch.int32, device=device)

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
        self.ok = False


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
    ex = PeerExchange(flat.numel(), flat.device, group)
    if not ex.ok:
        if ex.error:
            import sys
            print(f"[brdf_nerf_b200.ddp] peer mapping unavailable on rank {ex.rank} ({ex.error}); using torch.distributed.all_reduce",
                  file=sys.stderr)
        return None
    model.rebind_grads(ex.bucket)
    return ex
