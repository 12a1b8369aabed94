"""SpSBRDFNeRF — same constructor, attributes and `state_dict` as the reference module
(reference models/spsbrdfnerf.py:418-757), but every evaluation goes through the C-ABI CUDA library.

What is kept bit-for-bit from the reference surface
  * parameter names / shapes (`fc_net.{0,2,..}.{weight,bias}`, `sigma_from_xyz.0.*`,
    `feats_from_xyz.*`, `rgb_from_xyzdir.{0,2}.*`, `{k,theta_rpv,rhoc,b,c,theta,roughness}_from_xyz.{0,2}.*`,
    `grad_from_xyz.*`) and the construction / initialisation order, so `torch.manual_seed(s);
    load_model(args)` yields the same weights as the reference and checkpoints interchange;
  * attributes read by `inference`: number_of_outputs[_brdf], normal, sun_v, indirect_light, beta,
    roughness, RPV, MultiBRDF, rgb_padding, glossy_scale, args;
  * freeze / unfreeze / freeze_rest / check_nan_parms / print_parms.

What is different underneath: all parameters are views into ONE flat fp32 buffer (`flat_params`),
their gradients views into `flat_grads` — the buffer the wgrad kernels accumulate into and the
single NCCL all-reduce bucket of data-parallel training.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib as L


class Sine(nn.Module):
    """sin(w0 x) (reference models/nerf.py:23-33); parameterless, only here so that the Sequential
    indices (and with them the state_dict keys) match the reference."""

    def __init__(self, w0: float = 1.0):
        super().__init__()
        self.w0 = w0

    def forward(self, x):  # pragma: no cover - the CUDA path never calls it
        raise RuntimeError("Sine is a structural placeholder; evaluation runs in libbrdfnerf_b200")


def _siren_uniform(linear: nn.Linear, first: bool):
    fan_in = linear.weight.size(-1)
    bound = 1.0 / fan_in if first else math.sqrt(6.0 / fan_in)
    with torch.no_grad():
        linear.weight.uniform_(-bound, bound)


def _head(feat: int, out: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(feat, feat // 2), Sine(), nn.Linear(feat // 2, out), nn.Sigmoid())


class SpSBRDFNeRF(nn.Module):
    def __init__(self, args, layers=8, feat=256, mapping=False, mapping_sizes=[10, 4], skips=[4], siren=True,
                 t_embedding_dims=16, beta=True, roughness=True, normal="none", sun_v="none",
                 indirect_light=False, glossy_scale=1., MultiBRDF=False, dim_RPV=3, precision="fp32"):
        super().__init__()
        if not siren:
            raise NotImplementedError("siren=0 (ReLU trunk) is outside the CUDA hot path")
        # R19 of SURVEY §8a: `beta` is built; two optional output channels are refused
        if sun_v == "learned":
            raise NotImplementedError("sun_v='learned' raises NameError in the reference itself (spsbrdfnerf.py:697 reads the "
                                      "undefined `xyz_features_`, SURVEY App. C.2): there is no behaviour to mirror")
        if indirect_light:
            raise NotImplementedError("indirect_light=True (sky_color head, spsbrdfnerf.py:562-568,704-706) is not built: the "
                                      "reference reads it from the hard-coded channels out[..., 5:8] (spsbrdfnerf.py:154), which "
                                      "is only the sky colour when sun_v='learned' (itself broken, App. C.2); with "
                                      "sun_v='analystic' those channels are sky g, sky b and the next head's first output")
        if len(skips) > 1:
            raise NotImplementedError("one skip connection is supported")
        self.layers, self.skips, self.feat = layers, list(skips), feat
        self.t_embedding_dims = t_embedding_dims
        self.viewdir = bool(getattr(args, "input_viewdir", 0))
        self.input_sizes = [3, 3] if self.viewdir else [3, 0]          # spsbrdfnerf.py:458
        self.rgb_padding = 0.001
        self.beta, self.roughness, self.sun_v = beta, roughness, sun_v
        self.indirect_light, self.normal = indirect_light, normal
        self.glossy_scale, self.MultiBRDF, self.args = glossy_scale, MultiBRDF, args
        self.RPV = bool(args.funcM == True or args.funcF == True or args.funcH == True)   # noqa: E712 (as reference)
        self.dim_RPV = dim_RPV
        self.mapping_sizes = list(mapping_sizes)
        self.n_freq = mapping_sizes[0] if mapping else 0
        self.n_freq_dir = mapping_sizes[1] if mapping else 0
        self.precision = precision

        self.number_of_outputs = 4 + int(beta == True)            # noqa: E712  (+ beta, spsbrdfnerf.py:476-477)
        self.number_of_outputs_brdf = self.number_of_outputs
        if self.roughness == True:                                # noqa: E712
            self.number_of_outputs_brdf += 1
        elif self.RPV:
            self.number_of_outputs_brdf += 3 * sum(int(f == True) for f in (args.funcM, args.funcF, args.funcH))  # noqa: E712
        else:
            self.number_of_outputs_brdf += 3 * (int(args.b == True) + int(args.c == True))   # noqa: E712

        in0 = 2 * mapping_sizes[0] * 3 if mapping else 3
        # --- same creation order as the reference so that a seeded init reproduces its weights ---
        fc = [nn.Linear(in0, feat), Sine(30.0)]
        for i in range(1, layers):
            fc += [nn.Linear(feat + in0 if i in skips else feat, feat), Sine()]
        self.fc_net = nn.Sequential(*fc)
        self.sigma_from_xyz = nn.Sequential(nn.Linear(feat, 1), nn.Softplus())
        self.feats_from_xyz = nn.Linear(feat, feat)
        dir_in = (2 * mapping_sizes[1] * 3 if mapping else 3) if self.viewdir else 0      # spsbrdfnerf.py:505-509, 534
        self.rgb_from_xyzdir = nn.Sequential(nn.Linear(feat + dir_in, feat // 2), Sine(), nn.Linear(feat // 2, 3), nn.Sigmoid())
        for i in range(layers):                     # fc_net.apply(sine_init)
            _siren_uniform(self.fc_net[2 * i], first=False)
        _siren_uniform(self.fc_net[0], first=True)  # fc_net[0].apply(first_layer_sine_init)
        if beta == True:                                          # noqa: E712  spsbrdfnerf.py:571-575: [features | t] -> beta
            self.beta_from_xyz = nn.Sequential(nn.Linear(t_embedding_dims + feat, feat // 2), Sine(),
                                               nn.Linear(feat // 2, 1), nn.Softplus())
        if normal in ("analystic_learned", "learned"):
            self.grad_from_xyz = nn.Linear(feat, 3)
        if self.roughness == True:                                # noqa: E712
            self.roughness_from_xyz = _head(feat, 1)
        if args.funcM == True:                                    # noqa: E712
            self.k_from_xyz = _head(feat, dim_RPV)
        if args.funcF == True:                                    # noqa: E712
            self.theta_rpv_from_xyz = _head(feat, dim_RPV)
        if args.funcH == True:                                    # noqa: E712  (funcH == 2 creates no head)
            self.rhoc_from_xyz = _head(feat, dim_RPV)
        if args.b == True:                                        # noqa: E712
            self.b_from_xyz = _head(feat, 1)
        if args.c == True:                                        # noqa: E712
            self.c_from_xyz = _head(feat, 1)
        if args.theta == True:                                    # noqa: E712
            self.theta_from_xyz = _head(feat, 1)

        self._flat: Optional[torch.Tensor] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._handle = None
        self._synced_version = -1
        self._dirty = 0
        self._ws_cache: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ reference utility surface
    def freeze(self, layer_name):
        for name, p in self.named_parameters():
            if layer_name in name or layer_name == "all":
                p.requires_grad = False

    def unfreeze(self, layer_name):
        for name, p in self.named_parameters():
            if layer_name in name:
                p.requires_grad = True

    def freeze_rest(self, layer_name):
        for name, p in self.named_parameters():
            if layer_name not in name:
                p.requires_grad = False

    def check_nan_parms(self, keyword=""):
        """Reference: one host sync per parameter tensor; here one reduction over the flat buffer."""
        self._ensure_flat()
        bad = bool(torch.isnan(self._flat).any())
        if bad:
            print(f"{keyword} check_nan_parms: NaN in parameters")
        return bad

    def print_parms(self, only_name=False):
        total = 0
        for name, p in self.named_parameters():
            total += p.numel()
            if only_name:
                print(f"{name} | gra {p.requires_grad} | ")
            else:
                d = p.data
                print(f"{name} | gra {p.requires_grad} |  min {d.min():.5f} mean {d.mean():.5f} max {d.max():.5f}")
        print("Total parameter number: ", total)

    # ------------------------------------------------------------------ flat parameter storage
    @property
    def flat_params(self) -> torch.Tensor:
        self._ensure_flat()
        return self._flat

    @property
    def flat_grads(self) -> torch.Tensor:
        self._ensure_flat()
        return self._flat_grad

    def _param_list(self):
        return list(self.named_parameters())

    _ALIGN = 8      # every tensor starts on a 32-byte boundary of the flat buffer (128-bit vector loads)

    def _layout(self):
        """[(name, param, offset)] and the padded total length of the flat buffer."""
        out, off = [], 0
        for name, p in self._param_list():
            out.append((name, p, off))
            off += -(-p.numel() // self._ALIGN) * self._ALIGN
        return out, off

    def _flat_ok(self) -> bool:
        if self._flat is None:
            return False
        lay, total = self._layout()
        for _, p, off in lay:
            if p.device != self._flat.device or p.data_ptr() != self._flat.data_ptr() + off * 4:
                return False
        return total == self._flat.numel()

    def _ensure_flat(self):
        """(Re)build the flat parameter / gradient buffers, e.g. after `.to(device)`."""
        if self._flat_ok():
            return
        lay, n = self._layout()
        dev = lay[0][1].device
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        grad = torch.zeros(n, dtype=torch.float32, device=dev)
        for _, p, off in lay:
            m = p.numel()
            flat[off:off + m].copy_(p.data.reshape(-1).to(torch.float32))
            if p.grad is not None:
                grad[off:off + m].copy_(p.grad.reshape(-1))
            p.data = flat[off:off + m].view(p.shape)
            p.grad = grad[off:off + m].view(p.shape)
        self._flat, self._flat_grad = flat, grad
        self._synced_version = -1
        if self._handle is not None:
            L.load().bn_mlp_destroy(self._handle)
            self._handle = None

    def rebind_grads(self, bucket: torch.Tensor):
        """Make `bucket` (fp32, same length and device as `flat_params`) the flat gradient buffer: every p.grad becomes a
        view of it.  Used by the data-parallel trainer to place the bucket in peer-mapped memory (brdf_nerf_b200.ddp)."""
        self._ensure_flat()
        if bucket.numel() != self._flat.numel() or bucket.device != self._flat.device or bucket.dtype != torch.float32:
            raise ValueError("gradient bucket must match flat_params in length, device and dtype")
        bucket.copy_(self._flat_grad)
        for _, p, off in self._layout()[0]:
            p.grad = bucket[off:off + p.numel()].view(p.shape)
        self._flat_grad = bucket

    def offsets(self) -> Dict[str, int]:
        return {name: off for name, _, off in self._layout()[0]}

    def grad_views(self, flat: torch.Tensor):
        """Per-parameter views of a flat gradient buffer laid out like `flat_params`."""
        return [flat[off:off + p.numel()].view(p.shape) if p.requires_grad else None for _, p, off in self._layout()[0]]

    # ------------------------------------------------------------------ C handle
    def _linear_names(self) -> List[Optional[str]]:
        names: List[Optional[str]] = [None] * L.BN_NUM_LINEAR
        for l in range(self.layers):
            names[l] = f"fc_net.{2 * l}"
        names[L.BN_LIN_SIGMA] = "sigma_from_xyz.0"
        names[L.BN_LIN_FEATS] = "feats_from_xyz"
        names[L.BN_LIN_RGB0] = "rgb_from_xyzdir.0"
        names[L.BN_LIN_RGB2] = "rgb_from_xyzdir.2"
        if hasattr(self, "grad_from_xyz"):
            names[L.BN_LIN_GRAD] = "grad_from_xyz"
        for h, hn in enumerate(L.HEAD_NAMES):
            if hasattr(self, f"{hn}_from_xyz"):
                names[L.BN_LIN_HEAD0 + 2 * h] = f"{hn}_from_xyz.0"
                names[L.BN_LIN_HEAD0 + 2 * h + 1] = f"{hn}_from_xyz.2"
        return names

    def handle(self):
        """Opaque bn_mlp handle (created lazily on the module's CUDA device)."""
        self._ensure_flat()
        if not self._flat.is_cuda:
            raise L.BnError("SpSBRDFNeRF must live on a CUDA device: there is no CPU path")
        lib = L.load()
        if self._handle is None:
            cfg = L.MlpCfg()
            cfg.feat, cfg.layers = self.feat, self.layers
            cfg.skip_layer = self.skips[0] if self.skips else -1
            cfg.n_freq_xyz = self.n_freq
            cfg.normal_lr = int(hasattr(self, "grad_from_xyz"))
            cfg.viewdir, cfg.n_freq_dir = int(self.viewdir), int(self.n_freq_dir)
            cfg.t_dims = int(self.t_embedding_dims) if self.beta == True else 0      # noqa: E712
            for h, hn in enumerate(L.HEAD_NAMES):
                mod = getattr(self, f"{hn}_from_xyz", None)
                cfg.head_dim[h] = mod[2].out_features if mod is not None else 0
            cfg.precision = L.BN_PREC_BF16 if self.precision == "bf16" else L.BN_PREC_FP32
            offs = self.offsets()
            for i, nm in enumerate(self._linear_names()):
                cfg.w_off[i] = offs[nm + ".weight"] if nm else -1
                cfg.b_off[i] = offs[nm + ".bias"] if nm else -1
            cfg.n_params = self._flat.numel()
            h = C.c_void_p()
            with torch.cuda.device(self._flat.device):
                L.check(lib.bn_mlp_create(C.byref(cfg), C.byref(h)))
            self._handle = h
            self._synced_version = -1
        return self._handle

    def _weights_version(self):
        """Changes whenever a parameter is written in place through autograd-visible tensors: `opt.step()`, `p.add_()`,
        `load_state_dict` (copy_ into the views).  The flat buffer's own counter does NOT move when a parameter view is
        updated (every view has its own counter), so the views' counters are part of the key.  Writes through `p.data`
        bypass every counter: call `mark_weights_dirty()` after those."""
        return (self._flat._version, sum(p._version for p in self.parameters()), self._dirty)

    def mark_weights_dirty(self):
        """The master weights were changed behind torch's version counters (raw kernels, `p.data` writes)."""
        self._dirty += 1

    def sync_weights(self, force=False):
        """Refresh the packed (bf16 / transposed) weight copies if the master weights changed."""
        h = self.handle()
        v = self._weights_version()
        if force or v != self._synced_version:
            L.check(L.load().bn_mlp_sync_weights(h, L.ptr(self._flat), L.stream_ptr()))
            self._synced_version = v

    def _load_from_state_dict(self, *a, **kw):
        super()._load_from_state_dict(*a, **kw)
        self._dirty += 1

    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16"):
            raise ValueError(precision)
        if precision != self.precision:
            self.precision = precision
            if self._handle is not None:
                L.load().bn_mlp_destroy(self._handle)
                self._handle = None

    def __del__(self):
        try:
            if self._handle is not None:
                L.load().bn_mlp_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ evaluation
    def mlp_flags(self, sigma_only=False, apply_brdf=False, apply_theta=False, nr_an_on=False, nr_lr_on=False,
                  train=False) -> int:
        f = 0
        if sigma_only:
            return L.MLP_SIGMA_ONLY
        if train:
            f |= L.MLP_TRAIN
        if self.beta == True:                                     # noqa: E712  every full forward emits the channel (:708-711)
            f |= L.MLP_BETA
        if nr_an_on:
            f |= L.MLP_NORMAL_AN
        if nr_lr_on:
            f |= L.MLP_NORMAL_LR
        if apply_brdf:
            if self.roughness == True:                            # noqa: E712
                f |= L.MLP_ROUGH
            elif self.RPV:
                f |= L.MLP_RPV
            else:               # reference else-branch (spsbrdfnerf.py:742-755): b, c and theta are emitted independently,
                want_theta = bool(apply_theta and self.args.theta == True)       # noqa: E712  e.g. shell_hapke + theta only
                if self.args.b == True or self.args.c == True or want_theta:     # noqa: E712
                    f |= L.MLP_HAPKE
                if want_theta:
                    f |= L.MLP_HAPKE_THETA
        return f

    def out_channels(self, flags: int) -> int:
        n = L.load().bn_mlp_out_channels(self.handle(), flags)
        if n < 0:
            L.check(n)
        return n

    def workspace(self, n_points: int, flags: int, tag: Optional[str] = "ws") -> torch.Tensor:
        """MLP workspace of a call.  `tag=None`: a fresh buffer owned by the caller (the autograd bridges keep it alive in
        their ctx, so a later forward cannot overwrite the activations an earlier backward still needs); a tag names a
        cached buffer that is reused by every call with that tag (tape-free Trainer path, no-grad inference)."""
        need = L.load().bn_mlp_workspace_bytes(self.handle(), n_points, flags)
        if tag is None:
            return torch.empty(int(need) + 256, dtype=torch.uint8, device=self._flat.device)
        ws = self._ws_cache.get(tag)
        if ws is None or ws.numel() < need or ws.device != self._flat.device:
            ws = torch.empty(int(need * 1.0) + 256, dtype=torch.uint8, device=self._flat.device)
            self._ws_cache[tag] = ws
        return ws

    def forward(self, input_xyz_, input_dir=None, input_sun_dir=None, input_t=None, sigma_only=False,
                apply_brdf=False, apply_theta=False, nr_an_on=False, nr_lr_on=False, sun_ray=False, mode="train"):
        """Reference signature (spsbrdfnerf.py:662): (B,3) points -> (B,1) sigma or packed (B,C)."""
        from ..autograd import PointsFunction
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.viewdir and not sigma_only and input_dir is None:
            raise ValueError("this model was built with input_viewdir=1: forward needs input_dir (spsbrdfnerf.py:689-690)")
        dirs = input_dir.contiguous() if (self.viewdir and input_dir is not None) else None
        if self.beta == True and not sigma_only and input_t is None:      # noqa: E712
            raise ValueError("this model was built with beta=True: forward needs input_t (spsbrdfnerf.py:708-709)")
        t = input_t.detach().float().contiguous() if (self.beta == True and input_t is not None) else None   # noqa: E712
        return PointsFunction.apply(self, input_xyz_.contiguous(), dirs, t, bool(sigma_only), bool(apply_brdf),
                                    bool(apply_theta), bool(nr_an_on), bool(nr_lr_on), need_grad, *self.parameters())
