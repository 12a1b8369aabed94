"""ctypes binding of libbrdfnerf_b200.so (include/brdfnerf_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised.  Build it with `python -m brdf_nerf_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbrdfnerf_b200.so")

BN_NUM_LINEAR = 37
BN_NUM_HEADS = 8
BN_LIN_SIGMA, BN_LIN_FEATS, BN_LIN_RGB0, BN_LIN_RGB2, BN_LIN_GRAD, BN_LIN_HEAD0 = 16, 17, 18, 19, 20, 21
BN_PREC_FP32, BN_PREC_BF16 = 0, 1
BN_BRDF_NONE, BN_BRDF_MICROFACET, BN_BRDF_RPV, BN_BRDF_HAPKE = 0, 1, 2, 3
BN_IRR_ONES, BN_IRR_COS, BN_IRR_SUNVIS = 0, 1, 2
MLP_SIGMA_ONLY, MLP_TRAIN, MLP_NORMAL_AN, MLP_NORMAL_LR = 1, 2, 4, 8
MLP_ROUGH, MLP_RPV, MLP_HAPKE, MLP_HAPKE_THETA, MLP_BETA = 16, 32, 64, 128, 256
HEAD_NAMES = ("roughness", "k", "theta_rpv", "rhoc", "b", "c", "theta", "beta")     # BN_HEAD_* order


class ShadeCfg(C.Structure):
    _fields_ = [("n_channels", C.c_int), ("normal_ch", C.c_int), ("param_ch", C.c_int), ("brdf_ch", C.c_int),
                ("brdf_type", C.c_int), ("funcM", C.c_int), ("funcF", C.c_int), ("funcH", C.c_int),
                ("hapke_b", C.c_int), ("hapke_c", C.c_int), ("hapke_theta", C.c_int), ("shell_hapke", C.c_int),
                ("multi_brdf", C.c_int), ("irr_mode", C.c_int), ("hpk_scl", C.c_float), ("fresnel_f0", C.c_float)]


class MlpCfg(C.Structure):
    _fields_ = [("feat", C.c_int), ("layers", C.c_int), ("skip_layer", C.c_int), ("n_freq_xyz", C.c_int),
                ("normal_lr", C.c_int), ("viewdir", C.c_int), ("n_freq_dir", C.c_int), ("t_dims", C.c_int), ("head_dim", C.c_int * BN_NUM_HEADS), ("precision", C.c_int),
                ("w_off", C.c_int64 * BN_NUM_LINEAR), ("b_off", C.c_int64 * BN_NUM_LINEAR), ("n_params", C.c_int64)]


_P, _I, _F, _L, _Z, _D = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t, C.c_double

_SIGS = {
    "bn_abi_version": (C.c_int, []),
    "bn_last_error": (C.c_char_p, []),
    "bn_device_check": (C.c_int, [_I]),
    "bn_launch_count": (C.c_ulonglong, []),
    "bn_profile_enable": (C.c_int, [_I]),
    "bn_profile_collect": (C.c_int, [_I, _P, _P, _P]),
    "bn_sample_stratified": (C.c_int, [_P, _P, _I, _P, _P, _P, _I, _I, _P]),
    "bn_sample_guided": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P]),
    "bn_merge_samples": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "bn_lambertian_render_loss": (C.c_int, [_P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _F, _F, _I, _P, _P, _P, _P, _I, _I, _P]),
    "bn_coarse_to_fine": (C.c_int, [_P, _P, _P, _F, _P, _P, _P, _P, _P, _F, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    _I, _I, _I, _P]),
    "bn_sort_rows": (C.c_int, [_P, _P, _I, _I, _P]),
    "bn_permute_samples": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "bn_composite_sigma": (C.c_int, [_P, _P, _P, _F, _P, _P, _P, _P, _P, _I, _I, _P]),
    "bn_composite_forward": (C.c_int, [_P, _P, _I, _I, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P]),
    "bn_composite_backward": (C.c_int, [_P, _P, _I, _I, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P]),
    "bn_shade_rays_forward": (C.c_int, [C.POINTER(ShadeCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "bn_shade_rays_backward": (C.c_int, [C.POINTER(ShadeCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "bn_brdf_points_forward": (C.c_int, [C.POINTER(ShadeCfg), _P, _P, _P, _I, _I, _P]),
    "bn_brdf_points_backward": (C.c_int, [C.POINTER(ShadeCfg), _P, _P, _P, _I, _I, _P]),
    "bn_loss_color_depth": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _F, _F, _I, _P, _P, _P, _I, _I, _P]),
    "bn_loss_regularizers": (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _I, _F, _P, _F, _P, _P, _P, _P, _P, _I, _I, _P]),
    "bn_count_nan": (C.c_int, [_P, C.c_longlong, _P, _P]),
    "bn_mlp_create": (C.c_int, [C.POINTER(MlpCfg), C.POINTER(_P)]),
    "bn_mlp_destroy": (None, [_P]),
    "bn_mlp_sync_weights": (C.c_int, [_P, _P, _P]),
    "bn_mlp_out_channels": (C.c_int, [_P, _I]),
    "bn_mlp_workspace_bytes": (_Z, [_P, _L, _I]),
    "bn_mlp_forward": (C.c_int, [_P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _P, _Z, _P]),
    "bn_mlp_trunk_forward": (C.c_int, [_P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _L, _L, _P, _P, _Z, _P]),
    "bn_mlp_write_t": (C.c_int, [_P, _P, _I, _I, _I, _L, _L, _P, _Z, _P]),
    "bn_mlp_heads_forward": (C.c_int, [_P, _P, _L, _I, _P, _I, _P, _Z, _P]),
    "bn_mlp_backward": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "bn_mlp_normals_forward": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    "bn_mlp_normals_backward": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "bn_debug_chain_trace": (C.c_int, [_P, _P]),
    "bn_debug_ws_tensor": (C.c_int, [_P, _L, _I, _I, _I, _P, _P]),
    "bn_debug_gemm_epi": (C.c_int, [_I, _P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong, _P, _P, _P, _I, _I,
                                    C.c_longlong, _I, C.c_longlong, _P]),
    "bn_debug_gemm": (C.c_int, [_I, _I, _P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong, C.c_longlong, _I, C.c_longlong, _P]),
    "bn_allreduce_p2p": (C.c_int, [_P, _P, _P, _L, _I, _I, _I, _P]),
    "bn_allreduce_p2p_flag_words": (C.c_int, []),
    "bn_peer_alloc": (C.c_int, [_Z, C.POINTER(_P)]),
    "bn_peer_free": (C.c_int, [_P]),
    "bn_peer_export": (C.c_int, [_P, _P]),
    "bn_peer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "bn_peer_close": (C.c_int, [_P]),
    "bn_adam_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P]),
    "bn_adam_step_graph": (C.c_int, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P]),
    "bn_rays_from_rpc": (C.c_int, [_P, _P, _P, C.c_longlong, _I, _D, _D, _I, _I, _I, _F, _F, _F, _F, _P, _P, _I, _P, _P, _P]),
    "bn_dsm_points": (C.c_int, [_P, _I, _P, C.c_longlong, _D, _D, _D, _D, _I, _I, _P, _P, _P, _P, _P]),
    "bn_dsm_workspace_bytes": (_Z, [_I, _I, _I, _F]),
    "bn_dsm_rasterize": (C.c_int, [_P, _I, _I, C.c_longlong, _D, _D, _D, _I, _I, _I, _F, _P, _P, _P, _Z, _P]),
    "bn_dsm_accumulate": (C.c_int, [_P, _I, _I, C.c_longlong, _D, _D, _D, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "bn_dsm_finalize": (C.c_int, [_I, _I, _I, _F, _P, _Z, _P, _P, _P]),
    "bn_dsm_normals_from_points": (C.c_int, [_P, _I, _I, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m brdf_nerf_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            if name.startswith("bn_mlp_normals"):
                continue
            raise ImportError(f"{LIB_PATH} does not export {name}")
        fn.restype = res
        fn.argtypes = args
    if lib.bn_abi_version() != 1:
        raise ImportError("libbrdfnerf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGS)


class BnError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise BnError(f"libbrdfnerf_b200 error {rc}: {load().bn_last_error().decode()}")


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise BnError("brdf_nerf_b200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise BnError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise BnError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
