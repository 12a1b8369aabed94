"""brdf_nerf_b200 — B200-native (sm_100a) implementation of BRDF-NeRF's ray-rendering hot path.

Drop-in surface (same names / arguments as the reference):
    brdf_nerf_b200.rendering.render_rays(models, args, rays, ts, ...)
    brdf_nerf_b200.models.load_model(args) -> SpSBRDFNeRF
Everything below that surface runs through the C-ABI library `libbrdfnerf_b200.so`
(include/brdfnerf_b200.h); there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
