"""Model factory with the reference's name (reference models/__init__.py:6-17)."""
from .spsbrdfnerf import SpSBRDFNeRF


def load_model(args, precision: str = "fp32"):
    """`precision`: 'fp32' (CUDA-core parity mode) or 'bf16' (tcgen05 throughput mode)."""
    if args.model != "spsbrdf-nerf":
        raise ValueError(f"model {args.model} is not implemented on the B200 hot path (only spsbrdf-nerf)")
    return SpSBRDFNeRF(args, layers=args.fc_layers, mapping=args.mapping, feat=args.fc_feat,
                       t_embedding_dims=args.t_embbeding_tau, beta=args.beta, roughness=args.roughness,
                       normal=args.normal, indirect_light=args.indirect_light, glossy_scale=args.glossy_scale,
                       sun_v=args.sun_v, MultiBRDF=args.MultiBRDF, dim_RPV=args.dim_RPV, siren=args.siren,
                       precision=precision)
