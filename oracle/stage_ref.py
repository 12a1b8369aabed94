"""ORACLE (test infrastructure only — never imported by the product path).

Stages the UNMODIFIED reference files of the hot path from the read-only tree (/root/reference, or $BRDFNERF_REF) into
`oracle/_ref/` so that they travel to the GPU box with the snapshot (the directory is git-ignored: reference sources never
enter this repository's history; it is NOT gpurun-ignored).  On the box `bench.py --impl reference`, bench.py's
`reference_cuda_eager` leg and `tests/test_gpu_reference_live.py` run the real thing from there:

    python -m oracle.stage_ref          # also run by __graft_entry__.build() whenever the reference tree is present

Staged: rendering.py, train_utils.py, metrics.py, models/*.py, BRDF/*.py — the import closure of
`rendering.render_rays`, `models.load_model` and the losses; byte-identical copies, checked by sha256.
`rasterio` and `kornia` (imported by train_utils.py:9 / metrics.py:7, unused on the path, not installed) are stubbed at
import time by oracle/ref_harness.py, exactly as for the live tree.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("BRDFNERF_REF", "/root/reference")
FILES = ["rendering.py", "train_utils.py", "metrics.py", "models/__init__.py", "models/nerf.py", "models/satnerf.py",
         "models/snerf.py", "models/spsbrdfnerf.py", "BRDF/Hapke.py", "BRDF/RPV.py", "BRDF/basic_func.py",
         "BRDF/microfacet.py", "LICENSE"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose=True) -> str:
    if not os.path.isfile(os.path.join(SRC, "rendering.py")):
        raise RuntimeError(f"reference tree not found at {SRC}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
        assert manifest[rel] == _sha(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference files into {DST}", file=sys.stderr)
    return DST


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "rendering.py")) and os.path.isfile(os.path.join(DST, "MANIFEST.json"))


if __name__ == "__main__":
    stage()
