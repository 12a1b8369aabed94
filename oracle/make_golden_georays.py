"""ORACLE (test infrastructure only).  Generates tests/golden/georays.npz: a strided pixel subset of a 2048 x 2048 image of
the synthetic RPC camera (oracle/georays_np.synthetic_rpc) and
  * ref_rays_ecef       : output of the LIVE reference's get_rays (datasets/satellite_rgb_dep.py:23-78, cs='ecef') driven by
    the oracle's restatement of rpcm's RPC model (rpcm itself is not installed: everything in get_rays except
    rpc.localization is the reference's own code here);
  * ref_rays_ecef_norm  : the LIVE reference's normalize_rays on them, ref_sun: its get_sun_dirs row;
  * restated_rays_utm   : the oracle's get_rays with cs='utm' (the reference's default; its utm branch needs pyproj, absent:
    NOT a reference output, parity unpinned), restated_lonlat_max: the restated localisation at max_alt.
Run here:   python -m oracle.make_golden_georays
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import georays_np as G          # noqa: E402
from oracle import ref_harness as RH        # noqa: E402

MIN_ALT, MAX_ALT = -25.0, 95.0
CENTER_ECEF = (799000.0, -5452800.0, 3200200.0)
RANGE = 400.0
SUN = (62.5, 148.0)


def main():
    rpc = G.synthetic_rpc(0)
    cols, rows = np.meshgrid(np.arange(0, 2048, 41), np.arange(0, 2048, 43))
    cols, rows = cols.flatten().astype(np.float64), rows.flatten().astype(np.float64)
    ref = RH.ref_get_rays(cols, rows, rpc, MIN_ALT, MAX_ALT, cs="ecef")
    ref_n = RH.ref_normalize_rays(ref, RANGE, CENTER_ECEF)
    sun = RH.ref_sun_dirs(SUN[0], SUN[1], 1)
    utm = G.get_rays(cols, rows, rpc, MIN_ALT, MAX_ALT, cs="utm")
    lon, lat = rpc.localization(cols, rows, np.full(cols.shape, MAX_ALT))
    path = os.path.join(ROOT, "tests", "golden", "georays.npz")
    np.savez_compressed(path, cols=cols, rows=rows, min_alt=MIN_ALT, max_alt=MAX_ALT, center=np.asarray(CENTER_ECEF),
                        scene_range=RANGE, sun_el_az=np.asarray(SUN), ref_rays_ecef=ref.numpy(), ref_rays_ecef_norm=ref_n.numpy(),
                        ref_sun=sun.numpy(), restated_rays_utm=utm, restated_lonlat_max=np.stack([lon, lat], 1))
    print(f"georays: {cols.size} pixels, {rpc.last_iterations} iterations, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
