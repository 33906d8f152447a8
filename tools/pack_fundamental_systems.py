"""Repacks the reference's precomputed fundamental systems (DATA tables shipped as package data next to
gpfy/harmonics/fund_set.py: fundamental_system/fs_{dim}D.npz, key degree_{n} -> [N(dim, n), dim] unit
vectors) into ONE compressed archive gpfy_b200/data/fundamental_systems.npz with keys d{dim}_n{degree}.

Run in the build container only (needs /root/reference):  python tools/pack_fundamental_systems.py
Identical V_l are required for parity: the harmonics of a complete degree are defined by these points
(spherical_harmonics.py:119-122), so they are inputs of the path, like the golden vectors."""
import glob
import os
import re

import numpy as np

SRC = "/root/reference/src/gpfy/fundamental_system"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpfy_b200", "data", "fundamental_systems.npz")

out = {}
for f in sorted(glob.glob(os.path.join(SRC, "fs_*D.npz"))):
    dim = int(re.search(r"fs_(\d+)D", f).group(1))
    with np.load(f) as z:
        for k in z.files:
            out[f"d{dim}_n{int(k.split('_')[1])}"] = np.asarray(z[k], dtype=np.float64)
np.savez_compressed(DST, **out)
print(f"{len(out)} tables -> {DST} ({os.path.getsize(DST) / 1e6:.2f} MB)")
