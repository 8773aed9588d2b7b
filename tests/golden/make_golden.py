#!/usr/bin/env python3
"""
tests/golden/make_golden.py — regenerate the committed fixtures from the reference tree.

The reference ships exactly two result-pinning artefacts
(runners/results/mountain_car_cuda_policy.npz and
runners/results/continuous_mountain_car_cuda_policy.npz: 200x200 grids, all nine
arrays of the saved-policy format).  They are copied here in reduced form
(compressed; `states_space` — which is a pure function of the grid — replaced by
its SHA-256) so the CPU test-suite can pin the oracle and the file format
without /root/reference being present.

Run in the build container:  python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

for env in ("mountain_car", "continuous_mountain_car"):
    src = REF / "runners" / "results" / f"{env}_cuda_policy.npz"
    d = np.load(src)
    out = {k: d[k] for k in d.files if k != "states_space"}
    ss = np.ascontiguousarray(d["states_space"])
    out["states_space_sha256"] = np.frombuffer(hashlib.sha256(ss.tobytes()).digest(), dtype=np.uint8)
    out["states_space_shape"] = np.array(ss.shape, dtype=np.int64)
    out["file_keys"] = np.array(d.files)
    out["file_dtypes"] = np.array([str(d[k].dtype) for k in d.files])
    np.savez_compressed(HERE / f"{env}_golden.npz", **out)
    print(env, {k: (v.dtype, v.shape) for k, v in out.items()})

# --- utils/barycentric.py (the reference's numba CPU lookup, importable here) -------------
# Golden (weights, indices) for random points incl. out-of-range ones; pins
# oracle_inference_weights and the N2 lookup kernel.
sys.path.insert(0, str(REF))
from itertools import product  # noqa: E402

from utils.barycentric import get_barycentric_weights_and_indices  # noqa: E402

rng = np.random.default_rng(20261017)
inf = {}
for D, bins in ((2, 37), (4, 11), (6, 5)):
    lo = np.array([-1.2, -0.07, -np.pi, -15.0, -2.5, -8.0][:D], dtype=np.float32)
    hi = np.array([0.6, 0.07, np.pi, 15.0, 2.5, 8.0][:D], dtype=np.float32)
    shape = np.full(D, bins, dtype=np.int32)
    shape[0] += 3
    strides = np.ones(D, dtype=np.int64)
    for d in range(D - 2, -1, -1):
        strides[d] = strides[d + 1] * shape[d + 1]
    strides = strides.astype(np.int32)
    corner_bits = np.array(list(product([0, 1], repeat=D)), dtype=np.int32)
    pts = (lo + (hi - lo) * rng.uniform(-0.1, 1.1, size=(1500, D))).astype(np.float32)
    pts[:D] = lo  # exact corners
    pts[D:2 * D] = hi
    w, idx = get_barycentric_weights_and_indices(pts, lo, hi, shape, strides, corner_bits)
    inf.update({f"d{D}_lo": lo, f"d{D}_hi": hi, f"d{D}_shape": shape, f"d{D}_strides": strides,
                f"d{D}_corner_bits": corner_bits, f"d{D}_points": pts, f"d{D}_weights": w, f"d{D}_indices": idx})
np.savez_compressed(HERE / "barycentric_inference_golden.npz", **inf)
print("barycentric_inference_golden:", {k: v.shape for k, v in inf.items() if "weights" in k})
