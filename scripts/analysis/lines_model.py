#!/usr/bin/env python3
"""Line-touch model of the evaluation sweep's gathers under a REAL policy (captured by scripts/k5_policies.py into
gpurun_out/k5_policies.npz): how many 128-byte lines does one 32-lane gather request touch, for the plain layout, for
x-lines padded to 24 / 32 floats, and for a pair shadow P[i] = (V[i], V[i+1])?  The L1 data pipe pays one wavefront per
line per request, so this predicts the pipe cost of a layout before any kernel is written.
    python scripts/analysis/reach.py && python scripts/analysis/lines_model.py policy_16"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]

# ---- padded x-lines ----
disp = np.load(ROOT / "scripts" / "data" / "k5_disp.npy").astype(np.int64).reshape(9, 20,20,20,20,20, 6)
pol = np.load(ROOT / "gpurun_out" / "k5_policies.npz")[sys.argv[1]].reshape(20,20,20,20,20,20).astype(np.int64)  # (x,i1..i5)
rng = np.random.default_rng(0)
# sample contiguous chunks of the internal order (i1..i5, x): take full (i4,i5,x) slabs for random (i1,i2,i3)
def run(pad, lane_map="consec"):
    XS = pad
    tot_lines = 0; tot_req = 0
    for _ in range(60):
        i1, i2, i3 = rng.integers(0, 20, 3)
        a = pol[:, i1, i2, i3]                      # (x, i4, i5)
        a = np.transpose(a, (1, 2, 0))              # (i4, i5, x)
        d = disp[:, i1, i2, i3]                     # (A, i4, i5, 6)
        dd = np.take_along_axis(d[:, :, :, None, :], a[None, ..., None], axis=0)[0]   # (i4,i5,x,6)
        I4, I5, X = np.meshgrid(np.arange(20), np.arange(20), np.arange(20), indexing="ij")
        c = [np.clip(v + dd[..., k], 0, 18) for k, v in ((1, i1), (2, i2), (3, i3), (4, I4), (5, I5))]
        cx = np.clip(X + dd[..., 0], 0, 18)
        line = (((c[0] * 20 + c[1]) * 20 + c[2]) * 20 + c[3]) * 20 + c[4]
        base = (line * XS + cx).reshape(-1)         # internal order (i4,i5,x) flattened = consecutive states
        live = ((X > 0) & (X < 19)).reshape(-1)
        nst = base.size
        for w0 in range(0, nst - 31, 32):
            b = base[w0:w0 + 32][live[w0:w0 + 32]]
            if b.size == 0: continue
            for off in (0, XS, 20 * XS, 400 * XS + XS, 8000 * XS):     # a few window offsets
                for xb in (0, 1):
                    tot_lines += len(np.unique((b + off + xb) // 32))
                    tot_req += 1
    return tot_lines / tot_req
for pad in (20, 32, 24):
    print(sys.argv[1], "x-line pitch", pad, "lines per request", round(run(pad), 3))

# ---- pair shadow ----
disp = np.load(ROOT / "scripts" / "data" / "k5_disp.npy").astype(np.int64).reshape(9, 20,20,20,20,20, 6)
pol = np.load(ROOT / "gpurun_out" / "k5_policies.npz")[sys.argv[1]].reshape(20,20,20,20,20,20).astype(np.int64)
rng = np.random.default_rng(0)
tot = {"scalar": 0, "pair_shadow": 0}; nreq = {"scalar": 0, "pair_shadow": 0}; nwarps = 0
for _ in range(60):
    i1, i2, i3 = rng.integers(0, 20, 3)
    a = np.transpose(pol[:, i1, i2, i3], (1, 2, 0))
    d = disp[:, i1, i2, i3]
    dd = np.take_along_axis(d[:, :, :, None, :], a[None, ..., None], axis=0)[0]
    I4, I5, X = np.meshgrid(np.arange(20), np.arange(20), np.arange(20), indexing="ij")
    c = [np.clip(v + dd[..., k], 0, 18) for k, v in ((1, i1), (2, i2), (3, i3), (4, I4), (5, I5))]
    cx = np.clip(X + dd[..., 0], 0, 18)
    line = (((c[0] * 20 + c[1]) * 20 + c[2]) * 20 + c[3]) * 20 + c[4]
    base = (line * 20 + cx).reshape(-1)
    live = ((X > 0) & (X < 19)).reshape(-1)
    for w0 in range(0, base.size - 31, 32):
        b = base[w0:w0 + 32][live[w0:w0 + 32]]
        if b.size == 0: continue
        nwarps += 1
        for off in (0, 20, 400, 8000 + 20, 160000, 3200000 + 400):
            # scalar: two requests (x-bit 0 and 1), 4-byte elements, 128-byte lines
            for xb in (0, 1):
                tot["scalar"] += len(np.unique((b + off + xb) * 4 // 128)); nreq["scalar"] += 1
            # pair shadow: one request, element i is the 8-byte pair (V[i], V[i+1])
            lo = (b + off) * 8 // 128; hi = ((b + off) * 8 + 7) // 128
            tot["pair_shadow"] += len(np.unique(np.concatenate([lo, hi]))); nreq["pair_shadow"] += 1
print(sys.argv[1], {k: round(tot[k] / nreq[k], 3) for k in tot}, "lines per request;",
      "wavefronts per window per warp: scalar", round(2 * tot["scalar"] / nreq["scalar"], 2), "pair-shadow", round(tot["pair_shadow"] / nreq["pair_shadow"], 2))
