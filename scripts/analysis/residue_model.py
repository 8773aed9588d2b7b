#!/usr/bin/env python3
"""Plane-staged sweep, model 3 (not built): residue-balanced warps — if the planner permuted the states of a plane so that
every warp holds 32 distinct (in-plane offset mod 32), how many warps would a plane need?  (15.4 instead of 13; natural order pays a
mean conflict degree of 1.72.)
Inputs: scripts/data/k5_disp.npy (python scripts/analysis/reach.py) and scripts/data/k5_policy16.npz (the converged K5 policy,
scripts/k5_policies.py).  CPU only."""
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
import numpy as np
disp = np.load(str(ROOT / 'scripts' / 'data') + '/k5_disp.npy').astype(np.int16).reshape(9,20,20,20,20,20,6)  # (A, xd,t1,w1,t2,w2, 6) for x index 10
pol = np.load(str(ROOT / 'scripts' / 'data') + '/k5_policy16.npz')['policy'].reshape(20,20,20,20,20,20)  # x,xd,t1,w1,t2,w2
rng = np.random.default_rng(0)
X, XD = np.meshgrid(np.arange(20), np.arange(20), indexing='ij')   # x, xd
mx = []; mx_nat = 0; cnt = 0; confl_nat = []
for _ in range(3000):
    p = tuple(int(v) for v in rng.integers(0, 20, 4))
    a = pol[(slice(None), slice(None)) + p]             # (x, xd)
    d = disp[(slice(None), slice(None)) + p]            # (A, xd, 6)
    dx = d[a, XD, 0]; dxd = d[a, XD, 1]
    live = (X > 0) & (X < 19)                            # x edges are absorbing (|x| > 2.4)
    cx = np.clip(X + dx, 0, 18); cxd = np.clip(XD + dxd, 0, 18)
    ip = cxd * 20 + cx
    # storage order within plane: xd slow, x fast -> state index s = xd*20 + x
    s = (XD * 20 + X)
    order = np.argsort(s.ravel())
    ipf = ip.ravel()[order]; lv = live.ravel()[order]
    res = ipf[lv] % 32
    b = np.bincount(res, minlength=32)
    mx.append(b.max())
    # natural order conflicts: per warp of 32 consecutive states, max multiplicity of residue among live lanes (distinct addresses)
    for w0 in range(0, 400, 32):
        r = ipf[w0:w0+32][lv[w0:w0+32]]
        if r.size == 0: continue
        u = np.unique(r)          # same address -> broadcast
        m = np.bincount(u % 32, minlength=32).max()
        confl_nat.append(m)
mx = np.array(mx)
print("live states per plane: 360; ideal warps 11.25 (12)")
print("residue-balanced warps per plane: mean %.2f, p50 %d p90 %d max %d" % (mx.mean(), np.percentile(mx,50), np.percentile(mx,90), mx.max()))
print("natural order: mean conflict degree per warp-LDS (same slot parity assumed): %.3f; hist" % np.mean(confl_nat), np.bincount(confl_nat))
