#!/usr/bin/env python3
"""Plane-staged sweep, model 2: shared-memory bank-conflict degree of a warp's corner load in natural state order
under the converged K5 policy (mean 1.86), and with every cell's V-planes bank-aligned (1.18) — the experiment that was built,
measured (conflicts 99 M -> 79 M, but more loads) and removed.
Inputs: scripts/data/k5_disp.npy (python scripts/analysis/reach.py) and scripts/data/k5_policy16.npz (the converged K5 policy,
scripts/k5_policies.py).  CPU only."""
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
import numpy as np
disp = np.load(str(ROOT / 'scripts' / 'data') + '/k5_disp.npy').astype(np.int16).reshape(9,20,20,20,20,20,6)
pol = np.load(str(ROOT / 'scripts' / 'data') + '/k5_policy16.npz')['policy'].reshape(20,20,20,20,20,20)
rng = np.random.default_rng(0)
X, XD = np.meshgrid(np.arange(20), np.arange(20), indexing='ij')   # arrays indexed [x, xd]
deg_un=[]; deg_al=[]; ncell_w=[]
for _ in range(2000):
    p = tuple(int(v) for v in rng.integers(0, 20, 4))
    a = pol[(slice(None), slice(None)) + p]             # [x, xd]
    d = disp[(slice(None), slice(None)) + p]            # (A, xd, 6)
    dx = d[a, XD, 0]; dxd = d[a, XD, 1]
    live = (X > 0) & (X < 19)
    cx = np.clip(X + dx, 0, 18); cxd = np.clip(XD + dxd, 0, 18)
    ip = cxd * 20 + cx
    t = XD * 20 + X
    # phase per cell(action): representative nearest the middle
    phase = {}
    for act in np.unique(a[live]):
        m = live & (a == act)
        tt = t[m]; pp = (ip[m] - tt) % 32
        i = np.argmin(np.abs(2*tt - 400))
        phase[int(act)] = int(pp[i])
    order = np.argsort(t.ravel())
    ipf = ip.ravel()[order]; lv = live.ravel()[order]; af = a.ravel()[order]
    al = np.array([(32 - ((phase.get(int(x), 0) + 2) & 28)) & 31 for x in af])
    for w0 in range(0, 400, 32):
        sl = slice(w0, w0+32)
        m = lv[sl]
        if m.sum() == 0: continue
        # unaligned: each cell reads its own slot (distinct planes) -> distinct addresses even if ip equal; bank = ip mod 32 (+slot parity ignored)
        key_un = np.stack([af[sl][m], ipf[sl][m]], 1)
        u = np.unique(key_un, axis=0)
        deg_un.append(np.bincount(u[:,1] % 32, minlength=32).max())
        key_al = np.stack([af[sl][m], (ipf[sl][m] + al[sl][m])], 1)
        u = np.unique(key_al, axis=0)
        deg_al.append(np.bincount(u[:,1] % 32, minlength=32).max())
        ncell_w.append(len(np.unique(af[sl][m])))
print("unaligned mean degree %.3f hist" % np.mean(deg_un), np.bincount(deg_un))
print("aligned   mean degree %.3f hist" % np.mean(deg_al), np.bincount(deg_al))
print("cells per warp hist", np.bincount(ncell_w))
