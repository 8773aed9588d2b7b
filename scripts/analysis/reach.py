#!/usr/bin/env python3
"""Successor-cell displacement of every (state, action) at K5 for one x slab (the dynamics are translation-invariant in
x), computed with the CPU oracle's step function: per-dimension histograms and the table scripts/data/k5_disp.npy that
lines_model.py and box_model.py read.   python scripts/analysis/reach.py"""
import sys, numpy as np, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
(ROOT / "scripts" / "data").mkdir(exist_ok=True)
from oracle import cpu_oracle
from dynamicprogramming_b200 import envs
spec = envs.REGISTRY["double_cartpole_swingup"]
bins = 20
axes = [np.asarray(v, np.float32) for v in spec.bins_space(bins).values()]
# one x value (index 10): 3.2M states over the other 5 dims
sub = [axes[0][10:11]] + axes[1:]
o = cpu_oracle.CpuPolicyIteration("double_cartpole_swingup", sub, spec.actions, 0.999, 1e-4, 10, 1)
# patch the grid to the full one
st = 1
for d in range(5, -1, -1):
    o.grid.shape[d] = bins; o.grid.strides[d] = st; o.grid.lo[d] = float(axes[d][0]); o.grid.hi[d] = float(axes[d][-1]); st *= bins
n = o.n_states
S = o.states_space
# state indices per dim
sidx = np.stack([np.clip(np.rint((S[:, d] - axes[d][0]) / (axes[d][-1] - axes[d][0]) * (bins - 1)), 0, bins - 1).astype(int) for d in range(6)], 1)
Cn = 64
idx = np.empty((n, Cn), np.int32); w = np.empty((n, Cn), np.float32); r = np.empty(n, np.float32); t = np.empty(n, np.uint8); nxt = np.empty((n, 6), np.float32)
alld = []
for a in range(len(spec.actions)):
    o.lib.oracle_rows(C.byref(o.grid), o.step, cpu_oracle._p(S, C.c_float), C.c_int64(n), C.c_float(float(spec.actions[a])),
                      cpu_oracle._p(idx, C.c_int32), cpu_oracle._p(w, C.c_float), cpu_oracle._p(r, C.c_float), cpu_oracle._p(t, C.c_uint8), cpu_oracle._p(nxt, C.c_float))
    base = idx[:, 0].astype(np.int64)
    cell = np.stack([(base // (bins ** (5 - d))) % bins for d in range(6)], 1)
    disp = cell - sidx
    alld.append(disp)
    print("action", a, spec.actions[a], "terminated frac", t.mean())
    for d in range(6):
        v, c = np.unique(disp[:, d], return_counts=True)
        print("   dim", d, dict(zip(v.tolist(), (c / n).round(4).tolist())))
alld = np.stack(alld, 0)  # (A, n, 6)
np.save(ROOT / "scripts" / "data" / "k5_disp.npy", alld.astype(np.int8))
print("spread across actions (max-min disp) per dim:")
sp = alld.max(0) - alld.min(0)
for d in range(6):
    v, c = np.unique(sp[:, d], return_counts=True)
    print("   dim", d, dict(zip(v.tolist(), (c / n).round(4).tolist())))
