#!/usr/bin/env python3
"""Plane-staged sweep, model 1: how many (x, xdot) V-planes does a state-plane of K5 need under a real policy,
and how many of them are new when a CTA walks consecutive state-planes and keeps the previous step's planes in its slots?
(36.5 needed, 15.6 new per state-plane; the planner measures 14.9.)
Inputs: scripts/data/k5_disp.npy (python scripts/analysis/reach.py) and scripts/data/k5_policy16.npz (the converged K5 policy,
scripts/k5_policies.py).  CPU only."""
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
import numpy as np, sys, itertools, collections
disp = np.load(str(ROOT / 'scripts' / 'data') + '/k5_disp.npy').astype(np.int16).reshape(9,20,20,20,20,20,6)  # (A, xd,t1,w1,t2,w2, 6)
# outer displacement independent of xd?
o = disp[..., 2:]
print("outer disp independent of xd:", bool((o == o[:, :1]).all()))
od = o[:, 0]   # (A, t1,w1,t2,w2, 4)
pol = np.load(str(ROOT / 'scripts' / 'data') + '/k5_policy16.npz')['policy'].reshape(20,20,20,20,20,20)  # x,xd,t1,w1,t2,w2
# per plane action histogram
A = 9
cnt = np.zeros((A,20,20,20,20), np.int32)
for a in range(A):
    cnt[a] = (pol == a).sum(axis=(0,1))
nact = (cnt > 0).sum(0)
print("actions per plane: hist", np.bincount(nact.ravel(), minlength=10))
print("mean actions per plane", nact.mean())
# states covered if we only stage the top-k actions per plane
srt = np.sort(cnt, axis=0)[::-1]
for k in range(1, 6):
    print("top", k, "actions cover", srt[:k].sum() / cnt.sum())
idx = np.stack(np.meshgrid(*[np.arange(20)]*4, indexing='ij'), -1)  # (20,20,20,20,4)
cell = np.clip(idx[None] + od, 0, 18)     # (A, ..., 4) lower-corner plane coords
np.save(str(ROOT / 'scripts' / 'data' / 'k5_cell.npy'), cell.astype(np.int8)); np.save(str(ROOT / 'scripts' / 'data' / 'k5_cnt.npy'), cnt)
# distinct cells per plane among used actions
def plane_id(c): return ((c[...,0]*20 + c[...,1])*20 + c[...,2])*20 + c[...,3]
cid = plane_id(cell)   # (A, 20,20,20,20)
used = cnt > 0
dist = np.zeros((20,20,20,20), int)
for p in np.ndindex(20,20,20,20):
    dist[p] = len(set(cid[:, p[0],p[1],p[2],p[3]][used[:, p[0],p[1],p[2],p[3]]].tolist()))
print("distinct successor 4D-cells per plane: hist", np.bincount(dist.ravel(), minlength=10), "mean", dist.mean())
# distinct V planes per plane (union of 16 corners over used actions)
offs = np.array([[(c>>k)&1 for k in range(4)] for c in range(16)])
tot = 0
for p in np.ndindex(20,20,20,20):
    s = set()
    for a in np.nonzero(used[(slice(None),)+p])[0]:
        c = cell[(a,)+p]
        for o_ in offs:
            s.add(tuple(c + o_))
    tot += len(s)
print("mean distinct V-planes needed per state-plane:", tot/160000, "-> L2->smem bytes per sweep w/o reuse across planes: %.2f GB" % (tot*1600/1e9))


# ---- slot reuse between consecutive state-planes (what csrc/plane_plan.cuh does)
cell = np.load(str(ROOT / 'scripts' / 'data' / 'k5_cell.npy')).astype(np.int64)   # (A,20,20,20,20,4)
cnt = np.load(str(ROOT / 'scripts' / 'data' / 'k5_cnt.npy'))                       # (A,20,20,20,20) states of each action per plane
offs = np.array([[(c>>k)&1 for k in range(4)] for c in range(16)])
cid = ((cell[...,0]*20 + cell[...,1])*20 + cell[...,2])*20 + cell[...,3]
# population per distinct cell per plane
def cells_of(p):
    d = collections.Counter()
    for a in range(9):
        n = cnt[(a,)+p]
        if n: d[int(cid[(a,)+p])] += int(n)
    return d
def vps(c):
    t = np.array([(c//8000)%20, (c//400)%20, (c//20)%20, c%20])
    out = []
    for o_ in offs:
        u = t + o_
        out.append(int(((u[0]*20+u[1])*20+u[2])*20+u[3]))
    return out
rng = np.random.default_rng(2)
def simulate(T, nsamp=400, L=20, mode='prev'):
    loads = planes = 0; fb = 0; tot = 0; needsz = []; unionsz = []; kcells=[]
    for _ in range(nsamp):
        p0 = [int(v) for v in rng.integers(0, 20, 3)]
        prev = set()
        for i in range(L):
            p = (p0[0], p0[1], p0[2], i)
            d = cells_of(p)
            need = set(); k = 0
            for c, n in d.items():
                tot += n
                if n < T: fb += n; continue
                k += 1
                need.update(vps(c))
            new = need - prev
            loads += len(new); planes += 1
            needsz.append(len(need)); unionsz.append(len(need | prev)); kcells.append(k)
            prev = need
    return loads/planes, fb/tot, np.percentile(needsz, [50, 90, 100]), np.percentile(unionsz, [50, 90, 99, 100]), np.percentile(kcells,[50,90,100])
for T in (1, 2, 4, 8, 16, 32):
    l, f, ns, us, kc = simulate(T)
    print("T", T, "loads/plane %.1f (%.2f GB/sweep)" % (l, l*160000*1600/1e9), "fallback states %.3f%%" % (100*f), "need p50/p90/max", ns, "union(prev,cur) p50/90/99/max", us, "cells", kc)
