#!/usr/bin/env python3
"""Could a tile of states stage its V neighbourhood in shared memory?  For tiles of x-lines (boxes in the five non-x
dimensions) under a REAL policy: bounding box of the successor cells and the exact number of distinct V lines, against
the 2 560 lines 200 KB of shared memory hold.  Result for the converged K5 policy: a 2x2x2x2x2 tile needs a 1 296-line
box (median) for 32 x-lines of states — no reuse to harvest.
    python scripts/analysis/reach.py && python scripts/analysis/box_model.py gpurun_out/k5_policies.npz policy_16"""
import sys, numpy as np, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
disp = np.load(ROOT / "scripts" / "data" / "k5_disp.npy").astype(np.int32)      # (A, 20^5, 6)  dims order: x, xdot, th1, w1, th2, w2 ; index over (i1..i5) row-major
A = disp.shape[0]
disp = disp.reshape(A, 20, 20, 20, 20, 20, 6)
pol = np.load(sys.argv[1])[sys.argv[2]].reshape(20, 20, 20, 20, 20, 20)   # (x, i1..i5)
rng = np.random.default_rng(0)
def analyse(tile, n_samples=400):
    t = np.array(tile)
    nt = 20 // t   # tiles per dim (assume divides or floor)
    vols, exact, nst = [], [], []
    for _ in range(n_samples):
        o = rng.integers(0, nt) * t
        sl = tuple(slice(o[k], o[k] + t[k]) for k in range(5))
        p = pol[(slice(None),) + sl]                   # (20, t1..t5) actions
        idx = np.stack(np.meshgrid(*[np.arange(o[k], o[k] + t[k]) for k in range(5)], indexing="ij"), -1)  # (t.., 5)
        # cells: for each x, state -> idx + disp[a][state][1:6]
        d = disp[:, sl[0], sl[1], sl[2], sl[3], sl[4], 1:]       # (A, t.., 5)
        cells = []
        for x in range(1, 19):   # live x only
            a = p[x]
            dd = np.take_along_axis(d, a[None, ..., None].astype(np.int64), axis=0)[0]   # (t..,5)
            c = idx + dd
            # clamp like the engine: cell in [0, 18]; wrap displacements (+-18/19) produce far cells: keep
            c = np.clip(c, 0, 18)
            cells.append(c.reshape(-1, 5))
        cells = np.unique(np.concatenate(cells, 0), axis=0)
        lo, hi = cells.min(0), cells.max(0) + 1        # upper corner +1
        vols.append(np.prod(hi - lo + 1))
        # exact distinct lines: cells + {0,1}^5
        corners = (cells[:, None, :] + np.array(list(itertools.product([0, 1], repeat=5)))[None]).reshape(-1, 5)
        exact.append(len(np.unique(corners, axis=0)))
        nst.append(np.prod(t))
    vols, exact = np.array(vols), np.array(exact)
    print(f"tile {tile}: x-lines/tile {np.prod(t)}, box lines median {np.median(vols):.0f} p90 {np.percentile(vols,90):.0f} max {vols.max()}, "
          f"exact distinct lines median {np.median(exact):.0f} p90 {np.percentile(exact,90):.0f}; "
          f"fit<=2560: box {np.mean(vols<=2560):.2f} exact {np.mean(exact<=2560):.2f}; "
          f"lines per x-line: box {np.median(vols)/np.prod(t):.1f} exact {np.median(exact)/np.prod(t):.1f} (scalar kernel: 32)")
for tile in [(1,1,1,1,1),(1,1,1,1,4),(1,1,2,2,2),(2,2,2,2,2),(1,2,4,2,4),(2,2,4,2,4),(1,4,4,4,4),(2,4,4,4,4),(4,4,4,4,4),(1,1,1,4,10),(1,1,4,4,5),(1,2,5,4,5)]:
    analyse(tile, 200)
