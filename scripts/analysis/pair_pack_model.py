#!/usr/bin/env python3
"""Plane-staged sweep, item mode: shared-memory wavefronts of the pair loads under the converged K5 policy for
(a) pairs sorted by (cell, in-plane index) and (b) pairs packed so that the 16 lanes of a half-warp hold distinct 64-bit bank
pairs (lane = (offset / 2) mod 16).  Inputs: scripts/data/k5_disp.npy, scripts/data/k5_policy16.npz.  CPU only."""
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
disp = np.load(str(ROOT / 'scripts' / 'data') + '/k5_disp.npy').astype(np.int16).reshape(9, 20, 20, 20, 20, 20, 6)
pol = np.load(str(ROOT / 'scripts' / 'data') + '/k5_policy16.npz')['policy'].reshape(20, 20, 20, 20, 20, 20)
rng = np.random.default_rng(0)
wf_sorted = []; wf_packed = []; hw_mixed = []; hw_mixed_s = []; hw_sorted = []; hw_packed = []; npairs = []; nsing = []
def wavefronts64(addrs):
    """one 64-bit load by up to 16 lanes (a half-warp): max number of distinct 8-byte words per bank pair"""
    a = np.unique(np.asarray(addrs) // 2)
    return np.bincount(a % 16, minlength=16).max() if a.size else 0
for _ in range(1500):
    pt = tuple(int(v) for v in rng.integers(0, 20, 4))
    a = pol[(slice(None), slice(None)) + pt]             # [x, xd]
    d = disp[(slice(None), slice(None)) + pt]            # (A, xd, 6)
    pairs = []                                           # (cell key, o)
    singles = 0
    for xd in range(20):
        x = 1
        while x <= 18:
            act = a[x, xd]; dd = d[act, xd]
            o = int(np.clip(xd + dd[1], 0, 18)) * 20 + int(np.clip(x + dd[0], 0, 18))
            if x + 1 <= 18 and a[x + 1, xd] == act and o % 2 == 0 and x + 1 + dd[0] <= 18:
                pairs.append((int(act), o)); x += 2
            else:
                singles += 1; x += 1
    npairs.append(len(pairs)); nsing.append(singles)
    pairs.sort()
    # (a) sorted: consecutive groups of 16 lanes
    w = 0
    for i in range(0, len(pairs), 16):
        w += wavefronts64([o for _, o in pairs[i:i + 16]])
    wf_sorted.append(w); hw_sorted.append((len(pairs) + 15) // 16)
    # (b) packed: per cell, half-warps = max multiplicity of the bank-pair class
    h = 0
    for act in set(k for k, _ in pairs):
        cls = np.array([(o // 2) % 16 for k, o in pairs if k == act])
        h += np.bincount(cls, minlength=16).max()
    wf_packed.append(h); hw_packed.append(h)
    # (c) packed across cells (slot pitch = 0 mod 32 words, so the bank of a float does not depend on its slot)
    cls = np.array([(o // 2) % 16 for k, o in pairs])
    hw_mixed.append(np.bincount(cls, minlength=16).max() if cls.size else 0)
print("pairs per plane %.1f, singles %.1f" % (np.mean(npairs), np.mean(nsing)))
print("sorted by (cell, index): %.2f half-warps, %.2f wavefronts per 64-bit pair load (ideal %.2f)" % (np.mean(hw_sorted), np.mean(wf_sorted), np.mean(npairs) / 16))
print("bank-aligned packing:    %.2f half-warps = wavefronts per pair load; lane efficiency %.2f" % (np.mean(hw_packed), np.mean(npairs) / 16 / np.mean(hw_packed)))
print("aligned across cells:    %.2f half-warps; lane efficiency %.2f" % (np.mean(hw_mixed), np.mean(npairs) / 16 / np.mean(hw_mixed)))
print("half-warps p50/p90/max", np.percentile(hw_packed, [50, 90, 100]))
