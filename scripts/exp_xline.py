#!/usr/bin/env python3
"""Experiment driver: x-line sweep configurations vs the scalar sweep on one environment.
   python scripts/exp_xline.py --env double_cartpole_swingup --bins 20 --configs "4,0,4,8,1:1,1,2,4,4;2,0,4,8,2:1,1,1,4,5" """
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

os.environ.setdefault("DPB200_XLINE", "off")   # the engine itself stays on the scalar sweep
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger

logger.remove()
from dynamicprogramming_b200 import _ffi, envs

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="double_cartpole_swingup")
ap.add_argument("--bins", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warm-sweeps", type=int, default=30)
ap.add_argument("--configs", default="4,0,4,8,1:1,1,2,4,4")
a = ap.parse_args()

eng = envs.make(a.env, bins=a.bins)
eng.build_table()
eng.policy_improvement()
eng.sweeps(a.warm_sweeps)
print("layout", eng.layout(), flush=True)
lib = _ffi.lib()
fn = lib.pi_debug_xline
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64),
               C.POINTER(C.c_double), C.POINTER(C.c_int32)]
first = True
for cfg in a.configs.split(";"):
    ms_new, ms_base = C.c_float(), C.c_float()
    mism = C.c_int64()
    wf = C.c_double()
    info = (C.c_int32 * 4)()
    rc = fn(eng._engine, cfg.encode(), a.iters, C.byref(ms_new), C.byref(ms_base) if first else None, C.byref(mism),
            C.byref(wf), info)
    if rc:
        print(cfg, "ERROR", lib.pi_last_error().decode()[:300], flush=True)
        continue
    if first:
        print(f"scalar sweep: {ms_base.value:.4f} ms  ({eng.n_states / ms_base.value / 1e6:.1f} G backups/s)", flush=True)
        first = False
    print(f"{cfg:>28s} regs {info[0]:3d} grid {info[1]:4d}x{info[2]:4d}: {ms_new.value:.4f} ms ({eng.n_states / ms_new.value / 1e6:6.1f} G/s) "
          f"mismatches={mism.value} window={wf.value:.4f}", flush=True)
eng.close()
