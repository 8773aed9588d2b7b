#!/usr/bin/env python3
"""Where the end-to-end step goes (K5): H2D policy + compaction + re-plan, 25 sweeps, D2H — wall-clock per call."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from dynamicprogramming_b200 import _ffi, envs
eng = envs.make("double_cartpole_swingup", bins=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
lib = _ffi.lib()
eng.build_table(); eng.sweeps(50); eng.policy_improvement()
n = eng.n_states
pol = torch.empty(n, dtype=torch.int32).pin_memory(); v = torch.empty(n, dtype=torch.float32).pin_memory()
pn, vn = pol.numpy(), v.numpy()
_ffi.check(lib.pi_copy_local_results(eng._engine, None, _ffi.ptr(pn)))
print(eng.eval_kernel_info()["kernel"][:50])
for rep in range(4):
    t0 = time.perf_counter(); _ffi.check(lib.pi_upload_policy_local(eng._engine, _ffi.ptr(pn))); t1 = time.perf_counter()
    eng.sweeps(25); t2 = time.perf_counter()
    _ffi.check(lib.pi_copy_local_results(eng._engine, _ffi.ptr(vn), None)); t3 = time.perf_counter()
    print("upload+compact+plan %.2f ms | 25 sweeps %.2f ms | D2H %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
