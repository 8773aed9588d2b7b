import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from loguru import logger
logger.remove()
from dynamicprogramming_b200 import envs, _ffi
import numpy as np
for env, bins in (("cartpole", 30), ("cartpole", 30), ("double_pendulum_swingup", 20)):
    eng = envs.make(env, bins=bins)
    eng.build_table()
    eng.policy_evaluation(); eng.policy_improvement()
    lib = _ffi.lib()
    v = np.empty(eng.n_states, np.float32); p = np.empty(eng.n_states, np.int32)
    t = time.perf_counter(); _ffi.check(lib.pi_copy_results(eng._engine, _ffi.ptr(v), _ffi.ptr(p))); t1 = time.perf_counter() - t
    t = time.perf_counter(); _ffi.check(lib.pi_copy_results(eng._engine, _ffi.ptr(v), _ffi.ptr(p))); t2 = time.perf_counter() - t
    t = time.perf_counter(); eng.close(); t3 = time.perf_counter() - t
    print(env, bins, "copy_results first", round(t1 * 1e3, 2), "second", round(t2 * 1e3, 2), "destroy", round(t3 * 1e3, 2), flush=True)
