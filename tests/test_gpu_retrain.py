"""N4 (runners/trial_runner.sh:33-60, runners/eval_metric.py:24-65): a persistent engine across autoresearch
trials.  `retrain()` / a pooled constructor must give exactly what a fresh engine gives for the new dynamics text, and the
second trial must start sweeping almost immediately (no context creation, no JIT compiles, no autotune)."""
import time

import numpy as np
import pytest

from dynamicprogramming_b200 import _ffi, engine, envs

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _edited(spec):
    """A 'trial': the same environment with an edited reward term (what the autoresearch agent changes)."""
    base = spec.cls

    class Edited(base):
        def _dynamics_cuda_src(self) -> str:
            src = super()._dynamics_cuda_src()
            assert "6.0f * gate" in src
            return src.replace("6.0f * gate", "4.5f * gate")

    return Edited


def _solve(eng, n_pi=2):
    for _ in range(n_pi):
        eng.policy_evaluation()
        eng.policy_improvement()
    return eng.download()


def test_retrain_equals_a_fresh_engine_bit_for_bit():
    spec = envs.REGISTRY["double_cartpole_swingup"]
    cfg = spec.config()
    cfg.max_eval_iter = 150
    Edited = _edited(spec)
    space, acts = spec.bins_space(8), spec.actions
    fresh = Edited(space, acts, cfg)
    v_ref, p_ref = _solve(fresh)
    fresh.close()

    eng = spec.cls(space, acts, cfg)            # trial 1: the original text
    _solve(eng)
    eng.__class__ = Edited                       # trial 2: same object, edited text
    assert eng.retrain() is True                 # the dynamics text changed: the builder was recompiled
    v2, p2 = _solve(eng)
    np.testing.assert_array_equal(p2, p_ref)
    np.testing.assert_array_equal(bits(v2), bits(v_ref))
    assert eng.retrain() is False                # same text again: table rebuilt, nothing compiled
    v3, p3 = _solve(eng)
    np.testing.assert_array_equal(p3, p_ref)
    np.testing.assert_array_equal(bits(v3), bits(v_ref))
    eng.close()


def test_pooled_constructor_reuses_the_engine_and_starts_sweeping_at_once():
    """bins 12 (the trial workload): the second trial's constructor + table build must take < 0.2 s and compile nothing
    but the edited table builder."""
    spec = envs.REGISTRY["double_cartpole_swingup"]
    cfg = spec.config()
    cfg.max_eval_iter, cfg.max_pi_iter = 100, 2
    Edited = _edited(spec)
    space, acts = spec.bins_space(12), spec.actions
    engine.keep_engines(True)
    try:
        t0 = time.perf_counter()
        a = spec.cls(space, acts, cfg)
        a.build_table()
        first = time.perf_counter() - t0
        a.run()                                   # ends with close(): the native engine is parked
        c0, h0 = (np.zeros(1, np.int64) for _ in range(2))
        _ffi.lib().pi_nvrtc_counters(c0.ctypes.data_as(_ffi.C.POINTER(_ffi.C.c_int64)), h0.ctypes.data_as(_ffi.C.POINTER(_ffi.C.c_int64)))
        t0 = time.perf_counter()
        b = Edited(space, acts, cfg)              # same grid / actions / config: takes over the parked engine
        b.build_table()
        second = time.perf_counter() - t0
        c1, h1 = (np.zeros(1, np.int64) for _ in range(2))
        _ffi.lib().pi_nvrtc_counters(c1.ctypes.data_as(_ffi.C.POINTER(_ffi.C.c_int64)), h1.ctypes.data_as(_ffi.C.POINTER(_ffi.C.c_int64)))
        assert (c1[0] - c0[0]) + (h1[0] - h0[0]) == 1, "only the edited table builder may be compiled (or fetched from the cache)"
        print(f"first trial ready in {first:.3f} s, second in {second:.3f} s")
        assert second < 0.2, (first, second)
        b.run()
        fresh_cfg = spec.config()
        fresh_cfg.max_eval_iter, fresh_cfg.max_pi_iter = 100, 2
    finally:
        engine.keep_engines(False)
    fresh = Edited(space, acts, fresh_cfg)
    fresh.run()
    np.testing.assert_array_equal(b.policy, fresh.policy)
    np.testing.assert_array_equal(bits(b.value_function), bits(fresh.value_function))
    assert b.total_eval_sweeps == fresh.total_eval_sweeps
