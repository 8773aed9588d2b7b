"""Pin the CPU oracle (oracle/pi_oracle.c) against the reference's own golden
artefacts and against known-answer cases.  CPU only."""
import hashlib

import numpy as np
import pytest

from dynamicprogramming_b200 import envs
from oracle import cpu_oracle


@pytest.fixture(scope="module")
def oracle_lib():
    return cpu_oracle.lib()


@pytest.mark.parametrize("env,min_agree,v_tol,max_outliers", [
    # continuous mountain car: the reference's golden V is reproduced to ~1e-6 relative
    ("continuous_mountain_car", 0.999, 1e-5, 0),
    # mountain car: cosf differs by 1 ulp at a few cell boundaries (3 boundary states)
    ("mountain_car", 0.995, 1e-5, 40),
])
def test_full_policy_iteration_reproduces_the_reference_golden_files(golden_dir, env, min_agree, v_tol, max_outliers):
    g = np.load(golden_dir / f"{env}_golden.npz")
    o = cpu_oracle.from_engine_env(env)  # 200 x 200, the runner's default config
    assert o.n_states == 40000
    np.testing.assert_array_equal(o.grid_shape, g["grid_shape"])
    np.testing.assert_array_equal(o.strides, g["strides"])
    np.testing.assert_array_equal(o.bounds_low, g["bounds_low"])
    np.testing.assert_array_equal(o.bounds_high, g["bounds_high"])
    np.testing.assert_array_equal(o.action_space, g["action_space"])
    digest = hashlib.sha256(np.ascontiguousarray(o.states_space).tobytes()).digest()
    assert digest == bytes(g["states_space_sha256"]), "states_space differs from the reference file"
    o.run()
    assert o.converged
    agree = float(np.mean(o.policy == g["policy"]))
    assert agree >= min_agree, agree
    V = g["value_function"]
    rel = np.abs(o.value_function - V) / np.maximum(np.abs(V), 1e-6)
    assert int((rel > v_tol).sum()) <= max_outliers, (float(rel.max()), int((rel > v_tol).sum()))


def test_golden_value_function_is_a_fixed_point_of_the_oracle_backup(golden_dir):
    """max |T_pi V - V| < theta on the reference's own converged V."""
    g = np.load(golden_dir / "continuous_mountain_car_golden.npz")
    o = cpu_oracle.from_engine_env("continuous_mountain_car")
    o.policy[:] = g["policy"]
    newV, delta = o.eval_sweep(np.ascontiguousarray(g["value_function"]))
    assert delta < 1.5e-4, delta


def test_inference_weights_match_the_reference_numba_function_bit_for_bit(golden_dir, oracle_lib):
    import ctypes as C

    g = np.load(golden_dir / "barycentric_inference_golden.npz")
    for D in (2, 4, 6):
        grid = cpu_oracle.OracleGrid()
        grid.n_dims = D
        for d in range(D):
            grid.shape[d] = int(g[f"d{D}_shape"][d])
            grid.strides[d] = int(g[f"d{D}_strides"][d])
            grid.lo[d] = float(g[f"d{D}_lo"][d])
            grid.hi[d] = float(g[f"d{D}_hi"][d])
        pts = np.ascontiguousarray(g[f"d{D}_points"])
        cb = np.ascontiguousarray(g[f"d{D}_corner_bits"])
        w = np.empty((len(pts), 1 << D), np.float32)
        idx = np.empty((len(pts), 1 << D), np.int32)
        oracle_lib.oracle_inference_weights(C.byref(grid), cpu_oracle._p(cb, C.c_int32), cpu_oracle._p(pts, C.c_float),
                                            C.c_int64(len(pts)), cpu_oracle._p(w, C.c_float),
                                            cpu_oracle._p(idx, C.c_int32))
        np.testing.assert_array_equal(idx, g[f"d{D}_indices"])
        np.testing.assert_array_equal(w.view(np.uint32), g[f"d{D}_weights"].view(np.uint32))


@pytest.mark.parametrize("env,bins", [("pendulum", 21), ("cartpole", 7), ("double_cartpole", 4)])
def test_rows_known_answers(env, bins):
    """Weights sum to one, indices stay inside the grid, corner order follows the
    reference (2-D: dim 0 is the high bit; N-D: bit d <-> dim d)."""
    o = cpu_oracle.from_engine_env(env, bins=bins)
    idx, w, r, t, nxt = o.rows(0)
    assert idx.min() >= 0 and idx.max() < o.n_states
    np.testing.assert_allclose(w.sum(axis=1), 1.0, atol=2e-6)
    assert (w >= 0).all()
    D = o.D
    st = o.strides.astype(np.int64)
    base = idx[:, 0]
    for c in range(1 << D):
        bits = [(c >> (1 - d)) & 1 for d in range(D)] if D == 2 else [(c >> d) & 1 for d in range(D)]
        np.testing.assert_array_equal(idx[:, c], base + int(np.dot(bits, st)))


def test_point_on_a_grid_node_has_weight_one_on_one_corner(oracle_lib):
    import ctypes as C

    o = cpu_oracle.from_engine_env("pendulum", bins=11)
    node = o.states_space[5 * 11 + 3]
    idxs = np.zeros(4, np.int32)
    wg = np.zeros(4, np.float32)
    oracle_lib.oracle_barycentric(C.byref(o.grid), cpu_oracle._p(node.copy(), C.c_float),
                                  cpu_oracle._p(idxs, C.c_int32), cpu_oracle._p(wg, C.c_float))
    k = int(np.argmax(wg))
    assert wg[k] == pytest.approx(1.0, abs=1e-5) and idxs[k] == 5 * 11 + 3


def test_gamma_zero_gives_the_reward_and_terminated_gives_the_reward():
    spec = envs.REGISTRY["mountain_car"]
    cfg = spec.config()
    cfg.gamma = 0.0
    o = cpu_oracle.from_engine_env("mountain_car", bins=31, config=cfg)
    o.policy[:] = 2
    V0 = np.random.default_rng(1).standard_normal(o.n_states).astype(np.float32)
    V0[o.terminal_mask.astype(bool)] = 0
    newV, _ = o.eval_sweep(V0)
    live = ~o.terminal_mask.astype(bool)
    np.testing.assert_array_equal(newV[live], np.float32(-1.0))          # V = r when gamma = 0
    np.testing.assert_array_equal(newV[~live], V0[~live])                 # absorbing states copy V
    # terminated transitions ignore V even with gamma > 0
    o2 = cpu_oracle.from_engine_env("mountain_car", bins=31)
    o2.policy[:] = 2
    _, _, r, t, _ = o2.rows(2)
    newV2, _ = o2.eval_sweep(V0)
    sel = (t == 1) & live
    assert sel.any()
    np.testing.assert_array_equal(newV2[sel], r[sel])


def test_lowest_action_index_wins_ties():
    """All-equal Q (gamma = 0, constant reward) must select action 0 everywhere."""
    spec = envs.REGISTRY["cartpole"]
    cfg = spec.config()
    cfg.gamma = 0.0
    o = cpu_oracle.from_engine_env("cartpole", bins=5, config=cfg)
    o.policy[:] = 1
    o.policy_improvement()
    live = ~o.terminal_mask.astype(bool)
    assert (o.policy[live] == 0).all()
    assert (o.policy[~live] == 1).all()  # terminal states are left untouched


def test_sync_interval_semantics():
    """Sweeps per evaluation are always 25k+1 (or max_eval_iter)."""
    o = cpu_oracle.from_engine_env("pendulum", bins=15)
    o.policy_evaluation()
    assert o.last_eval_sweeps % 25 == 1 or o.last_eval_sweeps == o.max_eval_iter
    spec = envs.REGISTRY["pendulum"]
    cfg = spec.config()
    cfg.max_eval_iter = 40
    o = cpu_oracle.from_engine_env("pendulum", bins=15, config=cfg)
    o.policy_evaluation()
    assert o.last_eval_sweeps == 40
