"""N2 — batched policy lookup (pi_lookup_actions / pi_lookup_create, include/dpb200.h) against
get_optimal_action of the reference (utils/barycentric.py:76-108): weights and indices are the
reference numba function's own outputs (tests/golden/barycentric_inference_golden.npz, generated
by importing utils/barycentric.py) and the oracle restatement of it."""
import numpy as np
import pytest

from dynamicprogramming_b200 import envs
from dynamicprogramming_b200.utils import PolicyLookup, get_optimal_action

pytestmark = pytest.mark.gpu


def _expected(weights, indices, policy, actions):
    """float32 sum in ascending corner order without contraction (what the kernel does) and the
    reference's own expression lambdas @ action_space[policy[flat_indices]]."""
    a = actions[policy[indices]].astype(np.float32)
    acc = np.zeros(len(weights), np.float32)
    for c in range(weights.shape[1]):
        acc = (acc + (weights[:, c] * a[:, c]).astype(np.float32)).astype(np.float32)
    ref = np.array([weights[i] @ a[i] for i in range(len(weights))], np.float32)
    return acc, ref


@pytest.mark.parametrize("D", [2, 4, 6])
def test_lookup_matches_the_reference_function_on_its_golden_points(D, golden_dir):
    g = np.load(golden_dir / "barycentric_inference_golden.npz")
    shape, lo, hi = g[f"d{D}_shape"], g[f"d{D}_lo"], g[f"d{D}_hi"]
    rng = np.random.default_rng(D)
    actions = np.linspace(-3.0, 5.0, 11, dtype=np.float32)
    policy = rng.integers(0, len(actions), int(np.prod(shape))).astype(np.int32)
    look = PolicyLookup(policy, actions, lo, hi, shape)
    got = look(g[f"d{D}_points"])
    seq, ref = _expected(g[f"d{D}_weights"], g[f"d{D}_indices"], policy, actions)
    np.testing.assert_array_equal(got.view(np.uint32), seq.view(np.uint32))        # bit-exact, same summation order
    np.testing.assert_allclose(got, ref, rtol=0, atol=4e-6)                        # BLAS dot order (tolerance 4e-6 abs)
    one = get_optimal_action(g[f"d{D}_points"][7], policy, actions, lo, hi, shape, g[f"d{D}_strides"], g[f"d{D}_corner_bits"])
    assert np.float32(one) == got[7]
    look.close()


@pytest.mark.parametrize("env,bins", [("pendulum", 31), ("cartpole", 9), ("double_cartpole_swingup", 6)])
def test_live_engine_and_loaded_policy_agree_with_the_oracle(env, bins, tmp_path):
    from oracle import cpu_oracle

    spec = envs.REGISTRY[env]
    cfg = spec.config()
    cfg.max_pi_iter, cfg.max_eval_iter = 2, 60
    eng = spec.make(bins=bins, config=cfg)
    eng.build_table()
    eng.policy_evaluation()
    eng.policy_improvement()
    D = eng.N_DIMS
    rng = np.random.default_rng(11)
    lo, hi = eng.bounds_low, eng.bounds_high
    pts = (lo + (hi - lo) * rng.uniform(-0.1, 1.1, (4000, D))).astype(np.float32)   # includes out-of-grid points (clamped)
    pts[:50] = eng.states_space[rng.integers(0, eng.n_states, 50)]                  # exactly on nodes
    pts[50] = lo
    pts[51] = hi
    live = eng.lookup_actions(pts)
    _, policy = eng.download()
    o = cpu_oracle.from_engine_env(env, bins=bins, config=cfg)
    w, idx = o.inference_weights(pts, eng.corner_bits)
    seq, ref = _expected(w, idx, policy, eng.action_space)
    np.testing.assert_array_equal(live.view(np.uint32), seq.view(np.uint32))
    # a query ON a grid node returns that node's action (up to the float32 rounding of the node coordinate)
    node_action = eng.action_space[policy[idx[:50][np.arange(50), np.argmax(w[:50], axis=1)]]]
    np.testing.assert_allclose(live[:50], node_action, rtol=0, atol=1e-4 * float(np.abs(eng.action_space).max()))
    np.testing.assert_allclose(live, ref, rtol=0, atol=4e-6 * float(np.abs(eng.action_space).max()))
    # saved-policy path: run() pulls the tensors and releases the device; save / load; same answers
    eng.policy, eng.value_function = policy, np.zeros(eng.n_states, np.float32)
    eng.save(tmp_path / "p")
    eng.close()
    loaded = type(eng).load(tmp_path / "p.npz")
    again = loaded.lookup_actions(pts)
    np.testing.assert_array_equal(again.view(np.uint32), live.view(np.uint32))


def test_lookup_argument_errors():
    with pytest.raises(ValueError):
        PolicyLookup(np.zeros(5, np.int32), np.ones(2, np.float32), [0, 0], [1, 1], [2, 2])
    from dynamicprogramming_b200._ffi import EngineError
    with pytest.raises(EngineError):
        PolicyLookup(np.full(4, 7, np.int32), np.ones(2, np.float32), [0, 0], [1, 1], [2, 2])   # action index out of range
    look = PolicyLookup(np.zeros(4, np.int32), np.ones(2, np.float32), [0, 0], [1, 1], [2, 2])
    assert look(np.zeros((0, 2), np.float32)).shape == (0,)
    with pytest.raises(ValueError):
        look(np.zeros((3, 5), np.float32))
    look.close()
