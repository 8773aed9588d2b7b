"""Host-side mirror of the reference interface: grid metadata, saved-policy
format, lazy states_space, the runner's --bins path.  CPU only (no engine)."""
import hashlib

import numpy as np
import pytest

from dynamicprogramming_b200 import envs
from dynamicprogramming_b200.engine import (CudaPIConfig, CudaPolicyIteration2D, CudaPolicyIteration4D,
                                            CudaPolicyIteration6D)


def _shell(spec, bins=None):
    """An engine object with everything but the device part (what load() builds)."""
    inst = spec.cls.__new__(spec.cls)
    for k, v in spec.kwargs.items():
        setattr(inst, k, v)
    bs = spec.bins_space(bins)
    inst._axes = [np.ascontiguousarray(np.asarray(v).astype(np.float32)) for v in bs.values()]
    inst._axis_names = list(bs)
    inst.n_states = int(np.prod([len(a) for a in inst._axes]))
    inst._states_space = None
    inst._engine = None
    inst.action_space = np.ascontiguousarray(spec.actions, dtype=np.float32)
    inst.n_actions = len(inst.action_space)
    inst.config = spec.config()
    inst._precompute_grid_metadata()
    return inst


def test_config_defaults_match_the_reference():
    c = CudaPIConfig()
    assert (c.gamma, c.theta, c.max_eval_iter, c.max_pi_iter, c.log_interval) == (0.99, 1e-4, 10_000, 50, 100)


@pytest.mark.parametrize("env", ["mountain_car", "continuous_mountain_car"])
def test_grid_metadata_equals_the_reference_files(golden_dir, env):
    g = np.load(golden_dir / f"{env}_golden.npz")
    inst = _shell(envs.REGISTRY[env])
    for key in ("bounds_low", "bounds_high", "grid_shape", "strides", "corner_bits", "action_space"):
        got, want = getattr(inst, key), g[key]
        assert got.dtype == want.dtype and got.shape == want.shape, key
        np.testing.assert_array_equal(got, want, err_msg=key)
    ss = inst.states_space
    assert ss.dtype == np.float32 and ss.shape == tuple(g["states_space_shape"])
    assert hashlib.sha256(np.ascontiguousarray(ss).tobytes()).digest() == bytes(g["states_space_sha256"])


@pytest.mark.parametrize("cls,D", [(CudaPolicyIteration2D, 2), (CudaPolicyIteration4D, 4), (CudaPolicyIteration6D, 6)])
def test_strides_and_corner_bits(cls, D):
    spec = next(s for s in envs.REGISTRY.values() if s.cls.N_DIMS == D)
    inst = _shell(spec, bins=4)
    assert inst.corner_bits.shape == (1 << D, D) and inst.corner_bits.dtype == np.int32
    assert inst.strides[-1] == 1 and inst.strides[0] == 4 ** (D - 1)
    # row-major, dim 0 slowest, last dim fastest
    ss = inst.states_space
    assert ss[1, D - 1] != ss[0, D - 1] and (ss[1, : D - 1] == ss[0, : D - 1]).all()


def test_bins_path_follows_the_runner():
    spec = envs.REGISTRY["pendulum"]
    default = spec.bins_space()
    assert all(v.dtype == np.float32 and len(v) == 200 for v in default.values())
    np.testing.assert_array_equal(default["theta"], np.linspace(-np.pi, np.pi, 200, dtype=np.float32))
    other = spec.bins_space(64)
    lo, hi = default["theta"][0], default["theta"][-1]
    np.testing.assert_array_equal(other["theta"], np.linspace(lo, hi, 64, dtype=np.float32))
    assert list(other) == ["theta", "theta_dot"]


def test_saved_policy_format_round_trip(tmp_path, golden_dir):
    g = np.load(golden_dir / "mountain_car_golden.npz")
    inst = _shell(envs.REGISTRY["mountain_car"])
    inst.value_function = g["value_function"]
    inst.policy = g["policy"]
    inst.save(tmp_path / "sub" / "policy")               # suffix added, parents created
    path = tmp_path / "sub" / "policy.npz"
    assert path.exists()
    d = np.load(path)
    assert list(d.files) == list(g["file_keys"])          # same nine keys, same order
    assert [str(d[k].dtype) for k in d.files] == list(g["file_dtypes"])
    assert d["states_space"].shape == (40000, 2)
    loaded = envs.REGISTRY["mountain_car"].cls.load(path)  # no GPU needed
    np.testing.assert_array_equal(loaded.policy, g["policy"])
    np.testing.assert_array_equal(loaded.value_function, g["value_function"])
    assert loaded.n_states == 40000 and loaded.n_actions == 3
    assert isinstance(loaded.config, CudaPIConfig)


def test_crane_save_adds_target_x(tmp_path):
    spec = envs.REGISTRY["overhead_crane"]
    inst = _shell(spec, bins=4)
    inst.value_function = np.zeros(inst.n_states, np.float32)
    inst.policy = np.zeros(inst.n_states, np.int32)
    inst.save(tmp_path / "crane.npz")
    d = np.load(tmp_path / "crane.npz")
    assert float(d["target_x"]) == -2.5
    assert spec.cls.load(tmp_path / "crane.npz").target_x == -2.5
    assert "-2.500000f" in inst._dynamics_cuda_src()


def test_terminal_functions_match_spec_thresholds():
    inst = _shell(envs.REGISTRY["cartpole"])
    mask, value = inst._terminal_fn(inst.states_space)
    x, th = inst.states_space[:, 0], inst.states_space[:, 2]
    assert value == 0.0 and mask.dtype == bool
    assert mask[(np.abs(x) > 2.4)].all() and not mask[(np.abs(x) < 2.3) & (np.abs(th) < 0.2)].any()
    crane = _shell(envs.REGISTRY["overhead_crane"], bins=12)
    m, _ = crane._terminal_fn(crane.states_space)
    assert crane._goal_mask.sum() >= 0 and m.sum() >= crane._goal_mask.sum()


def test_duplicate_axis_nodes_are_rejected():
    spec = envs.REGISTRY["pendulum"]
    inst = spec.cls.__new__(spec.cls)
    inst._axes = [np.array([0, 0, 1], np.float32), np.array([0, 1, 2], np.float32)]
    inst._axis_names = ["a", "b"]
    inst.n_states, inst.n_actions = 9, 1
    with pytest.raises(ValueError):
        inst._precompute_grid_metadata()


def test_streamed_save_equals_eager_savez(tmp_path):
    """states_space streamed slab by slab into the archive == what np.savez writes."""
    import zipfile

    for env, bins in (("double_cartpole", 4), ("cartpole", 7), ("pendulum", 33)):
        spec = envs.REGISTRY[env]
        a, b = _shell(spec, bins=bins), _shell(spec, bins=bins)
        rng = np.random.default_rng(1)
        for inst in (a, b):
            inst.value_function = rng.random(inst.n_states, dtype=np.float32)
            inst.policy = rng.integers(0, inst.n_actions, inst.n_states).astype(np.int32)
            rng = np.random.default_rng(1)
        _ = a.states_space                      # a: materialised -> np.savez path;  b: streamed
        assert b._states_space is None
        a.save(tmp_path / f"{env}_eager")
        b.save(tmp_path / f"{env}_streamed")
        assert b._states_space is None          # still never materialised
        za, zb = zipfile.ZipFile(tmp_path / f"{env}_eager.npz"), zipfile.ZipFile(tmp_path / f"{env}_streamed.npz")
        assert za.namelist() == zb.namelist()
        for name in za.namelist():              # same .npy bytes entry by entry (header + data)
            assert za.read(name) == zb.read(name), name
            assert zb.getinfo(name).compress_type == zipfile.ZIP_STORED
        loaded = spec.cls.load(tmp_path / f"{env}_streamed")
        assert loaded._states_space is None     # lazy: read on first use
        np.testing.assert_array_equal(loaded.states_space, a.states_space)
        assert loaded.n_states == a.n_states


def test_slabwise_terminal_mask_equals_the_full_array_path(monkeypatch):
    from dynamicprogramming_b200 import engine

    for env, bins in (("cartpole", 9), ("double_cartpole_swingup", 5), ("overhead_crane", 8), ("mountain_car", 50)):
        inst = _shell(envs.REGISTRY[env], bins=bins)
        full, v_full = inst._terminal_mask_and_value()
        monkeypatch.setattr(engine, "_CHUNK_STATES", 10)      # force the slab-by-slab path
        slab, v_slab = _shell(envs.REGISTRY[env], bins=bins)._terminal_mask_and_value()
        monkeypatch.undo()
        np.testing.assert_array_equal(full, slab)
        assert v_full == v_slab and full.dtype == bool
        if env == "overhead_crane":     # the goal mask the plugin stashes on itself covers the whole grid again
            other = _shell(envs.REGISTRY[env], bins=bins)
            monkeypatch.setattr(engine, "_CHUNK_STATES", 10)
            other._terminal_mask_and_value()
            monkeypatch.undo()
            np.testing.assert_array_equal(other._goal_mask, inst._goal_mask)
            assert other._goal_mask.shape == (inst.n_states,)
