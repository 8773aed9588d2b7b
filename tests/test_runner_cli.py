"""The runners' command line (runners/*_cuda.py:__main__, README.md:327-345): flags, --bins
semantics, train-or-load, saved-policy format."""
import numpy as np
import pytest

from dynamicprogramming_b200 import envs, runners


def test_every_reference_runner_has_a_cli_with_the_shared_flags():
    assert set(runners.RUNNERS) == {"pendulum_cuda", "mountain_car_cuda", "continuous_mountain_car_cuda", "cartpole_cuda",
                                    "cartpole_swingup_cuda", "double_pendulum_swingup_cuda", "overhead_crane_cuda",
                                    "double_cartpole_cuda", "double_cartpole_swingup_cuda"}
    for name, env in runners.RUNNERS.items():
        a = runners.build_parser(name).parse_args([])
        assert (a.episodes, a.steps, a.seed) == (5, 1000, 42) and not a.retrain and not a.no_plot and a.random is None
        assert a.bins == envs.REGISTRY[env].default_bins
        assert str(a.save_path) == f"results/{name}_policy.npz"
        b = runners.build_parser(name).parse_args(["--bins", "12", "--retrain", "--no-plot", "--random", "--save-path", "x.npz"])
        assert b.bins == 12 and b.retrain and b.no_plot and b.random == 5 and str(b.save_path) == "x.npz"
    c = runners.build_parser("overhead_crane_cuda").parse_args(["--target-x", "1.5", "--start-x", "0"])
    assert c.target_x == 1.5 and c.start_x == 0.0


def test_unknown_runner_and_help():
    assert runners.main([]) == 2
    assert runners.main(["not_a_runner"]) == 2
    assert runners.main(["--help"]) == 0


def test_load_path_needs_no_gpu(tmp_path, capsys, golden_dir):
    """--save-path exists and no --retrain => Cls.load(), exactly like the reference (pendulum_cuda.py:298-303);
    runs on the CPU box: the golden mountain-car policy is loaded and summarised without an engine."""
    g = np.load(golden_dir / "mountain_car_golden.npz")
    path = tmp_path / "mc.npz"
    data = {k: g[k] for k in g.files}
    if "states_space" not in data:   # the fixture is a reduced copy of runners/results/*.npz
        axes = [np.linspace(data["bounds_low"][d], data["bounds_high"][d], int(data["grid_shape"][d]), dtype=np.float32) for d in range(2)]
        data["states_space"] = np.column_stack([m.ravel() for m in np.meshgrid(*axes, indexing="ij")]).astype(np.float32)
    np.savez(path, **data)
    rc = runners.main(["mountain_car_cuda", "--save-path", str(path), "--episodes", "0", "--no-plot"])
    out = capsys.readouterr().out
    assert rc == 0 and "Loading existing policy" in out and "policy histogram" in out


@pytest.mark.gpu
def test_train_then_load_round_trip(tmp_path, capsys):
    path = tmp_path / "results" / "cartpole_cuda_policy.npz"
    assert runners.main(["cartpole_cuda", "--bins", "8", "--save-path", str(path), "--episodes", "3", "--no-plot"]) == 0
    out = capsys.readouterr().out
    assert "Training new policy" in out and out.count("-> action") == 3
    z = np.load(path)
    assert set(z.files) == {"value_function", "policy", "bounds_low", "bounds_high", "grid_shape", "strides", "corner_bits",
                            "action_space", "states_space"}
    assert z["policy"].dtype == np.int32 and z["value_function"].dtype == np.float32 and z["states_space"].shape == (8 ** 4, 4)
    assert runners.main(["cartpole_cuda", "--bins", "8", "--save-path", str(path), "--episodes", "2"]) == 0
    out = capsys.readouterr().out
    assert "Loading existing policy" in out and out.count("-> action") == 2
