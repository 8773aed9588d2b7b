"""
Run the reference's own runner scripts UNMODIFIED on the B200 engine.

    python -m dynamicprogramming_b200.compat /path/to/DynamicProgramming/runners/pendulum_cuda.py --bins 200 --retrain
    python -m dynamicprogramming_b200.compat runners/double_cartpole_swingup_cuda.py --bins 12 --episodes 3 --no-plot

Every `runners/*_cuda.py` of the reference does, at import time,

    import matplotlib.pyplot as plt                                      (e.g. runners/pendulum_cuda.py:33)
    from src.cuda_policy_iteration import CudaPolicyIteration2D, CudaPIConfig          (:36)

and the overhead crane additionally `import cupy as cp` inside its `_allocate_tensors_and_compile`
override to store a goal value through a boolean mask (runners/overhead_crane_cuda.py:193-206).
`install()` satisfies exactly those imports, before the script runs, by registering in `sys.modules`:

  * `src.cuda_policy_iteration` -> the engine classes of this package (same names, `GPU_AVAILABLE = True`);
  * `cupy` -> ONLY when the real cupy is not importable: the three names the crane's override touches
    (`asarray`, `bool_`, `get_default_memory_pool`).  Masks stay host arrays; the store itself is
    `DeviceArray.__setitem__` -> pi_set_values_buffer (include/dpb200.h).  Nothing is computed by it;
  * `matplotlib.pyplot` -> ONLY when matplotlib is not importable: a stub whose functions raise a clear
    error when called, so `--no-plot` runs work and plotting fails loudly instead of silently.

    python -m dynamicprogramming_b200.compat --serve      # one runner invocation per stdin line, engine kept (N4)

The script's own `train()`, `evaluate()`, argparse block, `--bins` handling and `save()/load()` run as
written; what changes is the class they subclass.  Nothing here touches oracle/.
"""
from __future__ import annotations

import importlib
import importlib.util
import runpy
import sys
import types
from pathlib import Path

from . import engine

_INSTALLED = False


def _importable(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def _engine_module() -> types.ModuleType:
    mod = types.ModuleType("src.cuda_policy_iteration")
    mod.__doc__ = "dynamicprogramming_b200 drop-in for src/cuda_policy_iteration.py"
    for name in ("CudaPIConfig", "CudaPolicyIteration2D", "CudaPolicyIteration4D", "CudaPolicyIteration6D"):
        setattr(mod, name, getattr(engine, name))
    mod.GPU_AVAILABLE = True   # src/cuda_policy_iteration.py:29-33
    mod.__all__ = ["CudaPIConfig", "CudaPolicyIteration2D", "CudaPolicyIteration4D", "CudaPolicyIteration6D", "GPU_AVAILABLE"]
    return mod


def _cupy_stub() -> types.ModuleType:
    import numpy as np

    mod = types.ModuleType("cupy")
    mod.__doc__ = "dynamicprogramming_b200.compat: the names reference subclasses use on engine buffers"
    mod.__dpb200_stub__ = True
    mod.bool_, mod.float32, mod.int32, mod.uint8 = np.bool_, np.float32, np.int32, np.uint8

    def asarray(a, dtype=None):
        return np.asarray(a, dtype=dtype)

    class _Pool:
        def free_all_blocks(self) -> None:   # the engine owns and frees its device memory (pi_destroy)
            return None

    mod.asarray = asarray
    mod.get_default_memory_pool = lambda: _Pool()

    def __getattr__(name):
        raise AttributeError(
            f"cupy.{name}: cupy is not installed; dynamicprogramming_b200.compat only provides asarray / bool_ / "
            "get_default_memory_pool for boolean-mask stores into engine buffers")

    mod.__getattr__ = __getattr__
    return mod


def _matplotlib_stub() -> tuple[types.ModuleType, types.ModuleType]:
    top = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    top.__dpb200_stub__ = plt.__dpb200_stub__ = True
    top.__path__ = []

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _missing(*a, **k):
            raise RuntimeError(f"matplotlib.pyplot.{name}: matplotlib is not installed — run the runner with --no-plot")

        return _missing

    plt.__getattr__ = __getattr__
    top.pyplot = plt
    top.use = lambda *a, **k: None
    return top, plt


def install() -> dict:
    """Register the drop-in modules; returns which were installed (idempotent)."""
    global _INSTALLED
    done = {"src.cuda_policy_iteration": True, "cupy_stub": False, "matplotlib_stub": False}
    pkg = sys.modules.get("src")
    if pkg is None or not getattr(pkg, "__dpb200__", False):
        pkg = types.ModuleType("src")
        pkg.__path__ = []
        pkg.__dpb200__ = True
        sys.modules["src"] = pkg
    mod = _engine_module()
    sys.modules["src.cuda_policy_iteration"] = mod
    pkg.cuda_policy_iteration = mod
    if "cupy" not in sys.modules and not _importable("cupy"):
        sys.modules["cupy"] = _cupy_stub()
    done["cupy_stub"] = bool(getattr(sys.modules.get("cupy"), "__dpb200_stub__", False))
    if "matplotlib" not in sys.modules and not _importable("matplotlib"):
        top, plt = _matplotlib_stub()
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = top, plt
    done["matplotlib_stub"] = bool(getattr(sys.modules.get("matplotlib"), "__dpb200_stub__", False))
    _INSTALLED = True
    return done


def uninstall() -> None:
    """Remove what install() registered (tests)."""
    global _INSTALLED
    for name in ("src.cuda_policy_iteration", "src"):
        m = sys.modules.get(name)
        if m is not None and (name != "src" or getattr(m, "__dpb200__", False)):
            del sys.modules[name]
    for name in ("cupy", "matplotlib.pyplot", "matplotlib"):
        if getattr(sys.modules.get(name), "__dpb200_stub__", False):
            del sys.modules[name]
    _INSTALLED = False


def load_runner(path: str | Path, run_name: str | None = None) -> dict:
    """Execute a reference runner file as a module (its `__main__` block does not run) and return its
    globals: `BINS_SPACE`, `ACTION_SPACE`, `train`, the `...Cuda` class, ..."""
    install()
    path = Path(path).resolve()
    return runpy.run_path(str(path), run_name=run_name or path.stem)


def run_runner(path: str | Path, argv: list[str]) -> None:
    """`python <runner> <argv...>` with the engine swapped in."""
    install()
    path = Path(path).resolve()
    old_argv = sys.argv
    sys.argv = [str(path), *argv]
    try:
        runpy.run_path(str(path), run_name="__main__")
    finally:
        sys.argv = old_argv


def serve(stream=None) -> int:
    """`--serve`: a long-lived trial loop.  Every line read from stdin is one runner invocation,
    e.g. `runners/double_cartpole_swingup_cuda.py --bins 12 --episodes 3 --no-plot` — what runners/trial_runner.sh:38-44
    passes to a fresh `python` per trial.  The script file is re-read for every line (so an edited dynamics / reward
    string takes effect), but the process, the CUDA context and the native engine stay: the runner's `train()` builds
    its object as always and the constructor takes over the parked engine of the previous trial (engine.keep_engines),
    recompiling only the table builder when the dynamics text changed.  One `[serve] ...` status line per trial."""
    import shlex
    import time
    import traceback

    engine.keep_engines(True)
    stream = sys.stdin if stream is None else stream
    n = 0
    for line in stream:
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if line in ("quit", "exit"):
            break
        argv = shlex.split(line)
        t0 = time.perf_counter()
        try:
            run_runner(argv[0], argv[1:])
            status = "ok"
        except SystemExit as ex:      # argparse / sys.exit inside the runner must not end the server
            status = f"exit {ex.code}"
        except Exception:             # noqa: BLE001 - a broken trial is reported, the loop goes on
            traceback.print_exc()
            status = "error"
        n += 1
        print(f"[serve] trial {n} {status} in {time.perf_counter() - t0:.3f} s: {line}", flush=True)
    engine.release_engines()
    return 0


def main(argv: list[str] | None = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    if argv[0] == "--serve":
        return serve()
    run_runner(argv[0], argv[1:])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
