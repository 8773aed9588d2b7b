"""C-ABI library: loads, exports every symbol include/dpb200.h declares, fails
loudly without a device, compiles every plugin's dynamics (CPU only, no compute)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from dynamicprogramming_b200 import _ffi, envs
from dynamicprogramming_b200.engine import CudaPIConfig

ROOT = Path(__file__).resolve().parents[1]


def _declared_functions() -> list[str]:
    text = (ROOT / "include" / "dpb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(pi_[a-z_0-9]+)\s*\(", text)
    return sorted(set(n for n in names if n not in ("pi_log_fn",)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _ffi.lib()
    declared = _declared_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dpb200.h but not exported"
    # and the binding table covers the header (no stale / missing prototypes)
    assert sorted(_ffi.SIGNATURES) == declared
    assert lib.pi_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.PiConfig) == 24
    assert C.sizeof(_ffi.PiShard) == 136
    assert C.sizeof(_ffi.PiGrid) == 4 + 4 * 6 + 4 * 6 + 4 * 6 + 4 + 8 * 6  # incl. padding before the pointers
    assert C.sizeof(_ffi.PiStats) == 56


@pytest.mark.parametrize("name", sorted(envs.REGISTRY))
def test_every_plugin_compiles_against_the_table_builder(name):
    spec = envs.REGISTRY[name]
    inst = spec.cls.__new__(spec.cls)
    for k, v in spec.kwargs.items():
        setattr(inst, k, v)
    n = C.c_int64()
    rc = _ffi.lib().pi_compile_check(inst._dynamics_cuda_src().encode(), spec.cls.N_DIMS, C.byref(n))
    assert rc == 0, _ffi.lib().pi_last_error().decode()
    assert n.value > 1000


def test_compile_error_carries_the_nvrtc_log():
    bad = "__device__ void step_dynamics(float a, float b, float u, float* x, float* y, float* r, bool* t) { oops; }"
    rc = _ffi.lib().pi_compile_check(bad.encode(), 2, None)
    assert rc == _ffi.PI_ERR_COMPILE
    msg = _ffi.lib().pi_last_error().decode()
    assert "oops" in msg and "error" in msg.lower()
    with pytest.raises(_ffi.EngineError) as ei:
        _ffi.check(rc)
    assert ei.value.code == _ffi.PI_ERR_COMPILE


def test_wrong_signature_is_a_compile_error():
    # a 2-D plugin handed to a 4-D engine: the generated call does not match
    src = envs.REGISTRY["pendulum"].cls.__new__(envs.REGISTRY["pendulum"].cls)._dynamics_cuda_src()
    assert _ffi.lib().pi_compile_check(src.encode(), 4, None) == _ffi.PI_ERR_COMPILE


def test_invalid_arguments():
    lib = _ffi.lib()
    assert lib.pi_compile_check(None, 2, None) == _ffi.PI_ERR_INVALID
    assert lib.pi_compile_check(b"", 9, None) == _ffi.PI_ERR_INVALID
    assert lib.pi_n_states(None) == 0
    lib.pi_destroy(None)  # no-op


def test_no_device_means_loud_failure_not_cpu_fallback():
    lib = _ffi.lib()
    if lib.pi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_ffi.EngineError) as ei:
        envs.make("pendulum", bins=8)
    assert ei.value.code == _ffi.PI_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_lookup_without_a_device_fails_loudly_too():
    from dynamicprogramming_b200.utils import PolicyLookup

    lib = _ffi.lib()
    if lib.pi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_ffi.EngineError) as ei:
        PolicyLookup(np.zeros(4, np.int32), np.ones(2, np.float32), [0, 0], [1, 1], [2, 2])
    assert ei.value.code == _ffi.PI_ERR_NO_DEVICE


def test_jit_sweep_kernels_compile_for_sm_100a_without_a_gpu():
    """The x-line sweep is compiled per grid at run time; the same NVRTC path is checked here on
    synthetic grids (every K / LV / roll variant the engine can select)."""
    lib = _ffi.lib()
    n = __import__("ctypes").c_int64()
    for D, bins, cfg in [(6, 20, b"4,0,4,8,1,1:1,1,1,4,10"), (6, 20, b"2,0,4,8,2,1:1,1,1,2,10"), (6, 8, b"2,1,2,4,2,4:1,1,1,2,4"),
                         (4, 40, b"4,0,4,8,1,2:1,2,4"), (4, 12, b"2,1,2,8,2,2:1,3,6"), (3, 20, b"4,0,2,8,1,1:4,4")]:
        assert lib.pi_xline_compile_check(D, bins, cfg, __import__("ctypes").byref(n)) == _ffi.PI_OK, lib.pi_last_error()
        assert n.value > 10_000
    assert lib.pi_xline_compile_check(6, 20, b"pair:64,8", None) == _ffi.PI_OK
    # the plane-staged sweep (TMA-staged V-planes): 6-D and 4-D; one state per thread with scalar / packed weight tree,
    # and item mode (two states per thread)
    assert lib.pi_xline_compile_check(6, 20, b"plane:0,0,2,2,0", None) == _ffi.PI_OK, lib.pi_last_error()
    assert lib.pi_xline_compile_check(6, 20, b"plane:", None) == _ffi.PI_OK, lib.pi_last_error()
    assert lib.pi_xline_compile_check(6, 20, b"plane:0,0,2,2,2", None) == _ffi.PI_OK, lib.pi_last_error()
    assert lib.pi_xline_compile_check(4, 12, b"plane:0,0,2,1,2", None) == _ffi.PI_OK, lib.pi_last_error()
    assert lib.pi_xline_compile_check(4, 20, b"plane:0,0,2,1,1", None) == _ffi.PI_OK, lib.pi_last_error()
    assert lib.pi_xline_compile_check(6, 15, b"plane:0,0,2,2,0", None) == _ffi.PI_ERR_INVALID   # a 15 x 15 plane is not a multiple of 16 bytes
    assert lib.pi_xline_compile_check(4, 40, b"plane:0,0,2,2,0", None) == _ffi.PI_ERR_INVALID   # a 40 x 40 plane has more than 1024 states
    assert lib.pi_xline_compile_check(6, 21, b"4,0,4,8,1,1:1,1,1,1,3", None) == _ffi.PI_ERR_INVALID   # 21 % 4 != 0
    assert lib.pi_xline_compile_check(2, 20, b"4,0,4,8,1,1:4", None) == _ffi.PI_ERR_INVALID


def test_wrong_number_of_dimensions_asserts_like_the_reference():
    spec = envs.REGISTRY["pendulum"]
    bins = {"a": np.linspace(0, 1, 4, dtype=np.float32)}
    with pytest.raises(AssertionError):
        spec.cls(bins, spec.actions, CudaPIConfig())


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; the product path must not route through it."""
    pkg = ROOT / "dynamicprogramming_b200"
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.](cpu_oracle|ref_runner|pi_oracle|_ref|_build)", re.M)
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        assert not pat.search(path.read_text()), f"{path} references oracle/"


def test_nvrtc_disk_cache(tmp_path, monkeypatch):
    """run-time compiles are answered from a cubin cache keyed by the exact text."""
    import ctypes as C

    from dynamicprogramming_b200 import _ffi, envs

    lib = _ffi.lib()
    monkeypatch.setenv("DPB200_CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.delenv("DPB200_CACHE", raising=False)
    spec = envs.REGISTRY["pendulum"]
    src = spec.cls.__new__(spec.cls)._dynamics_cuda_src()

    def counters():
        a, b = C.c_int64(), C.c_int64()
        lib.pi_nvrtc_counters(C.byref(a), C.byref(b))
        return a.value, b.value

    def compile_(text):
        n = C.c_int64()
        _ffi.check(lib.pi_compile_check(text.encode(), 2, C.byref(n)))
        return n.value

    c0, h0 = counters()
    n1 = compile_(src)
    assert counters() == (c0 + 1, h0)
    files = list((tmp_path / "cache").glob("*.cubin"))
    assert len(files) == 1 and files[0].stat().st_size == n1
    assert compile_(src) == n1 and counters() == (c0 + 1, h0 + 1)            # hit: same bytes, no compile
    compile_(src + "\n// reward weights changed\n")                           # any edit of the text misses
    assert counters() == (c0 + 2, h0 + 1) and len(list((tmp_path / "cache").glob("*.cubin"))) == 2
    files[0].write_bytes(b"")                                                 # a truncated entry is ignored and replaced
    assert compile_(src) == n1 and counters() == (c0 + 3, h0 + 1) and files[0].stat().st_size == n1
    monkeypatch.setenv("DPB200_CACHE", "off")
    compile_(src)
    assert counters() == (c0 + 4, h0 + 1)
    with pytest.raises(_ffi.EngineError):                                     # errors are never cached
        compile_("this is not CUDA")
    monkeypatch.delenv("DPB200_CACHE")
    with pytest.raises(_ffi.EngineError):
        compile_("this is not CUDA")
