#!/usr/bin/env python3
"""
oracle/build_ref.py — compile the REFERENCE's own CUDA kernels into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under dynamicprogramming_b200/ may import,
link or execute anything produced here; only tests/, __graft_entry__.smoke()
and bench.py (--impl reference / cpu_baseline) use it, and only as the checker
or as the baseline being compared against.

What it does (in the build container, where /root/reference exists):
  1. imports the reference's src/cuda_policy_iteration.py and runners/*_cuda.py
     *from where they lie* with stub `cupy` / `matplotlib` modules on
     sys.modules (cupy is not installed in this image; matplotlib is only used
     for plotting),
  2. captures, per environment, the exact CUDA source string the reference
     hands to cp.RawModule (src/cuda_policy_iteration.py:288-289, :696-697,
     :1128-1129): `self._dynamics_cuda_src() + generic_kernels`,
  3. appends a ~20-line probe kernel (ours) that calls the reference's
     `step_dynamics` + `get_barycentric_Nd` device functions and dumps the
     indices / weights / reward / terminated they produce — the bit-exact
     oracle for the transition table,
  4. compiles with NVRTC (same default options cupy uses: no fast-math,
     fmad on, precise division) to an sm_100 cubin + PTX,
  5. writes oracle/_ref/<env>.cubin, <env>.ptx and manifest.json (launch
     metadata only: kernel names, D, default bins/actions/config).

No reference source text is written into the repository: the outputs are
binaries (git-ignored; they travel to the GPU box with the snapshot) and PTX
kept only for inspection (also git-ignored).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--arch sm_100a]
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"

ENVS = {
    # name: (module, class, D)
    "pendulum": ("runners.pendulum_cuda", "PendulumCuda", 2),
    "mountain_car": ("runners.mountain_car_cuda", "MountainCarCuda", 2),
    "continuous_mountain_car": ("runners.continuous_mountain_car_cuda", "ContinuousMountainCarCuda", 2),
    "cartpole": ("runners.cartpole_cuda", "CartPoleCuda", 4),
    "cartpole_swingup": ("runners.cartpole_swingup_cuda", "CartPoleSwingUpCuda", 4),
    "double_pendulum_swingup": ("runners.double_pendulum_swingup_cuda", "DoublePendulumSwingUpCuda", 4),
    "overhead_crane": ("runners.overhead_crane_cuda", "OverheadCraneCuda", 4),
    "double_cartpole": ("runners.double_cartpole_cuda", "DoubleCartPoleCuda", 6),
    "double_cartpole_swingup": ("runners.double_cartpole_swingup_cuda", "DoubleCartPoleSwingUpCuda", 6),
}

KERNEL_NAMES = {
    2: ("policy_eval_kernel", "policy_improve_kernel", "get_barycentric_2d"),
    4: ("policy_eval_kernel_4d", "policy_improve_kernel_4d", "get_barycentric_4d"),
    6: ("policy_eval_kernel_6d", "policy_improve_kernel_6d", "get_barycentric_6d"),
}


def _probe_kernel(D: int) -> str:
    """Our wrapper around the reference's device functions (step 3)."""
    C = 1 << D
    s_args = ", ".join(f"states[s * {D} + {d}]" for d in range(D))
    ns_decl = ", ".join(f"ns{d}" for d in range(D))
    ns_ptrs = ", ".join(f"&ns{d}" for d in range(D))
    ns_vals = ", ".join(f"ns{d}" for d in range(D))
    bary = KERNEL_NAMES[D][2]
    return f'''
extern "C" __global__ void oracle_probe_rows(
    const float* states, const float* actions, int a_idx,
    const float* b_low, const float* b_high, const int* g_shape, const int* strides,
    int n_states, int* out_idx, float* out_w, float* out_reward, unsigned char* out_term,
    float* out_next)
{{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    float {ns_decl}, reward; bool terminated;
    step_dynamics({s_args}, actions[a_idx], {ns_ptrs}, &reward, &terminated);
    int idxs[{C}]; float wgts[{C}];
    {bary}({ns_vals}, b_low, b_high, g_shape, strides, idxs, wgts);
    for (int c = 0; c < {C}; ++c) {{ out_idx[s * {C} + c] = idxs[c]; out_w[s * {C} + c] = wgts[c]; }}
    out_reward[s] = reward; out_term[s] = terminated ? 1 : 0;
    float nsv[{D}] = {{ {ns_vals} }};
    for (int d = 0; d < {D}; ++d) out_next[s * {D} + d] = nsv[d];
}}
'''


class _Nvrtc:
    """Minimal ctypes binding to the toolkit's libnvrtc (the same library the
    product links), so oracle and product use one compiler version."""

    def __init__(self) -> None:
        last = None
        for name in ("/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"):
            try:
                self.lib = ctypes.CDLL(name)
                break
            except OSError as exc:  # pragma: no cover
                last = exc
        else:  # pragma: no cover
            raise RuntimeError(f"libnvrtc not found: {last}")
        self.lib.nvrtcGetErrorString.restype = ctypes.c_char_p

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.nvrtcGetErrorString(rc).decode()}")

    def version(self) -> tuple[int, int]:
        a, b = ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.nvrtcVersion(ctypes.byref(a), ctypes.byref(b)), "nvrtcVersion")
        return a.value, b.value

    def compile(self, src: str, name: str, opts: list[str]) -> tuple[bytes, bytes, str]:
        prog = ctypes.c_void_p()
        self._check(
            self.lib.nvrtcCreateProgram(ctypes.byref(prog), src.encode(), name.encode(), 0, None, None),
            "nvrtcCreateProgram",
        )
        arr = (ctypes.c_char_p * len(opts))(*[o.encode() for o in opts])
        rc = self.lib.nvrtcCompileProgram(prog, len(opts), arr)
        n = ctypes.c_size_t()
        self.lib.nvrtcGetProgramLogSize(prog, ctypes.byref(n))
        log = ctypes.create_string_buffer(n.value)
        self.lib.nvrtcGetProgramLog(prog, log)
        if rc != 0:
            raise RuntimeError(f"NVRTC failed for {name}:\n{log.value.decode()}")
        self.lib.nvrtcGetPTXSize(prog, ctypes.byref(n))
        ptx = ctypes.create_string_buffer(n.value)
        self.lib.nvrtcGetPTX(prog, ptx)
        self.lib.nvrtcGetCUBINSize(prog, ctypes.byref(n))
        cubin = ctypes.create_string_buffer(n.value)
        self.lib.nvrtcGetCUBIN(prog, cubin)
        self.lib.nvrtcDestroyProgram(ctypes.byref(prog))
        return cubin.raw, ptx.raw.rstrip(b"\0"), log.value.decode()


def _install_stubs() -> dict:
    """Stub modules so the reference imports without cupy/matplotlib/gymnasium."""
    captured: dict = {}

    cp = types.ModuleType("cupy")

    class _RawModule:
        def __init__(self, code: str = "", **kw) -> None:
            captured["code"] = code

        def get_function(self, name: str):
            captured.setdefault("functions", []).append(name)
            return name

    cp.RawModule = _RawModule
    sys.modules["cupy"] = cp

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    for extra in ("matplotlib.patches", "matplotlib.animation", "matplotlib.colors", "matplotlib.cm"):
        sys.modules[extra] = types.ModuleType(extra)
    return captured


def capture_sources(reference: Path) -> dict:
    """Return {env: dict(source=..., D=..., meta=...)} from the reference tree."""
    import numpy as np

    captured = _install_stubs()
    sys.path.insert(0, str(reference))
    out = {}
    for env, (modname, clsname, D) in ENVS.items():
        mod = importlib.import_module(modname)
        cls = getattr(mod, clsname)
        inst = cls.__new__(cls)
        inst.n_states = 1
        if env == "overhead_crane":
            # train() default (runners/overhead_crane_cuda.py:248-262)
            inst.target_x = -2.5
        captured.clear()
        # the reference module binds `cp` at import time only if cupy imported
        core = importlib.import_module("src.cuda_policy_iteration")
        core.cp = sys.modules["cupy"]
        inst._compile_cuda_module()
        meta = {
            "D": D,
            "bins": {k: [float(v[0]), float(v[-1]), int(len(v))] for k, v in mod.BINS_SPACE.items()},
            "actions": [float(a) for a in np.asarray(mod.ACTION_SPACE, dtype=np.float32)],
            "eval_kernel": KERNEL_NAMES[D][0],
            "improve_kernel": KERNEL_NAMES[D][1],
            "probe_kernel": "oracle_probe_rows",
        }
        if env == "overhead_crane":
            meta["target_x"] = -2.5
        out[env] = {"source": captured["code"], "D": D, "meta": meta}
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--arch", default="sm_100a")
    ap.add_argument("--keep-ptx", action="store_true", default=True)
    args = ap.parse_args()

    reference = Path(args.reference)
    if not (reference / "src" / "cuda_policy_iteration.py").exists():
        print(f"[build_ref] {reference} not present — keeping prebuilt oracle/_ref/ as is")
        return 0

    OUT.mkdir(parents=True, exist_ok=True)
    nvrtc = _Nvrtc()
    sources = capture_sources(reference)
    manifest = {"nvrtc_version": list(nvrtc.version()), "arch": args.arch, "envs": {}}
    for env, item in sources.items():
        src = item["source"] + _probe_kernel(item["D"])
        # cupy.RawModule defaults: only the architecture flag (no fast-math).
        cubin, ptx, log = nvrtc.compile(src, f"{env}.cu", [f"--gpu-architecture={args.arch}"])
        (OUT / f"{env}.cubin").write_bytes(cubin)
        if args.keep_ptx:
            (OUT / f"{env}.ptx").write_bytes(ptx)
        manifest["envs"][env] = item["meta"]
        print(f"[build_ref] {env}: cubin {len(cubin)} B")
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
