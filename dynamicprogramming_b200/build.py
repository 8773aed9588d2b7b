"""
Build libdpb200.so in-tree with nvcc for sm_100a (no torch, no JIT cache):

    python -m dynamicprogramming_b200.build

The shared object lands next to the package (dynamicprogramming_b200/libdpb200.so)
so it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libdpb200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found; the B200 engine cannot be built")
    return nvcc


def sources() -> list[Path]:
    return [CSRC / "dpb200.cu"]


def needs_rebuild() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG.parent / "include" / "dpb200.h")
    return any(d.stat().st_mtime > t for d in deps)


def stringify_jit_sources() -> None:
    """csrc/*_src.cuh (kernels compiled at run time by NVRTC) -> csrc/*_src.inc raw-string
    literals that dpb200.cu #includes."""
    for src in CSRC.glob("*_src.cuh"):
        inc = src.with_suffix(".inc")
        text = 'R"XLSRC(\n' + src.read_text() + '\n)XLSRC"\n'
        if not inc.exists() or inc.read_text() != text:
            inc.write_text(text)


def build(force: bool = False, verbose: bool = False) -> Path:
    stringify_jit_sources()
    if not force and not needs_rebuild():
        return LIB
    cuda_lib = Path(_nvcc()).resolve().parents[1] / "lib64"
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(s) for s in sources()]
    cmd += ["-o", str(LIB), f"-L{cuda_lib}", "-lnvrtc", "-ldl", "-Xlinker", f"-rpath={cuda_lib}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
