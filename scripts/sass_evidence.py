#!/usr/bin/env python3
"""SASS evidence for the JIT-compiled (NVRTC) sweep kernels — no GPU needed.  Compiles the K5 configurations through the
library's own code path (pi_xline_compile_check with DPB200_DUMP_JIT), disassembles the cubins with cuobjdump and writes
profiles/r02_sass_<kernel>.txt: resource usage, opcode histogram, and the section of the hot loop that shows the
instructions the design claims (UBLKCP = cp.async.bulk TMA copies, SYNCS = mbarrier, LDS gathers from the staged V-planes,
FFMA chain; LDG.E.CONSTANT gathers with immediate offsets for gp_sweep; FFMA2 / FMUL2 and 128-bit window loads for xl_sweep)."""
import collections, ctypes as C, os, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
tmp = tempfile.mkdtemp()
os.environ["DPB200_CACHE"] = "off"; os.environ["DPB200_DUMP_JIT"] = tmp
from dynamicprogramming_b200 import _ffi
lib = _ffi.lib()
CASES = [("ps_sweep", b"plane:", "pi_plane_sweep.cu", "plane-staged sweep, K5 default = item mode (two states per thread: 224 consumer threads + producer warp, 64 slots, chunk 20, 2 CTAs/SM; pairs in packed f32x2 halves)", r"UBLKCP|UBLKPF|SYNCS|LDS|FFMA2|FMUL2"),
         ("ps_sweep_one", b"plane:0,0,2,2,0", "pi_plane_sweep.cu", "plane-staged sweep, one state per thread (67 slots, chunk 20, 2 CTAs/SM, scalar weight tree)", r"UBLKCP|SYNCS|LDS|FFMA"),
         ("gp_sweep", b"pair:128,8,2,8,1", "pi_pair_sweep.cu", "gather sweep, one state per thread, strides as immediates, gathers in groups of 8", r"LDG\.E\.CONSTANT|FFMA"),
         ("xl_sweep", b"2,0,4,8,2,1:1,1,1,2,10", "pi_xline_sweep.cu", "x-line sweep K = 2 (packed f32x2 math, vector window loads)", r"FFMA2|FMUL2|LDG\.E\.(64|128)|UBLKPF")]
for name, cfg, fname, what, pat in CASES:
    n = C.c_int64()
    _ffi.check(lib.pi_xline_compile_check(6, 20, cfg, C.byref(n)))
    cubin = Path(tmp) / (fname + ".cubin")
    res = subprocess.run(["cuobjdump", "-res-usage", str(cubin)], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
    lines = [re.sub(r"\s*/\*[0-9a-f]{4}\*/\s*", "", l.split("/* 0x")[0]).rstrip(" ;") for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        toks = [t for t in l.split() if not t.startswith("@")]
        if toks: ops[toks[0]] += 1
    out = [f"{name}: {what}", f"compiled by NVRTC for sm_100a through pi_xline_compile_check(6, 20, {cfg.decode()!r}); cubin {n.value} bytes; {len(lines)} SASS instructions",
           "", "== cuobjdump -res-usage"] + [l for l in res.splitlines() if "REG" in l or "Function" in l] + ["", "== opcode histogram (all instructions of the kernel)"]
    out += ["  %5d  %s" % (c, o) for o, c in ops.most_common(40)]
    hot = [i for i, l in enumerate(lines) if re.search(pat, l)]
    out += ["", f"== instructions matching /{pat}/: {len(hot)}; first 60 in program order"]
    out += ["  " + lines[i] for i in hot[:60]]
    if name == "ps_sweep":   # the pair path: from the first packed multiply to the last packed fma
        a = next(i for i, l in enumerate(lines) if "FMUL2" in l); b = max(i for i, l in enumerate(lines) if "FFMA2" in l)
        seg = collections.Counter(([t for t in l.split() if not t.startswith("@")] or ["?"])[0] for l in lines[a:b + 1])
        out += ["", f"== the pair path (instructions {a}..{b}: one warp backs up 64 states): opcode histogram"]
        out += ["  %5d  %s" % (c, o) for o, c in seg.most_common(14)]
    (ROOT / "profiles" / f"r02_sass_{name}.txt").write_text("\n".join(out) + "\n")
    print(name, len(lines), "instructions;", {k: v for k, v in ops.items() if re.match(r"UBLKCP|SYNCS|LDS$|LDS\.|FFMA|FMUL|LDG|UBLKPF", k)})
