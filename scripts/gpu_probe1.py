#!/usr/bin/env python3
"""First GPU probe: parity of the engine against the reference's own kernels
(oracle/_ref cubins) + first timings.  Development aid, not part of the suite."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from loguru import logger  # noqa: E402

logger.remove()
logger.add(sys.stderr, level="WARNING")

import torch  # noqa: E402

from dynamicprogramming_b200 import envs  # noqa: E402
from dynamicprogramming_b200.engine import CudaPIConfig  # noqa: E402
from oracle import ref_runner  # noqa: E402

OUT = {}


def compare_env(name: str, bins: int) -> dict:
    res = {"bins": bins}
    spec = envs.REGISTRY[name]
    cfg = spec.config()
    eng = spec.make(bins=bins)
    ref = ref_runner.from_engine_env(name, bins=bins)
    N, A, D = eng.n_states, eng.n_actions, eng.N_DIMS
    res.update(N=N, A=A, D=D)
    eng.build_table()
    term = eng._terminal_mask_host.astype(bool)

    # 1. rows vs reference probe
    bad_idx = bad_w = bad_r = bad_t = 0
    for a in range(A):
        idx, w, r, t = eng.expand_rows(a)
        ridx, rw, rr, rt, _ = ref.probe_rows(a)
        live = ~term
        tt = (t == 1)
        bad_t += int(np.sum(tt[live] != (rt[live] == 1)))
        ok = live & ~tt & (rt == 0)
        bad_idx += int(np.sum(idx[ok] != ridx[ok]))
        bad_w += int(np.sum(w[ok].view(np.uint32) != rw[ok].view(np.uint32)))
        bad_r += int(np.sum(r[live].view(np.uint32) != rr[live].view(np.uint32)))
    res["rows"] = dict(bad_idx=bad_idx, bad_w=bad_w, bad_reward=bad_r, bad_term=bad_t)

    # 2. one sweep from a random V under constant policies, 3. improvement from the same V
    rng = np.random.default_rng(0)
    V0 = rng.standard_normal(N).astype(np.float32) * 10
    sweep_bad = {}
    Q_ref = np.empty((A, N), dtype=np.float32)
    for a in range(A):
        pol = np.full(N, a, dtype=np.int32)
        eng.upload_policy(pol)
        eng.upload_values(V0)
        eng.sweeps(1)
        v_mine, _ = eng.download()
        ref.d_policy.copy_(torch.from_numpy(pol).cuda())
        ref.d_value_function.copy_(torch.from_numpy(V0).cuda())
        v_ref = ref.sweep_once()
        Q_ref[a] = v_ref
        nb = int(np.sum(v_mine.view(np.uint32) != v_ref.view(np.uint32)))
        if nb:
            sweep_bad[a] = nb
    res["sweep_bitdiff_by_action"] = sweep_bad

    eng.upload_policy(np.zeros(N, dtype=np.int32))
    eng.upload_values(V0)
    eng.policy_improvement()
    _, p_mine = eng.download()
    ref.d_policy.zero_()
    ref.d_value_function.copy_(torch.from_numpy(V0).cuda())
    ref.improve_launch()
    p_ref = ref.d_policy.cpu().numpy()
    res["improve_diff"] = int(np.sum(p_mine != p_ref))
    # reference self-consistency: argmax of its eval-kernel Q vs its improve kernel
    best = np.zeros(N, dtype=np.int32)
    mq = np.full(N, -1.0e30, dtype=np.float32)
    for a in range(A):
        better = Q_ref[a] > mq
        mq = np.where(better, Q_ref[a], mq)
        best = np.where(better, a, best)
    best[term] = 0
    res["ref_self_inconsistent"] = int(np.sum(best != p_ref))
    eng.close()
    return res


def full_run(name: str, bins: int, max_pi=None) -> dict:
    spec = envs.REGISTRY[name]
    cfg = spec.config()
    if max_pi:
        cfg.max_pi_iter = max_pi
    eng = spec.make(bins=bins, config=cfg)
    torch.cuda.synchronize()
    t0 = time.time()
    eng.run()
    t_mine = time.time() - t0
    ref = ref_runner.from_engine_env(name, bins=bins, config=cfg)
    torch.cuda.synchronize()
    t0 = time.time()
    ref.run()
    torch.cuda.synchronize()
    t_ref = time.time() - t0
    V, P = eng.value_function, eng.policy
    out = dict(bins=bins, N=eng.n_states, pi_iters=(eng.pi_iterations, ref.pi_iterations),
               sweeps=(eng.total_eval_sweeps, ref.total_sweeps),
               policy_diff=int(np.sum(P != ref.policy)),
               v_bitdiff=int(np.sum(V.view(np.uint32) != ref.value_function.view(np.uint32))),
               v_maxrel=float(np.max(np.abs(V - ref.value_function) / np.maximum(np.abs(ref.value_function), 1e-6))),
               t_mine_s=t_mine, t_ref_s=t_ref, stats=eng.stats)
    return out


def timing(name: str, bins: int, n_sweeps: int = 50) -> dict:
    spec = envs.REGISTRY[name]
    eng = spec.make(bins=bins)
    t0 = time.time()
    eng.build_table()
    torch.cuda.synchronize()
    out = dict(bins=bins, N=eng.n_states, A=eng.n_actions, build_wall_s=time.time() - t0)
    eng.sweeps(5)
    _, ms = eng.sweeps(n_sweeps)
    out["sweep_ms"] = ms / n_sweeps
    out["backups_per_s"] = eng.n_states / (ms / n_sweeps * 1e-3)
    eng.policy_improvement()
    st = eng.engine_stats()
    out["stats"] = st
    eng.close()
    return out


def ref_timing(name: str, bins: int, n_sweeps: int = 10) -> dict:
    ref = ref_runner.from_engine_env(name, bins=bins)
    for _ in range(2):
        ref.eval_launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_sweeps):
        ref.eval_launch()
        ref._max_abs_diff()
        ref.d_value_function, ref.d_new_value_function = ref.d_new_value_function, ref.d_value_function
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_sweeps
    e0.record()
    ref.improve_launch()
    e1.record()
    torch.cuda.synchronize()
    return dict(bins=bins, N=ref.n_states, sweep_ms=ms, improve_ms=e0.elapsed_time(e1))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    small = {"pendulum": 41, "mountain_car": 50, "continuous_mountain_car": 50, "cartpole": 9,
             "cartpole_swingup": 9, "double_pendulum_swingup": 9, "overhead_crane": 9,
             "double_cartpole": 5, "double_cartpole_swingup": 6}
    OUT["compare"] = {}
    for name, b in small.items():
        try:
            OUT["compare"][name] = compare_env(name, b)
        except Exception as exc:  # noqa: BLE001
            OUT["compare"][name] = {"error": repr(exc)}
        print(name, json.dumps(OUT["compare"][name]), flush=True)

    OUT["full"] = {}
    for name, b, mp in [("mountain_car", 200, None), ("continuous_mountain_car", 200, None), ("pendulum", 200, None),
                        ("cartpole", 20, None), ("double_pendulum_swingup", 12, 10), ("double_cartpole_swingup", 8, 6)]:
        try:
            OUT["full"][name] = full_run(name, b, mp)
        except Exception as exc:  # noqa: BLE001
            OUT["full"][name] = {"error": repr(exc)}
        print("FULL", name, json.dumps(OUT["full"][name]), flush=True)

    OUT["timing"] = {}
    for name, b in [("cartpole", 30), ("double_pendulum_swingup", 50), ("double_cartpole_swingup", 12),
                    ("double_cartpole_swingup", 20)]:
        key = f"{name}@{b}"
        try:
            OUT["timing"][key] = timing(name, b)
            OUT["timing"][key]["ref"] = ref_timing(name, b)
        except Exception as exc:  # noqa: BLE001
            OUT["timing"][key] = {"error": repr(exc)}
        print("TIME", key, json.dumps(OUT["timing"][key]), flush=True)

    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "probe1.json").write_text(json.dumps(OUT, indent=1))
