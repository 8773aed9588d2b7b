"""World-size-2 test (gloo, CPU) of the sharding logic: contiguous state ranges,
needs-driven exchange plan, and the invariant that sharded Jacobi sweeps are
bit-identical to the unsharded ones.  The compute is the CPU oracle; the
partition / plan arithmetic is the package's (dynamicprogramming_b200.dist)."""
import os
import socket

import numpy as np
import pytest

from dynamicprogramming_b200 import dist as pdist


def test_partition_covers_everything_once():
    for n, w in [(10, 3), (64_000_000, 8), (7, 7), (2401, 2), (5, 8)]:
        edges = [pdist.shard_range(r, w, n) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        for idx in {0, n - 1, n // 2, n // 3}:
            r = pdist.owner_of(idx, w, n)
            lo, hi = edges[r]
            assert lo <= idx < hi


def test_plan_is_degenerate_all_gather_when_rows_reach_everywhere():
    n, w = 1000, 4
    base = np.arange(0, 900, 7)
    need = pdist.plan_exchange(base, span=99, rank=1, world=w, n_states=n)
    assert need[1] == (0, 0)
    assert need[0] == (0, 250) and need[2] == (500, 750)
    assert need[3][0] == 750 and need[3][1] == 896 + 99 + 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, env: str, bins: int, sweeps: int, out_dir: str) -> None:
    import torch
    import torch.distributed as td

    from oracle import cpu_oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    td.init_process_group("gloo", rank=rank, world_size=world)
    o = cpu_oracle.from_engine_env(env, bins=bins)
    N = o.n_states
    rng = np.random.default_rng(7)
    o.policy[:] = rng.integers(0, o.n_actions, N).astype(np.int32)
    lo, hi = pdist.shard_range(rank, world, N)
    span = int(o.strides.sum())
    # rows of MY states under the current policy -> what I need from each peer
    bases = []
    for a in range(o.n_actions):
        idx, w, r, t, _ = o.rows(a)
        sel = (o.policy[lo:hi] == a) & (t[lo:hi] == 0) & (o.terminal_mask[lo:hi] == 0)
        bases.append(idx[lo:hi, 0][sel])
    need = pdist.plan_exchange(np.concatenate(bases), span, rank, world, N)
    all_need = [None] * world
    td.all_gather_object(all_need, need)
    give = [all_need[r][rank] for r in range(world)]   # what peer r needs from me

    V = np.zeros(N, dtype=np.float32)
    V[o.terminal_mask.astype(bool)] = 0.0
    sentinel = np.float32(12345.0)
    for _ in range(sweeps):
        newV_full, _ = o.eval_sweep(V)                 # oracle computes all; keep only my slice
        nxt = np.full(N, sentinel, dtype=np.float32)   # anything I do not own or receive is poison
        nxt[lo:hi] = newV_full[lo:hi]
        reqs = []
        recv_bufs = {}
        for r in range(world):
            if r == rank:
                continue
            glo, ghi = give[r]
            if ghi > glo:
                reqs.append(td.isend(torch.from_numpy(nxt[glo:ghi].copy()), dst=r))
            nlo, nhi = need[r]
            if nhi > nlo:
                buf = torch.empty(nhi - nlo, dtype=torch.float32)
                recv_bufs[r] = (nlo, nhi, buf)
                reqs.append(td.irecv(buf, src=r))
        for q in reqs:
            q.wait()
        for r, (nlo, nhi, buf) in recv_bufs.items():
            nxt[nlo:nhi] = buf.numpy()
        V = nxt
    np.save(os.path.join(out_dir, f"slice_{rank}.npy"), V[lo:hi])
    np.save(os.path.join(out_dir, f"need_{rank}.npy"), np.array(need))
    td.destroy_process_group()


@pytest.mark.parametrize("env,bins", [("cartpole", 7), ("pendulum", 24)])
def test_sharded_sweeps_are_bit_identical_to_unsharded(tmp_path, env, bins):
    import torch.multiprocessing as mp

    from oracle import cpu_oracle

    world, sweeps = 2, 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, env, bins, sweeps, str(tmp_path)), nprocs=world, join=True)

    o = cpu_oracle.from_engine_env(env, bins=bins)
    rng = np.random.default_rng(7)
    o.policy[:] = rng.integers(0, o.n_actions, o.n_states).astype(np.int32)
    V = np.zeros(o.n_states, dtype=np.float32)
    for _ in range(sweeps):
        V, _ = o.eval_sweep(V)
    got = np.concatenate([np.load(tmp_path / f"slice_{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got.view(np.uint32), V.view(np.uint32))
    # the plan must be a strict subset of an all-gather for a local stencil
    need0 = np.load(tmp_path / "need_0.npy")
    lo1, hi1 = pdist.shard_range(1, world, o.n_states)
    assert need0[1][1] - need0[1][0] <= hi1 - lo1
