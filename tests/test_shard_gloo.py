"""CPU, world_size 2 over gloo: the N>1 host path -- block partition of the problems, no data-path collective, and the
final all-gather of per-problem result rows reproduces the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_motion_planning_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_rows(lo, hi):
    """Deterministic stand-in for the device-produced rows of problems lo..hi-1."""
    ids = np.arange(lo, hi)
    rng_rows = [np.random.default_rng(1000 + i).uniform(0, 1, 3) for i in ids]
    return torch.tensor(np.column_stack([ids, np.array(rng_rows).reshape(len(ids), 3)]) if len(ids) else np.zeros((0, 4)),
                        dtype=torch.float32)


def _worker(rank, world, port, n_problems, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(n_problems, rank, world)
        rows = _fake_rows(lo, hi)
        equal = n_problems % world == 0          # equal shards: the single-collective fast path the bench uses
        allrows = shard.gather_result_rows(rows, equal_shards=equal)
        q.put((rank, lo, hi, allrows.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_problems", [7, 8, 1])
def test_gather_matches_single_process(n_problems):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_problems, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_rows(0, n_problems).numpy()
    covered = []
    for rank, lo, hi, allrows in got:
        assert np.array_equal(allrows, want)        # every rank ends with the identical, id-ordered table
        covered += list(range(lo, hi))
    assert sorted(covered) == list(range(n_problems))   # shards are disjoint and complete
    assert shard.summarize(torch.from_numpy(want))["n_problems"] == n_problems


def test_partitions():
    for n in (0, 1, 5, 256, 1000):
        for w in (1, 2, 4, 8):
            b = [shard.shard_range(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    costs = np.array([1.0] * 100 + [3.5] * 100)            # maze graphs then kuka7 graphs
    b = shard.balanced_shards(costs, 4)
    assert b[0] == 0 and b[-1] == 200 and np.all(np.diff(b) >= 0)
    per = [costs[b[i]:b[i + 1]].sum() for i in range(4)]
    assert max(per) - min(per) <= 3.5 * 2
    assert np.array_equal(shard.balanced_shards(np.ones(8), 8), np.arange(9))


def test_summarize_search_rows():
    # (problem id, success, path cost, search checks, uncommitted speculative checks, explored nodes): eval_gnn.py:128-134
    rows = torch.tensor([[0, 1, 2.0, 100, 3, 20], [1, 0, 0.0, 300, 9, 80], [2, 1, 4.0, 200, 0, 40]], dtype=torch.float32)
    s_ = shard.summarize_search(rows)
    assert s_ == {"n_problems": 3, "n_success": 2, "collision_checks_mean": 200.0, "path_cost_mean": 3.0, "speculative_checks_mean": 4.0}
