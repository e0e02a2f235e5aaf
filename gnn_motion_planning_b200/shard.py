"""Multi-GPU plumbing for the hot path: independent planning problems are block-partitioned over ranks (one process
per GPU); nothing is exchanged during graph build / forward / collision checks; the ONLY collective is the all-gather
of the per-problem result rows at the end (the reduction eval_gnn does over its `solutions` list,
eval_gnn.py:120-134).  torch.distributed is the transport: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_problems, rank, world):
    """Contiguous block partition: rank r owns problems [lo, hi); sizes differ by at most one."""
    base, rem = divmod(n_problems, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(costs, world):
    """Contiguous partition of problems into `world` blocks with near-equal total cost (e.g. edges per problem, so a
    mixed-environment sweep does not leave the maze ranks idle while the kuka ranks work).  Returns [world+1] bounds."""
    costs = np.asarray(costs, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(costs)])
    bounds = [0]
    for r in range(1, world):
        target = csum[-1] * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), len(costs))
        bounds.append(b)
    bounds.append(len(costs))
    return np.array(bounds, dtype=np.int64)


def gather_result_rows(rows, group=None, out=None, equal_shards=False):
    """All-gather the per-problem result rows [n_local, W] of every rank -> [n_total, W], ordered by rank (= by
    problem id for contiguous shards).  Shards may have different sizes: rows are padded to the largest shard.
    equal_shards=True (every rank holds the same number of rows, e.g. the weak-scaling bench): ONE collective straight into
    `out` ([world * n_local, W], allocated if None), no size exchange, no host synchronisation."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    if equal_shards:
        if out is None:
            out = rows.new_empty((world * rows.shape[0], rows.shape[1]))
        if rows.is_cuda:
            dist.all_gather_into_tensor(out, rows.contiguous(), group=group)
        else:
            dist.all_gather(list(out.view(world, rows.shape[0], rows.shape[1]).unbind(0)), rows.contiguous(), group=group)
        return out
    n_local = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in counts]
    n_max = max(counts)
    padded = rows.new_zeros((n_max, rows.shape[1]))
    padded[:rows.shape[0]] = rows
    out = rows.new_empty((world * n_max, rows.shape[1]))
    if rows.is_cuda:
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        chunks = list(out.view(world, n_max, rows.shape[1]).unbind(0))
        dist.all_gather(chunks, padded, group=group)
    out = out.view(world, n_max, rows.shape[1])
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def summarize(rows):
    """The reduction of eval_gnn.py:128-134 on gathered rows (problem id, E_g, #free edges, best logit)."""
    return {"n_problems": int(rows.shape[0]), "edges_total": float(rows[:, 1].sum()), "free_edges_total": float(rows[:, 2].sum()),
            "best_logit_mean": float(rows[:, 3].mean()) if rows.shape[0] else 0.0}


def summarize_search(rows):
    """The reduction of eval_gnn.py:128-134 on gathered planner rows (problem id, success, path cost, search checks, uncommitted
    speculative checks, explored nodes) -- `search.result_rows`."""
    ok = rows[:, 1] > 0
    n_ok = int(ok.sum())
    return {"n_problems": int(rows.shape[0]), "n_success": n_ok, "collision_checks_mean": float(rows[:, 3].mean()) if rows.shape[0] else 0.0,
            "path_cost_mean": float(rows[ok, 2].mean()) if n_ok else 0.0, "speculative_checks_mean": float(rows[:, 4].mean()) if rows.shape[0] else 0.0}
