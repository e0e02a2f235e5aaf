"""Batched collision checks on the GPU (C ABI: ``gmp_maze_*``, ``gmp_arm_*``).

Replaces the one-state / one-edge-at-a-time ``env._state_fp`` / ``env._edge_fp`` calls of the
reference planner loop (eval_gnn.py:215, smoother.py:209) with one launch per batch.
"""
import numpy as np
import torch

from . import _lib

_DT = {torch.float32: _lib.GMP_DTYPE_F32, torch.float64: _lib.GMP_DTYPE_F64}


def _dev_index(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


@torch.no_grad()
def maze_state_fp(states, maps, problem=None, want_counted=False):
    """states [n,2] f32|f64 cuda, maps [P,15,15] u8 cuda, problem [n] i32 cuda or None -> free u8 [n] (, counted u8 [n])."""
    for name, t in (("states", states), ("maps", maps)):
        _lib.require_cuda(t, name)
    lib = _lib.load()
    _lib.handle(_dev_index(states))
    if states.dtype not in _DT:
        raise TypeError("states must be float32 or float64")
    states = states.reshape(-1, 2).contiguous()
    maps = maps.to(torch.uint8).contiguous()
    n = states.shape[0]
    if problem is not None:
        problem = problem.to(device=states.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=states.device)
    counted = torch.empty(n, dtype=torch.uint8, device=states.device) if want_counted else None
    _lib.check(lib.gmp_maze_state_fp(_lib.ptr(states), _DT[states.dtype], _lib.ptr(maps), _lib.ptr(problem), n,
                                     _lib.ptr(free), _lib.ptr(counted), _lib.stream_ptr(states.device)))
    return (free, counted) if want_counted else free


@torch.no_grad()
def maze_edge_fp(a, b, maps, problem=None, want_checks=False):
    """a, b [n,2] f32|f64 cuda -> free u8 [n] (, n_checks i32 [n] = collision_check_count increments)."""
    for name, t in (("a", a), ("b", b), ("maps", maps)):
        _lib.require_cuda(t, name)
    lib = _lib.load()
    _lib.handle(_dev_index(a))
    if a.dtype not in _DT or b.dtype != a.dtype:
        raise TypeError("a and b must both be float32 or both float64")
    a = a.reshape(-1, 2).contiguous()
    b = b.reshape(-1, 2).contiguous()
    maps = maps.to(torch.uint8).contiguous()
    n = a.shape[0]
    if problem is not None:
        problem = problem.to(device=a.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=a.device)
    checks = torch.empty(n, dtype=torch.int32, device=a.device) if want_checks else None
    _lib.check(lib.gmp_maze_edge_fp(_lib.ptr(a), _lib.ptr(b), _DT[a.dtype], _lib.ptr(maps), _lib.ptr(problem), n,
                                    _lib.ptr(free), _lib.ptr(checks), _lib.stream_ptr(a.device)))
    return (free, checks) if want_checks else free


@torch.no_grad()
def maze_edge_fp_graph(v, edge_index, node_ptr_d, edge_ptr_d, maps, n_edges_total, problem_of_graph=None,
                       want_checks=False, free_out=None, checks_out=None):
    """Check every edge of a packed batch of graphs (endpoints gathered from v).

    v [N_total,2] f32 cuda; edge_index [2,>=E_total] i64 cuda (local ids); node_ptr_d / edge_ptr_d: DEVICE int32 [B+1].
    """
    lib = _lib.load()
    _lib.handle(_dev_index(v))
    B = node_ptr_d.numel() - 1
    free = free_out if free_out is not None else torch.empty(n_edges_total, dtype=torch.uint8, device=v.device)
    checks = checks_out
    if want_checks and checks is None:
        checks = torch.empty(n_edges_total, dtype=torch.int32, device=v.device)
    _lib.check(lib.gmp_maze_edge_fp_graph(_lib.ptr(v), _lib.ptr(edge_index), edge_index.stride(0), _lib.ptr(node_ptr_d),
                                          _lib.ptr(edge_ptr_d), _lib.ptr(problem_of_graph), B, n_edges_total,
                                          _lib.ptr(maps), _lib.ptr(free), _lib.ptr(checks), _lib.stream_ptr(v.device)))
    return (free, checks) if want_checks else free


@torch.no_grad()
def result_rows(logits, edge_free, edge_ptr_d, first_problem_id=0, out=None):
    """Per-problem result rows [B,4] = (problem id, E_g, #free edges, best logit) -- the all-gather payload."""
    lib = _lib.load()
    B = edge_ptr_d.numel() - 1
    rows = out if out is not None else torch.empty((B, 4), dtype=torch.float32, device=logits.device)
    _lib.check(lib.gmp_result_rows(_lib.ptr(logits), _lib.ptr(edge_free), _lib.ptr(edge_ptr_d), B, int(first_problem_id),
                                   _lib.ptr(rows), _lib.stream_ptr(logits.device)))
    return rows
