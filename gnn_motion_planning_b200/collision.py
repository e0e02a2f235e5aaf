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
    if maps.dtype != torch.uint8 or not maps.is_contiguous():
        raise TypeError("maps must be a contiguous uint8 tensor [P,15,15] (the reference's maze_files maps are float64: convert "
                        "once with (maps != 0).to(torch.uint8))")
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


@torch.no_grad()
def maze3_state_fp(states, maps, problem=None):
    """3-D stick maze states [n,3] f32|f64 cuda -> (free u8 [n], n_checks i32 [n], k i32 [n])   (maze_env.py:279-291)."""
    _lib.require_cuda(states, "states")
    if states.dtype not in _DT:
        raise TypeError("states must be float32 or float64")
    states = states.reshape(-1, 3).contiguous()
    maps = maps.to(torch.uint8).contiguous()
    n = states.shape[0]
    if problem is not None:
        problem = problem.to(device=states.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=states.device)
    checks = torch.empty(n, dtype=torch.int32, device=states.device)
    k = torch.empty(n, dtype=torch.int32, device=states.device)
    _lib.check(_lib.load().gmp_maze3_state_fp(_lib.ptr(states), _DT[states.dtype], _lib.ptr(maps), _lib.ptr(problem), n, _lib.ptr(free),
                                              _lib.ptr(checks), _lib.ptr(k), _lib.stream_ptr(states.device)))
    return free, checks, k


@torch.no_grad()
def maze3_edge_fp(a, b, maps, problem=None):
    """3-D stick maze edges a, b [n,3] f32|f64 cuda -> (free u8 [n], n_checks i32 [n], k i32 [n])   (maze_env.py:316-347)."""
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    if a.dtype not in _DT or b.dtype != a.dtype:
        raise TypeError("a and b must both be float32 or both float64")
    a, b = a.reshape(-1, 3).contiguous(), b.reshape(-1, 3).contiguous()
    maps = maps.to(torch.uint8).contiguous()
    n = a.shape[0]
    if problem is not None:
        problem = problem.to(device=a.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=a.device)
    checks = torch.empty(n, dtype=torch.int32, device=a.device)
    k = torch.empty(n, dtype=torch.int32, device=a.device)
    _lib.check(_lib.load().gmp_maze3_edge_fp(_lib.ptr(a), _lib.ptr(b), _DT[a.dtype], _lib.ptr(maps), _lib.ptr(problem), n, _lib.ptr(free),
                                             _lib.ptr(checks), _lib.ptr(k), _lib.stream_ptr(a.device)))
    return free, checks, k


@torch.no_grad()
def maze_sample_points(maps, problem_of_slot, stream_of_slot, n_points, seed, first_draw=None, cap_collided=None):
    """Batched ``env.sample_n_points(n_points, need_negative=True)`` (maze_env.py:85-100) with the counter-based device RNG
    (``gmp_maze_sample_points``): -> (free [S,n,2] f64, collided [S,cap,2] f64, n_collided [S] i32, n_draws [S] i64), all CUDA.
    Slot s continues stream ``stream_of_slot[s]`` at draw ``first_draw[s]`` (0 if None)."""
    _lib.require_cuda(maps, "maps")
    lib = _lib.load()
    dev = maps.device
    prob = torch.as_tensor(problem_of_slot).to(device=dev, dtype=torch.int32).contiguous()
    strm = torch.as_tensor(stream_of_slot).to(device=dev, dtype=torch.int64).contiguous()
    fd = None if first_draw is None else torch.as_tensor(first_draw).to(device=dev, dtype=torch.int64).contiguous()
    S = prob.numel()
    cap = int(cap_collided if cap_collided is not None else 8 * n_points)
    free = torch.empty((S, n_points, 2), dtype=torch.float64, device=dev)
    coll = torch.empty((S, max(cap, 1), 2), dtype=torch.float64, device=dev)
    n_coll = torch.empty(S, dtype=torch.int32, device=dev)
    n_draws = torch.empty(S, dtype=torch.int64, device=dev)
    _lib.check(lib.gmp_maze_sample_points(_lib.ptr(maps.to(torch.uint8).contiguous()), _lib.ptr(prob), _lib.ptr(strm), _lib.ptr(fd), S,
                                          int(n_points), cap, int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(free), _lib.ptr(coll),
                                          _lib.ptr(n_coll), _lib.ptr(n_draws), _lib.stream_ptr(dev)))
    return free, coll, n_coll, n_draws


# ------------------------------------------------------------------------------------------------ arms
ARM_KUKA7, ARM_KUKA14, ARM_KUKA13, ARM_UR5, ARM_SNAKE7 = 0, 1, 2, 3, 4


def arm_model_info(model):
    """-> (dof, lower [dof] f64, upper [dof] f64): KukaEnv.pose_range (kuka_env.py:57-60)."""
    import ctypes
    lib = _lib.load()
    dof = ctypes.c_int32(0)
    lo, hi = np.zeros(14), np.zeros(14)
    _lib.check(lib.gmp_arm_model_info(int(model), ctypes.addressof(dof), lo.ctypes.data, hi.ctypes.data))
    return dof.value, lo[:dof.value].copy(), hi[:dof.value].copy()


def pack_boxes(problems_obstacles, device):
    """list (per problem) of [(halfExtents[3], basePosition[3]), ...] -> (boxes [O_total,6] f64 cuda, box_ptr [P+1] i32 cuda)."""
    rows, ptr = [], [0]
    for obs in problems_obstacles:
        for h, p in obs:   # ur5s_6_3000.pkl has ragged entries such as [0.01, 0.01, array([0.84])]
            rows.append(np.array([float(np.ravel(x)[0]) for x in list(h) + list(p)], np.float64))
        ptr.append(len(rows))
    boxes = torch.from_numpy(np.array(rows, np.float64).reshape(-1, 6)).to(device)
    return boxes, torch.from_numpy(np.array(ptr, np.int32)).to(device)


def snake_obstacles(maps):
    """SnakeEnv.create_maze (snake_env.py:63-71): half extents (0.7, 0.7, 1) at (1.4 i - 10.5, 1.4 j - 10.5, 0) for every
    occupied cell map[i, j], reference loop order (j outer, i inner) -> list (per problem) of (half, pos) pairs."""
    out = []
    for m in np.asarray(maps):
        out.append([((0.7, 0.7, 1.0), (1.4 * i - 10.5, 1.4 * j - 10.5, 0.0)) for j in range(m.shape[0]) for i in range(m.shape[1]) if m[i, j]])
    return out


@torch.no_grad()
def arm_state_fp(model, states, boxes, box_ptr, problem=None, want_counted=False):
    _lib.require_cuda(states, "states")
    lib = _lib.load()
    _lib.handle(_dev_index(states))
    if states.dtype not in _DT:
        raise TypeError("states must be float32 or float64")
    dof = arm_model_info(model)[0]
    states = states.reshape(-1, dof).contiguous()
    n = states.shape[0]
    if problem is not None:
        problem = problem.to(device=states.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=states.device)
    counted = torch.empty(n, dtype=torch.uint8, device=states.device) if want_counted else None
    _lib.check(lib.gmp_arm_state_fp(int(model), _lib.ptr(states), _DT[states.dtype], _lib.ptr(boxes), _lib.ptr(box_ptr),
                                    _lib.ptr(problem), n, _lib.ptr(free), _lib.ptr(counted), _lib.stream_ptr(states.device)))
    return (free, counted) if want_counted else free


@torch.no_grad()
def arm_edge_fp(model, a, b, boxes, box_ptr, problem=None, rrt_eps=0.5, want_checks=False):
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    lib = _lib.load()
    _lib.handle(_dev_index(a))
    if a.dtype not in _DT or b.dtype != a.dtype:
        raise TypeError("a and b must both be float32 or both float64")
    dof = arm_model_info(model)[0]
    a = a.reshape(-1, dof).contiguous()
    b = b.reshape(-1, dof).contiguous()
    n = a.shape[0]
    if problem is not None:
        problem = problem.to(device=a.device, dtype=torch.int32).contiguous()
    free = torch.empty(n, dtype=torch.uint8, device=a.device)
    checks = torch.empty(n, dtype=torch.int32, device=a.device) if want_checks else None
    _lib.check(lib.gmp_arm_edge_fp(int(model), _lib.ptr(a), _lib.ptr(b), _DT[a.dtype], _lib.ptr(boxes), _lib.ptr(box_ptr),
                                   _lib.ptr(problem), n, float(rrt_eps), _lib.ptr(free), _lib.ptr(checks),
                                   _lib.stream_ptr(a.device)))
    return (free, checks) if want_checks else free


_arm_ws = {}          # device index -> scratch of the fast graph form (grown on demand, reused across calls)
_max_boxes = {}       # (data_ptr, numel) of a box_ptr tensor -> its largest per-problem box count


def _max_boxes_of(box_ptr):
    key = (box_ptr.data_ptr(), box_ptr.numel())
    if key not in _max_boxes:
        if len(_max_boxes) > 64:
            _max_boxes.clear()
        _max_boxes[key] = int((box_ptr[1:] - box_ptr[:-1]).max()) if box_ptr.numel() > 1 else 0
    return _max_boxes[key]


def arm_last_undecided(device):
    """Number of interpolated states the fp32 filter of the last fast-form call on `device` left to the exact fp64 kernel
    (diagnostics; synchronises)."""
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    ws, off = _arm_ws.get(idx), _arm_ws.get((idx, "counter_off"))
    if ws is None or off is None:
        return 0
    return int(ws[off:off + 4].view(torch.int32)[0])


@torch.no_grad()
def arm_edge_fp_graph(model, v, edge_index, node_ptr_d, edge_ptr_d, boxes, box_ptr, n_edges_total, rrt_eps=0.5,
                      problem_of_graph=None, want_checks=False, free_out=None, checks_out=None, mode="fast"):
    """Every edge of a packed batch of graphs against its problem's boxes (kuka_env.py:389-411, one launch set).
    mode "fast" (default): fp32 filter + exact fp64 decision of the undecided states, lanes pulling edges from per-CTA
    queues; mode "exact": the fp64 thread-per-edge form.  Bit-identical results (tests/test_gpu_arm.py)."""
    lib = _lib.load()
    if mode == "fast":
        _lib.handle(_dev_index(v))
        B = node_ptr_d.numel() - 1
        free = free_out if free_out is not None else torch.empty(n_edges_total, dtype=torch.uint8, device=v.device)
        checks = checks_out
        if want_checks and checks is None:
            checks = torch.empty(n_edges_total, dtype=torch.int32, device=v.device)
        need = lib.gmp_arm_edge_graph_workspace_bytes(v.shape[0], n_edges_total)
        ws = _arm_ws.get(_dev_index(v))
        if ws is None or ws.numel() < need:
            ws = _arm_ws[_dev_index(v)] = torch.empty(int(need * 1.1) + 1024, dtype=torch.uint8, device=v.device)
        _arm_ws[(_dev_index(v), "counter_off")] = (v.shape[0] + 255) // 256 * 256
        _lib.check(lib.gmp_arm_edge_fp_graph_fast(int(model), _lib.ptr(v), v.shape[0], _lib.ptr(edge_index), edge_index.stride(0),
                                                  _lib.ptr(node_ptr_d), _lib.ptr(edge_ptr_d), _lib.ptr(problem_of_graph), B, n_edges_total,
                                                  _lib.ptr(boxes), _lib.ptr(box_ptr), _max_boxes_of(box_ptr), float(rrt_eps), _lib.ptr(ws),
                                                  ws.numel(), _lib.ptr(free), _lib.ptr(checks), _lib.stream_ptr(v.device)))
        return (free, checks) if want_checks else free
    _lib.handle(_dev_index(v))
    B = node_ptr_d.numel() - 1
    free = free_out if free_out is not None else torch.empty(n_edges_total, dtype=torch.uint8, device=v.device)
    checks = checks_out
    if want_checks and checks is None:
        checks = torch.empty(n_edges_total, dtype=torch.int32, device=v.device)
    # endpoint states are checked once per node (cached form): same booleans and counts, ~3 of (2 + K) evaluations per edge fewer
    flags = torch.empty(max(v.shape[0], 1), dtype=torch.uint8, device=v.device)
    _lib.check(lib.gmp_arm_edge_fp_graph_cached(int(model), _lib.ptr(v), v.shape[0], _lib.ptr(edge_index), edge_index.stride(0),
                                                _lib.ptr(node_ptr_d), _lib.ptr(edge_ptr_d), _lib.ptr(problem_of_graph), B, n_edges_total,
                                                _lib.ptr(boxes), _lib.ptr(box_ptr), float(rrt_eps), _lib.ptr(flags), _lib.ptr(free),
                                                _lib.ptr(checks), _lib.stream_ptr(v.device)))
    return (free, checks) if want_checks else free
