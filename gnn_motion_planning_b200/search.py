"""Batched planner loop: reference ``eval_gnn.explore`` (eval_gnn.py:168-276) for MANY maze problems at once, with the
inner lazy tree search on the device.

The reference runs one problem at a time: sample -> ``create_data`` -> ``model(...)`` -> a host loop that takes the arg-max of a
masked dense policy and collision-checks ONE edge per Python iteration.  ``explore_batch`` keeps the semantics problem by
problem -- same samples (each problem owns a NumPy ``RandomState`` whose stream is the one ``np.random.seed(seed)`` would give
the reference), same graphs, same search order, same ``collision_check_count`` -- and runs every stage as one batched launch:

    sampling     ``gmp_maze_state_fp`` over the draws of all active problems (the RNG stream stays on the host: the reference
                 couples every draw to the outcome of the previous check, ``maze_env.py:85-100``)
    graphs       ``gmp_knn_graph``         (``create_data``, eval_gnn.py:150-165)
    logits       ``gmp_explorer_forward``  (sparse ``[E]``; the dense ``[N,N]`` matrix never exists)
    search       ``gmp_maze_tree_search``  one CTA per problem, optional speculative edge checks (``spec_k``)

Problems that exhaust their graph are resampled and continue in the next round (eval_gnn.py:235-247) with their tree, their
explored-edge list and their counters kept on the device.
"""
import numpy as np
import torch

from . import _lib, collision, graph
from .environment.env_config import LIMITS, RRT_EPS

STATUS_SUCCESS, STATUS_EXHAUSTED, STATUS_CAPACITY = 1, 2, 3


class TreeSearchState:
    """Persistent per-problem search state on the device (one row per problem)."""

    def __init__(self, n_problems, cap_nodes, cap_explored_edges, device):
        z = lambda *shape: torch.zeros(*shape, dtype=torch.int32, device=device)  # noqa: E731
        self.S, self.cap_nodes, self.cap_elist = n_problems, int(cap_nodes), int(cap_explored_edges)
        self.explored, self.prev, self.path = z(n_problems, cap_nodes), z(n_problems, cap_nodes), z(n_problems, cap_nodes)
        self.elist = z(n_problems, cap_explored_edges)
        self.n_explored, self.n_elist, self.n_checks, self.n_spec = z(n_problems), z(n_problems), z(n_problems), z(n_problems)
        self.status, self.path_len = z(n_problems), z(n_problems)
        self.path_cost = torch.zeros(n_problems, dtype=torch.float32, device=device)
        self._ws = None

    def workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=device)
        return self._ws


@torch.no_grad()
def maze_tree_search(state, v, node_ptr_d, n_free_d, edge_index, edge_ptr_d, logits, goal64, maps, problem_of_graph, slot_of_graph,
                     n_nodes_total, n_edges_total, spec_k=1, first_round=True):
    """One search round for a packed batch of graphs (see ``gmp_maze_tree_search`` in include/gnnmp.h).  All arguments are CUDA
    tensors except the three integers; results land in ``state``."""
    lib = _lib.load()
    for name, t in (("v", v), ("edge_index", edge_index), ("logits", logits), ("goal64", goal64), ("maps", maps)):
        _lib.require_cuda(t, name)
    if maps.dtype != torch.uint8 or goal64.dtype != torch.float64 or logits.dtype != torch.float32:
        raise TypeError("maps must be uint8, goal float64, logits float32")
    B = node_ptr_d.numel() - 1
    ws = state.workspace(lib.gmp_tree_search_workspace_bytes(B, n_nodes_total, n_edges_total), v.device)
    _lib.check(lib.gmp_maze_tree_search(
        _lib.ptr(v), _lib.ptr(node_ptr_d), _lib.ptr(n_free_d), _lib.ptr(edge_index), edge_index.stride(0), _lib.ptr(edge_ptr_d),
        _lib.ptr(logits), _lib.ptr(goal64), _lib.ptr(maps), _lib.ptr(problem_of_graph), _lib.ptr(slot_of_graph), B, n_nodes_total,
        n_edges_total, int(spec_k), int(bool(first_round)), _lib.ptr(state.explored), _lib.ptr(state.n_explored), _lib.ptr(state.prev),
        _lib.ptr(state.elist), _lib.ptr(state.n_elist), _lib.ptr(state.n_checks), _lib.ptr(state.n_spec), _lib.ptr(state.status),
        _lib.ptr(state.path), _lib.ptr(state.path_len), _lib.ptr(state.path_cost), state.cap_nodes, state.cap_elist, _lib.ptr(ws), ws.numel(),
        _lib.stream_ptr(v.device)))


@torch.no_grad()
def result_rows(state, first_problem_id=0):
    """[S,6] f32 rows (problem id, success, path cost, search checks, uncommitted speculative checks, explored nodes): the
    per-problem tuple of eval_gnn.py:120-134 and the payload of the multi-GPU all-gather."""
    rows = torch.empty((state.S, 6), dtype=torch.float32, device=state.status.device)
    _lib.check(_lib.load().gmp_search_result_rows(_lib.ptr(state.status), _lib.ptr(state.path_cost), _lib.ptr(state.n_checks),
                                                  _lib.ptr(state.n_spec), _lib.ptr(state.n_explored), state.S, int(first_problem_id),
                                                  _lib.ptr(rows), _lib.stream_ptr(state.status.device)))
    return rows


class _Stream:
    """A problem's NumPy RandomState with a read-ahead buffer: draws come out in exactly the order np.random.uniform(-1, 1, 2)
    calls would produce them; draws that were fetched but not consumed stay queued for the next request (no get_state /
    set_state rewinds, which cost more than the sampling itself)."""

    def __init__(self, seed, lo=None, hi=None):
        self.rs = np.random.RandomState(int(seed))
        self.lo = -LIMITS[:2] if lo is None else np.asarray(lo, np.float64)      # maze: uniform(-LIMITS, LIMITS) (maze_env.py:131)
        self.hi = LIMITS[:2] if hi is None else np.asarray(hi, np.float64)       # arms: uniform(pose_range) (kuka_env.py:216)
        self.buf = np.zeros((0, len(self.lo)))

    def peek(self, n):
        if len(self.buf) < n:
            self.buf = np.concatenate([self.buf, self.rs.uniform(self.lo, self.hi, (n - len(self.buf), len(self.lo)))])
        return self.buf[:n]

    def consume(self, n):
        self.buf = self.buf[n:]


def _sample_batch(rngs, need, maps_d, problem_ids, device, check=None):
    """``env.sample_n_points(need[p], need_negative=True)`` (maze_env.py:85-100, kuka_env.py:194-209) for several problems at
    once: the draws of all problems go through ONE state-check launch per pass; every stream ends exactly where the reference's
    would.  ``check(states_f64_cuda, problem_i32_cuda) -> free u8`` defaults to the 2-D maze check against ``maps_d``."""
    if check is None:
        check = lambda st_, pr_: collision.maze_state_fp(st_, maps_d, pr_)  # noqa: E731
    P = len(rngs)
    free = [[] for _ in range(P)]          # per problem: list of [m,2] float64 chunks, concatenated at the end
    coll = [[] for _ in range(P)]
    n_got = [0] * P
    counted = np.zeros(P, np.int64)
    todo = [p for p in range(P) if need[p] > 0]
    while todo:
        draws = [rngs[p].peek(max(64, int((need[p] - n_got[p]) * 2.5))) for p in todo]
        allp = np.concatenate(draws)
        prob = np.repeat(np.asarray([problem_ids[p] for p in todo], np.int32), [len(d) for d in draws])
        ok = check(torch.from_numpy(allp).to(device), torch.from_numpy(prob).to(device)).cpu().numpy().astype(bool)
        off, nxt = 0, []
        for p, d in zip(todo, draws):
            f = ok[off:off + len(d)]
            off += len(d)
            missing = need[p] - n_got[p]
            idx = np.flatnonzero(f)
            used = len(d) if len(idx) < missing else int(idx[missing - 1]) + 1      # exactly the draws the reference would make
            d, f = d[:used], f[:used]
            rngs[p].consume(used)
            counted[p] += used              # every draw lies inside the limits: one counted lookup each (maze_env.py:272-276)
            free[p].append(d[f])
            coll[p].append(d[~f])
            n_got[p] += int(f.sum())
            if n_got[p] < need[p]:
                nxt.append(p)
        todo = nxt
    z = np.zeros((0, len(rngs[0].lo) if P else 2))
    return [np.concatenate(x) if x else z for x in free], [np.concatenate(x) if x else z for x in coll], counted


def _sample_batch_device(seed, streams, next_draw, need, maps_d, problem_ids):
    """The same contract through the counter-based device sampler (``gmp_maze_sample_points``): one launch for all problems, no
    host RNG.  ``next_draw`` (per problem) is advanced by the draws consumed."""
    n = need[0]
    assert all(x == n for x in need)
    free_d, coll_d, n_coll, n_draws = collision.maze_sample_points(maps_d, problem_ids, streams, n, seed, first_draw=next_draw,
                                                                   cap_collided=16 * n)
    free_h, coll_h, n_coll, n_draws = free_d.cpu().numpy(), coll_d.cpu().numpy(), n_coll.cpu().numpy(), n_draws.cpu().numpy()
    free = [free_h[i] for i in range(len(streams))]
    coll = [coll_h[i, :min(int(n_coll[i]), coll_h.shape[1])] for i in range(len(streams))]
    for i in range(len(streams)):
        next_draw[i] += int(n_draws[i])
    return free, coll, n_draws.astype(np.int64)


class _MazeAdapter:
    """2-D maze problems (rows of maps / init_states / goal_states)."""

    def __init__(self, maps, init_states, goal_states, problem_ids, seeds, sampler, dev):
        self.dim, self.dev, self.sampler = 2, dev, sampler
        self.problem_ids = [int(p) for p in problem_ids]
        self.seeds = list(seeds)
        self.maps = np.asarray(maps)
        self.maps_d = torch.as_tensor(np.ascontiguousarray(self.maps != 0).astype(np.uint8)).to(dev)
        self.init = np.asarray(init_states, np.float64)[self.problem_ids]
        self.goal = np.asarray(goal_states, np.float64)[self.problem_ids]
        self.rngs = [_Stream(s_) for s_ in seeds]
        self.next_draw = [0] * len(self.problem_ids)
        self.obs = [(np.argwhere(self.maps[p] == 1) / 15.0 - 0.5).astype(np.float32) for p in self.problem_ids]   # maze_env.py:73-79
        self.obs_width = 2

    def sample(self, which, n):
        pids = [self.problem_ids[p] for p in which]
        if self.sampler == "device":
            nd = [self.next_draw[p] for p in which]
            out = _sample_batch_device(0x9E3779B97F4A7C15, [self.seeds[p] for p in which], nd, [n] * len(which), self.maps_d, pids)
            for i, p in enumerate(which):
                self.next_draw[p] = nd[i]
            return out
        return _sample_batch([self.rngs[p] for p in which], [n] * len(which), self.maps_d, pids, self.dev)

    def search(self, st, v_d, node_ptr_d, n_free_d, ei, edge_ptr_d, logits, goal64, active, n_nodes, n_edges, spec_k, first):
        slots = torch.tensor(active, dtype=torch.int32, device=self.dev)
        probs = torch.tensor([self.problem_ids[p] for p in active], dtype=torch.int32, device=self.dev)
        maze_tree_search(st, v_d, node_ptr_d, n_free_d, ei, edge_ptr_d, logits, goal64, self.maps_d, probs, slots, n_nodes, n_edges,
                         spec_k=spec_k, first_round=first)


class _ArmAdapter:
    """Arm problems: list of (obstacles [(halfExtents, basePosition), ...], start, goal) as maze_files/kukas_*.pkl stores them."""

    def __init__(self, arm_model, problems, seeds, rrt_eps, dev):
        self.dev, self.arm_model, self.rrt_eps = dev, int(arm_model), float(rrt_eps)
        self.dim, lo, hi = collision.arm_model_info(arm_model)
        self.seeds = list(seeds)
        self.boxes_d, self.box_ptr_d = collision.pack_boxes([p[0] for p in problems], dev)
        self.init = np.asarray([np.asarray(p[1], np.float64).reshape(-1) for p in problems])
        self.goal = np.asarray([np.asarray(p[2], np.float64).reshape(-1) for p in problems])
        self.rngs = [_Stream(s_, lo, hi) for s_ in seeds]
        bp = self.box_ptr_d.cpu().numpy()
        boxes = self.boxes_d.cpu().numpy()
        self.obs = [boxes[bp[i]:bp[i + 1]].astype(np.float32) for i in range(len(problems))]     # FloatTensor(env.obstacles).view(-1, 6)
        self.obs_width = 6

    def sample(self, which, n):
        check = lambda st_, pr_: collision.arm_state_fp(self.arm_model, st_, self.boxes_d, self.box_ptr_d, pr_)  # noqa: E731
        return _sample_batch([self.rngs[p] for p in which], [n] * len(which), None, list(which), self.dev, check=check)

    def search(self, st, v_d, node_ptr_d, n_free_d, ei, edge_ptr_d, logits, goal64, active, n_nodes, n_edges, spec_k, first):
        lib = _lib.load()
        slots = torch.tensor(active, dtype=torch.int32, device=self.dev)
        B = node_ptr_d.numel() - 1
        ws = st.workspace(lib.gmp_tree_search_workspace_bytes(B, n_nodes, n_edges), self.dev)
        _lib.check(lib.gmp_arm_tree_search(
            self.arm_model, _lib.ptr(v_d), _lib.ptr(node_ptr_d), _lib.ptr(n_free_d), _lib.ptr(ei), ei.stride(0), _lib.ptr(edge_ptr_d),
            _lib.ptr(logits), _lib.ptr(goal64), _lib.ptr(self.boxes_d), _lib.ptr(self.box_ptr_d), _lib.ptr(slots), self.rrt_eps,
            _lib.ptr(slots), B, n_nodes, n_edges, int(spec_k), int(bool(first)), _lib.ptr(st.explored), _lib.ptr(st.n_explored),
            _lib.ptr(st.prev), _lib.ptr(st.elist), _lib.ptr(st.n_elist), _lib.ptr(st.n_checks), _lib.ptr(st.n_spec), _lib.ptr(st.status),
            _lib.ptr(st.path), _lib.ptr(st.path_len), _lib.ptr(st.path_cost), st.cap_nodes, st.cap_elist, _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(self.dev)))


@torch.no_grad()
def explore_batch(model, maps, init_states, goal_states, problem_ids, seeds, batch=100, t_max=100, k=10, loop=5, spec_k=1,
                  device=None, max_checks=20000, sampler="numpy", timings=None):
    """``explore(env, model, None, smooth=True, batch, t_max, k, smoother='none')`` for every problem of ``problem_ids``
    (rows of ``maps`` / ``init_states`` / ``goal_states``), seeded like ``np.random.seed(seeds[i])`` before the reference call.

    sampler: "numpy" -- every problem owns np.random.RandomState(seeds[i]): the reference's exact stream (parity tests);
             "device" -- the counter-based Philox sampler on the GPU (seeds[i] = stream id; a new stream, same semantics).

    Returns one dict per problem: success, path (float32 waypoints), path_nodes, explored (node ids in tree order), c_explore
    (= env.collision_check_count delta: sampling + edge + goal-region checks), spec_checks (speculative edge checks that were
    never committed; 0 when spec_k == 1), n_nodes, rounds, path_cost."""
    dev = torch.device(device if device is not None else "cuda")
    ad = _MazeAdapter(maps, init_states, goal_states, problem_ids, seeds, sampler, dev)
    return _explore_core(ad, model, batch, t_max, k, loop, spec_k, max_checks, timings)


@torch.no_grad()
def explore_batch_arm(model, arm_model, problems, seeds, rrt_eps=0.5, batch=100, t_max=100, k=10, loop=5, spec_k=1, device=None,
                      max_checks=20000, timings=None):
    """The same batched planner loop for the arm environments (``collision.ARM_KUKA7`` ...): ``problems[i]`` = (obstacles, start,
    goal) as the reference's problem files hold them; sampling follows ``KukaEnv.sample_n_points`` on the reference's NumPy
    stream, the search runs in ``gmp_arm_tree_search`` (``spec_k`` > 1 checks several candidate edges per iteration)."""
    dev = torch.device(device if device is not None else "cuda")
    ad = _ArmAdapter(arm_model, problems, seeds, rrt_eps, dev)
    return _explore_core(ad, model, batch, t_max, k, loop, spec_k, max_checks, timings)


def _explore_core(ad, model, batch, t_max, k, loop, spec_k, max_checks, timings):
    import time as _time
    dev = ad.dev
    P = len(ad.seeds)

    def _tick(name, t0):
        if timings is not None:                       # phase wall times (synchronised): sample / pack / graph / forward / search
            torch.cuda.synchronize(dev)
            timings[name] = timings.get(name, 0.0) + _time.perf_counter() - t0
        return _time.perf_counter()
    goal64 = torch.from_numpy(np.ascontiguousarray(ad.goal)).to(dev)
    n_batch = batch
    cap_nodes = 2 * (t_max + 2 * n_batch + 2) + 8
    st = TreeSearchState(P, cap_nodes, 2 + 4 * max_checks, dev)

    t0 = _time.perf_counter()
    new_free, new_coll, counted = ad.sample(list(range(P)), n_batch)
    free = [np.concatenate([ad.init[p][None], ad.goal[p][None], new_free[p]]) for p in range(P)]
    coll = [new_coll[p][:len(new_free[p])] for p in range(P)]                      # collided = collided[:len(free)] BEFORE init/goal join (:180-181)
    c_sample = np.asarray(counted, np.int64).copy()
    t0 = _tick("sample", t0)
    active = list(range(P))
    rounds = np.zeros(P, np.int64)
    last_v = [None] * P
    first = True
    while active:
        # ---- create_data for every active problem (eval_gnn.py:150-165), packed
        vs, n_free, k1 = [], [], []
        for p in active:
            f, c = free[p], coll[p]
            vs.append(np.concatenate([f, c]).astype(np.float32))                  # torch.FloatTensor(np.array(...))
            n_free.append(len(f))
            k1.append(graph.k1_of(k, len(f)))
            last_v[p] = vs[-1]
        node_ptr = np.concatenate([[0], np.cumsum([len(x) for x in vs])]).astype(np.int32)
        v_d = torch.from_numpy(np.concatenate(vs)).to(dev)
        t0 = _tick("pack", t0)
        ei, edge_ptr = graph.knn_graph_batch(v_d, node_ptr, np.asarray(n_free, np.int32), np.asarray(k1, np.int32))
        et = int(edge_ptr[-1])
        t0 = _tick("graph", t0)
        # ---- model(**data, **obs_data, loop=loop): sparse logits (eval_gnn.py:194)
        obss = [ad.obs[p] for p in active]
        obs_ptr = np.concatenate([[0], np.cumsum([len(o) for o in obss])]).astype(np.int32)
        obs_d = torch.from_numpy(np.concatenate(obss).reshape(-1, ad.obs_width)).to(dev)
        goal_d = torch.from_numpy(ad.goal[active].astype(np.float32)).to(dev)
        logits = model.forward_batch(v_d, ei, goal_d, obs_d, node_ptr, edge_ptr, obs_ptr, loop=loop)
        t0 = _tick("forward", t0)
        # ---- the search itself
        ad.search(st, v_d, torch.from_numpy(node_ptr).to(dev), torch.tensor(n_free, dtype=torch.int32, device=dev), ei,
                  torch.from_numpy(edge_ptr).to(dev), logits, goal64, active, int(node_ptr[-1]), et, spec_k, first)
        first = False
        status = st.status.cpu().numpy()
        t0 = _tick("search", t0)
        nxt = []
        for p in active:
            rounds[p] += 1
            if status[p] == STATUS_CAPACITY:
                raise _lib.GnnmpError("tree search capacity exceeded for problem %d (raise max_checks)" % p)
            if status[p] == STATUS_EXHAUSTED and (n_batch + len(free[p]) - 2) <= t_max:      # eval_gnn.py:239-240
                nxt.append(p)
        if nxt:                                                                               # resample (:242-247)
            nf, nc, cnt = ad.sample(nxt, n_batch)
            for i, p in enumerate(nxt):
                free[p] = np.concatenate([free[p], nf[i]])
                coll[p] = np.concatenate([coll[p], nc[i]])[:len(free[p])]
                c_sample[p] += cnt[i]
            t0 = _tick("sample", t0)
        active = nxt
    # ---- results
    status = st.status.cpu().numpy()
    rows = result_rows(st).cpu().numpy()
    n_expl, n_chk, n_spec, plen = (t.cpu().numpy() for t in (st.n_explored, st.n_checks, st.n_spec, st.path_len))
    explored, path = st.explored.cpu().numpy(), st.path.cpu().numpy()
    out = []
    for p in range(P):
        ok = status[p] == STATUS_SUCCESS
        nodes = path[p, :plen[p]].tolist() if ok else []
        out.append(dict(success=bool(ok), path_nodes=nodes, path=[last_v[p][i] for i in nodes], explored=explored[p, :n_expl[p]].tolist(),
                        c_explore=int(c_sample[p] + n_chk[p]), c_search=int(n_chk[p]), spec_checks=int(n_spec[p]), n_nodes=len(last_v[p]),
                        rounds=int(rounds[p]), path_cost=float(rows[p, 2]), row=rows[p]))
    return out


def path_cost(path):
    """eval_gnn.py:53-58."""
    path = np.array(path)
    return float(sum(np.linalg.norm(path[i + 1] - path[i]) for i in range(len(path) - 1)))


__all__ = ["TreeSearchState", "maze_tree_search", "explore_batch", "explore_batch_arm", "path_cost", "RRT_EPS"]
