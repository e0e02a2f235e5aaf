"""Host-side mirror of the hot-path call sites of reference ``smoother.py``: ``obs_data`` (:52-64),
``model_smooth`` (:233-246) and ``proposed_path_smootherv2`` (:194-216).  The smoother forward and
the edge checks run on the GPU; the steering loop is the reference's host logic (the caller)."""
from copy import deepcopy

import numpy as np
import torch


class DotDict(dict):
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def obs_data(env, free, collided):
    """smoother.py:52-64 (pads an all-zero row when a list is empty, truncates to 500 + 500)."""
    if not len(free):
        free.append([0. for _ in range(env.config_dim)])
    if not len(collided):
        collided.append([0. for _ in range(env.config_dim)])
    free = free[:500]
    collided = collided[:500]
    device = _device()
    return DotDict({
        'free': torch.FloatTensor(np.array(free)).to(device),
        'collided': torch.FloatTensor(np.array(collided)).to(device),
        'obstacles': torch.FloatTensor(np.asarray(env.obstacles)).to(device),
    })


def chain_edge_index(p):
    """smoother.py:238-241: (i+1 -> i), (i -> i+1), self loops appended."""
    a = torch.arange(1, p).reshape(1, -1)
    b = torch.arange(0, p - 1).reshape(1, -1)
    e = torch.cat((a, b), dim=0)
    e = torch.cat((e, e.flip(0)), dim=-1)
    loop = torch.arange(p)
    return torch.cat((e, torch.stack((loop, loop))), dim=-1)


def proposed_path_smootherv2(old_path, new_path, env):
    """smoother.py:194-216; the 2*(P-2) edge checks of a steering round are independent of each other only
    through `next_path`, which the reference updates in place while sweeping i -- kept sequential for parity."""
    K = int(np.ceil((np.linalg.norm(np.array(old_path) - np.array(new_path), axis=-1) / env.RRT_EPS).max()))
    path = deepcopy(old_path)
    for _ in range(K):
        diff = 0
        next_path = deepcopy(path)
        for i, ns in enumerate(zip(path[1:-1], new_path[1:-1])):
            i = i + 1
            old_n, new_n = ns
            dist = np.linalg.norm(old_n - new_n)
            if dist < env.RRT_EPS:
                next_path[i] = new_n
            else:
                next_path[i] = env.interpolate(old_n, new_n, env.RRT_EPS / dist)
            if not (env._edge_fp(next_path[i - 1], next_path[i]) and env._edge_fp(next_path[i + 1], next_path[i])):
                next_path[i] = path[i]
            else:
                diff += np.linalg.norm(next_path[i] - new_n)
        path = next_path
        if diff < 1e-5:
            return path
    return path


def model_smooth(model, free, collided, old_path, env, iter=5):
    """smoother.py:233-246."""
    device = _device()
    for _ in range(iter):
        data = obs_data(env, free, collided)
        data.path = torch.FloatTensor(np.array(old_path)).to(device)
        data.edge_index = chain_edge_index(len(old_path)).to(device)
        new_path = model(**data, loop=1).data.cpu().numpy()
        old_path = proposed_path_smootherv2(old_path, list(new_path), env)
    return old_path


# ---------------------------------------------------------------------------------------------------------------------------
# batched, device-resident form: smoother forwards and steering rounds for MANY maze problems per launch
@torch.no_grad()
def steer_rounds_batch(old_path_d, new_path_d, path_ptr_d, maps_d, problem_of_path_d, rrt_eps, want_cost=False):
    """``proposed_path_smootherv2`` (smoother.py:194-216) for a packed batch of 2-D maze paths on the device
    (``gmp_maze_steer_rounds``): -> (path [P_total,2] f32, n_checks [B] i32, n_rounds [B] i32[, path_cost [B] f32])."""
    from . import _lib
    B = path_ptr_d.numel() - 1
    dev = old_path_d.device
    out = torch.empty_like(old_path_d)
    checks = torch.empty(B, dtype=torch.int32, device=dev)
    rounds = torch.empty(B, dtype=torch.int32, device=dev)
    cost = torch.empty(B, dtype=torch.float32, device=dev) if want_cost else None
    _lib.check(_lib.load().gmp_maze_steer_rounds(_lib.ptr(old_path_d), _lib.ptr(new_path_d), _lib.ptr(path_ptr_d), _lib.ptr(maps_d),
                                                 _lib.ptr(problem_of_path_d), B, float(rrt_eps), _lib.ptr(out), _lib.ptr(checks),
                                                 _lib.ptr(rounds), _lib.ptr(cost), _lib.stream_ptr(dev)))
    return (out, checks, rounds, cost) if want_cost else (out, checks, rounds)


@torch.no_grad()
def model_smooth_batch(model, frees, collideds, paths, maps_d, problem_ids, rrt_eps=0.05, iter=5):
    """``model_smooth`` (smoother.py:233-246) for many maze problems at once: per iteration ONE batched smoother forward
    (``gmp_smoother_forward``) and ONE steering launch; paths stay on the device between iterations.
    frees / collideds / paths: per-problem lists of states (as the reference passes them).
    -> (list of float32 [P,2] paths, collision_check_count increments per problem [B], path costs [B])."""
    dev = maps_d.device
    B = len(paths)
    P = [len(p) for p in paths]
    path_ptr = np.concatenate([[0], np.cumsum(P)]).astype(np.int32)
    samples, n_free = [], []
    for f, c in zip(frees, collideds):                       # obs_data (smoother.py:52-64): pad empty lists, keep 500 + 500
        f = list(f) if len(f) else [[0.] * 2]
        c = list(c) if len(c) else [[0.] * 2]
        f, c = np.asarray(f[:500], np.float32).reshape(-1, 2), np.asarray(c[:500], np.float32).reshape(-1, 2)
        samples.append(np.concatenate([f, c]))
        n_free.append(len(f))
    sample_ptr = np.concatenate([[0], np.cumsum([len(x) for x in samples])]).astype(np.int32)
    samples_d = torch.from_numpy(np.concatenate(samples)).to(dev)
    eis = [chain_edge_index(p).numpy() for p in P]
    edge_ptr = np.concatenate([[0], np.cumsum([e.shape[1] for e in eis])]).astype(np.int32)
    ei_d = torch.from_numpy(np.concatenate(eis, 1)).to(dev)
    path_d = torch.from_numpy(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in paths])).to(dev)
    path_ptr_d = torch.from_numpy(path_ptr).to(dev)
    prob_d = torch.as_tensor(np.asarray(problem_ids, np.int32)).to(dev)
    total = torch.zeros(B, dtype=torch.int32, device=dev)
    cost = None
    for _ in range(iter):
        new_d = model.forward_batch(path_d, samples_d, ei_d, path_ptr, sample_ptr, n_free, edge_ptr, loop=1)
        path_d, checks, _, cost = steer_rounds_batch(path_d, new_d, path_ptr_d, maps_d, prob_d, rrt_eps, want_cost=True)
        total += checks
    out = path_d.cpu().numpy()
    return [out[path_ptr[i]:path_ptr[i + 1]] for i in range(B)], total.cpu().numpy(), cost.cpu().numpy()
