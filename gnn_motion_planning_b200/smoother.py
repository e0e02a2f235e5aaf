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
