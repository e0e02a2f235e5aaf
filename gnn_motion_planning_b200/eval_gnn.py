"""Host-side mirror of the hot-path call sites of reference ``eval_gnn.py``.

* ``Data``        -- the two-method stand-in for ``torch_geometric.data.Data`` the planner needs
                     (``.to(device)``, ``.to_dict()``; eval_gnn.py:151,194).
* ``create_data`` -- eval_gnn.py:150-165; the k-NN graph is built by the CUDA kernel (``gmp_knn_graph``).
* ``obs_data``    -- eval_gnn.py:25-36.
* ``explore``     -- eval_gnn.py:168-276, the CALLER of the hot path (boundary, not the hot path itself):
                     kept sequential and host-side exactly like the reference so that results can be
                     compared problem by problem; its three hot calls (create_data, model(...),
                     env._edge_fp) land in the CUDA kernels.
"""
from time import time

import numpy as np
import torch

from . import graph

loop = 5  # eval_gnn.py:14


class DotDict(dict):
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


class Data:
    """Attribute bag with ``.to(device)`` and ``.to_dict()`` (what eval_gnn.py uses of PyG's Data)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    def to_dict(self):
        return dict(self.__dict__)


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def obs_data(env, free, collided):
    """eval_gnn.py:25-36."""
    device = _device()
    return DotDict({
        'free': torch.FloatTensor(np.array(free)).to(device),
        'collided': torch.FloatTensor(np.array(collided))[:len(free)].to(device),
        'obstacles': torch.FloatTensor(np.asarray(env.obstacles)).to(device),
    })


def path_cost(path):
    path = np.array(path)
    cost = 0
    for i in range(0, len(path) - 1):
        cost += np.linalg.norm(path[i + 1] - path[i])
    return cost


def to_np(tensor):
    return tensor.data.cpu().numpy()


def eval_gnn(str, seed, env, indexes, model=None, model_s=None, use_tqdm=False, smooth=True, batch=500, t_max=500, k=30,
             weights_root=".", **kwargs):
    """eval_gnn.py:96-145: run `explore` over a list of problems and reduce the per-problem tuples."""
    import random
    from .str2name import str2name
    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)                                           # config.set_random_seed (config.py:48-51)
    import os
    if model is None:
        _, model, model_path, _, _ = str2name(str, make_env=False)
        model.load_state_dict(torch.load(os.path.join(weights_root, model_path), map_location="cpu"))
    if model_s is None and smooth:
        _, _, _, model_s, model_s_path = str2name(str, make_env=False)
        model_s.load_state_dict(torch.load(os.path.join(weights_root, model_s_path), map_location="cpu"))
    model.eval()
    if model_s is not None:
        model_s.eval()
    solutions, paths, smooth_paths = [], [], []
    for index in indexes:
        env.init_new_problem(index)
        result = explore(env, model, model_s, smooth, batch=batch, t_max=t_max, k=k, **kwargs)
        paths.append(result['path'])
        smooth_paths.append(result['smooth_path'])
        solutions.append((result['success'], path_cost(result['path']), path_cost(result['smooth_path']), result['c_explore'],
                          result['c_smooth'], result['total'], result['total_explore']))
    n_success = sum([s[0] for s in solutions])
    collision_explore = np.mean([s[3] for s in solutions])
    collision = np.mean([(s[3] + s[4]) for s in solutions])
    running_time = float(sum([s[5] for s in solutions if s[0]])) / max(n_success, 1)
    solution_cost = float(sum([(s[2]) for s in solutions if s[0]])) / max(n_success, 1)
    total_time = sum([s[5] for s in solutions])
    total_time_explore = sum([s[6] for s in solutions])
    return n_success, collision, running_time, solution_cost, total_time, paths, smooth_paths, collision_explore, total_time_explore


def create_data(free, collided, env, k):
    """eval_gnn.py:150-165.  Returns ``Data(goal, v, labels, edge_index)``; tensors live on the GPU."""
    device = _device()
    data = Data(goal=torch.FloatTensor(np.asarray(env.goal_state)).to(device))
    free_t = torch.FloatTensor(np.array(free)).reshape(len(free), -1)
    coll_t = torch.FloatTensor(np.array(collided)).reshape(len(collided), free_t.shape[1])
    data.v = torch.cat((free_t, coll_t), dim=0).to(device).contiguous()
    data.labels = torch.zeros(len(data.v), 3)
    data.labels[:len(free), 0] = 1
    data.labels[len(free):, 1] = 1
    data.labels[1, 2] = 1
    k1 = graph.k1_of(k, len(free))
    data.edge_index = graph.knn_graph_edges(data.v, len(free), k1)
    return data


@torch.no_grad()
def explore(env, model, model_s, smooth=True, batch=500, t_max=1000, k=30, smoother='model', loop=5):
    """eval_gnn.py:168-276, statement for statement, with one documented deviation: the explored-edge
    mask of eval_gnn.py:202 is applied with the torch<=1.x meaning the author ran with (a (2,M) ndarray index
    acts as the tuple (rows, cols)); under torch 2.x the literal statement zeroes whole rows and the search
    never starts (see tests/golden/make_golden.py)."""
    from . import smoother as smoother_mod
    c0 = env.collision_check_count
    t0 = time()
    forward = 0

    success = False
    path, smooth_path = [], []
    n_batch = batch
    free, collided = env.sample_n_points(n_batch, need_negative=True)
    collided = collided[:len(free)]
    free = [env.init_state] + [env.goal_state] + list(free)

    explored = [0]
    explored_edges = [[0, 0]]
    costs = {0: 0.}
    prev = {0: 0}

    data = create_data(free, collided, env, k)
    device = _device()

    while not success and (len(free) - 2) <= t_max:
        t1 = time()
        policy = model(**data.to(device).to_dict(), **obs_data(env, free, collided), loop=loop)
        policy = policy.cpu()
        forward += time() - t1
        v_np = to_np(data.v)
        n = len(v_np)
        collided_mask = (data.labels[:, 1] == 1)

        policy[torch.arange(n), torch.arange(n)] = 0
        policy[:, explored] = 0
        policy[:, collided_mask] = 0
        policy[collided_mask, :] = 0
        ee = np.array(explored_edges).reshape(2, -1)
        policy[torch.from_numpy(ee[0]), torch.from_numpy(ee[1])] = 0
        success = False
        while policy[explored, :].sum() != 0:
            sub = policy[explored, :]
            nz = torch.where(sub != 0)
            agent = sub[nz[0], nz[1]].argmax()
            end_a, end_b = int(nz[0][agent]), int(nz[1][agent])
            end_a = explored[end_a]
            explored_edges.extend([[end_a, end_b], [end_b, end_a]])
            if env._edge_fp(v_np[end_a], v_np[end_b]):
                explored.append(end_b)
                costs[end_b] = costs[end_a] + np.linalg.norm(v_np[end_a] - v_np[end_b])
                prev[end_b] = end_a
                policy[:, end_b] = 0
                if env.in_goal_region(v_np[end_b]):
                    success = True
                    path = [end_b]
                    node = end_b
                    while node != 0:
                        path.append(prev[node])
                        node = prev[node]
                    path.reverse()
                    break
            else:
                policy[end_a, end_b] = 0
                policy[end_b, end_a] = 0

        if not success:
            if not smooth:
                return []
            if (n_batch + len(free) - 2) > t_max:
                break
            new_free, new_collided = env.sample_n_points(n_batch, need_negative=True)
            free = free + list(new_free)
            collided = collided + list(new_collided)
            collided = collided[:len(free)]
            data = create_data(free, collided, env, k)

    c_explore = env.collision_check_count - c0
    c1 = env.collision_check_count
    t1 = time()
    if success and smooth:
        path = list(to_np(data.v)[path])
        if smoother == 'model':
            smooth_path = smoother_mod.model_smooth(model_s, free, collided, path, env)
        elif smoother == 'oracle':
            raise NotImplementedError("the classical joint_smoother is a training-label oracle, out of scope (SURVEY.md section 2)")
        else:
            smooth_path = path
    c_smooth = env.collision_check_count - c1
    if smooth:
        total_time = time()
        return {'c_explore': c_explore, 'c_smooth': c_smooth, 'data': data, 'explored': explored, 'forward': forward,
                'total': total_time - t0, 'total_explore': t1 - t0, 'success': success, 't0': t0, 'path': path,
                'smooth_path': smooth_path, 'explored_edges': explored_edges}
    else:
        return list(to_np(data.v)[path]), free, collided
