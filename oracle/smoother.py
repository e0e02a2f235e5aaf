"""ORACLE (test infrastructure, never the product path): CPU restatement of the reference GNN
path smoother forward ``ModelSmoother.forward`` -- reference ``model_smoother.py:104-142`` with
its add-aggregating ``MPNN`` (:22-39) -- and of the chain graph the caller builds
(``smoother.model_smooth``, reference ``smoother.py:238-241``).

Per loop iteration (:123-140):
    e2   = knn(x=nodes[P:], y=path, k=10).flip(0); e2[0] += P      sample -> path edges
    E    = coalesce(cat(edge_index, e2))
    x    = node_code([nodes | onehot(path,free,collided)])          Lin -> BatchNorm1d(eval) -> ReLU -> Lin
    h    = x + lin_1( sum_{e: dst=i} lin_0([x_j - x_i, x_j, x_i]) )
    path[1:-1] = smooth_node(h[:P])[1:-1];  nodes[:P] = path
returns path * scale.

PyG primitives restated per their published semantics (see oracle/knn_graph.py header for the
canonical distance rule); parity is pinned to the reference's own python through
``tests/golden/make_golden.py`` and UNPINNED at the PyG-primitive boundary.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import numpy as np
import torch


def chain_edge_index(p):
    """smoother.py:238-241: (i+1 -> i), (i -> i+1), then self loops appended."""
    a = torch.arange(1, p).reshape(1, -1)
    b = torch.arange(0, p - 1).reshape(1, -1)
    e = torch.cat((a, b), dim=0)
    e = torch.cat((e, e.flip(0)), dim=-1)
    loop = torch.arange(p)
    return torch.cat((e, torch.stack((loop, loop))), dim=-1)


def _knn_path_to_samples(samples, path, k):
    """knn(x=samples, y=path, k): for each path row the k nearest samples, canonical fp32 rule."""
    s = samples.to(torch.float32)
    p = path.to(torch.float32)
    d = torch.zeros(len(p), len(s), dtype=torch.float32)
    for c in range(s.shape[1]):
        diff = p[:, c:c + 1] - s[:, c].unsqueeze(0)
        d = d + diff * diff
    k = min(k, len(s))
    order = np.argsort(d.numpy(), axis=1, kind="stable")[:, :k]
    return torch.from_numpy(order.astype(np.int64))         # [P, k] sample indices


def _lin(x, sd, name):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


@torch.no_grad()
def smoother_forward(sd, path, free, collided, edge_index, loop=1, scale=1.0, dtype=torch.float32, knn_k=10):
    sd = {k: t.to(dtype) for k, t in sd.items() if t.is_floating_point()}
    path = path.to(dtype) / scale
    free = free.to(dtype) / scale
    collided = collided.to(dtype) / scale
    p, f = len(path), len(free)
    nodes = torch.cat((path, free, collided), dim=0)
    n = len(nodes)
    for _ in range(loop):
        nb = _knn_path_to_samples(nodes[p:], path, knn_k)              # [P,k]
        src2 = nb.reshape(-1) + p
        dst2 = torch.arange(p).repeat_interleave(nb.shape[1])
        key = torch.cat((edge_index[0].long() * n + edge_index[1].long(), src2 * n + dst2))
        key = torch.unique(key, sorted=True)                            # coalesce
        src, dst = key // n, key % n

        info = torch.zeros(n, 3, dtype=dtype)
        info[:p, 0] = 1
        info[p:p + f, 1] = 1
        info[p + f:, 2] = 1
        x = _lin(torch.cat((nodes, info), dim=-1), sd, "node_code.0")
        x = (x - sd["node_code.1.running_mean"]) / torch.sqrt(sd["node_code.1.running_var"] + 1e-5) \
            * sd["node_code.1.weight"] + sd["node_code.1.bias"]
        x = _lin(torch.relu(x), sd, "node_code.3")

        x_j, x_i = x[src], x[dst]
        msg = _lin(torch.relu(_lin(torch.cat((x_j - x_i, x_j, x_i), dim=-1), sd, "process.lin_0.0")),
                   sd, "process.lin_0.2")
        agg = torch.zeros(n, x.shape[1], dtype=dtype)
        agg.index_add_(0, dst, msg)
        h = x + _lin(torch.relu(_lin(agg, sd, "process.lin_1.0")), sd, "process.lin_1.2")
        new = _lin(h[:p], sd, "smooth_node")
        path = path.clone()
        path[1:-1] = new[1:-1]
        nodes = nodes.clone()
        nodes[:p] = path
    return path * scale
