"""Minimal stand-ins for the third-party packages the reference imports but this image lacks.

TEST INFRASTRUCTURE ONLY -- used by ``make_golden.py`` (run in the build container, where
``/root/reference`` exists) so that the reference's OWN ``model.py`` / ``model_smoother.py`` /
``eval_gnn.create_data`` / ``environment/maze_env.py`` code can be imported and executed to
produce golden vectors.  Nothing here is reference code: each stub restates the *published*
semantics of one PyG-family primitive, as used at the reference call sites:

* ``MessagePassing.propagate`` (flow source_to_target): ``x_j = x[edge_index[0]]``,
  ``x_i = x[edge_index[1]]``, aggregate at ``edge_index[1]`` with ``dim_size = N``; ``max``
  yields 0 for rows with no incoming edge (torch_scatter.scatter_max fill).  Call sites:
  reference ``model.py:33``, ``model_smoother.py:33``.
* ``knn(x, y, k)`` -> ``[2, len(y)*k]`` = (y index, x index), neighbours in ascending distance
  (torch_cluster).  Call sites ``model.py:132``, ``model_smoother.py:125``.
* ``knn_graph(x, k, loop=True)`` -> row 0 = neighbour, row 1 = centre.  ``eval_gnn.py:160,162``.
* ``coalesce(index, None, m, n)`` -> lexicographic sort by (row, col) + unique.
  ``eval_gnn.py:164``, ``model_smoother.py:128``.
* ``add_self_loops`` appends (i, i) for all i, no dedup.  ``smoother.py:241``.

Distances use the canonical rule of SURVEY.md App. C.1: fp32, squared L2 accumulated
left-to-right over dims, ties toward the lower index.
"""
import sys
import types

import numpy as np
import torch


def _sqdist_f32(y, x):
    """[len(y), len(x)] fp32 squared distances, accumulated dim by dim in fp32."""
    y = y.detach().cpu().to(torch.float32)
    x = x.detach().cpu().to(torch.float32)
    d = torch.zeros(len(y), len(x), dtype=torch.float32)
    for c in range(x.shape[1]):
        diff = y[:, c:c + 1] - x[:, c].unsqueeze(0)
        d = d + diff * diff
    return d


def knn(x, y, k, batch_x=None, batch_y=None, **kw):
    d = _sqdist_f32(y, x).numpy()
    k = min(k, len(x))
    # stable argsort == ties toward the lower index
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    rows = np.repeat(np.arange(len(y)), k)
    return torch.from_numpy(np.stack([rows, order.reshape(-1)]).astype(np.int64))


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **kw):
    e = knn(x, x, k if loop else k + 1)
    if not loop:
        e = e[:, e[0] != e[1]]
    return torch.stack([e[1], e[0]], dim=0)


def coalesce(index, value, m, n, op="add"):
    key = index[0].to(torch.int64) * n + index[1].to(torch.int64)
    key = torch.unique(key, sorted=True)
    return torch.stack([key // n, key % n], dim=0), value


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loop, loop])], dim=1), edge_attr


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kw):
        super().__init__()
        self.aggr = aggr
        assert flow == "source_to_target"

    def propagate(self, edge_index, x=None, edge_attr=None, **kw):
        xs = x if isinstance(x, (tuple, list)) else (x, x)
        x_j = xs[0][edge_index[0]]
        x_i = xs[1][edge_index[1]]
        if edge_attr is not None:
            msg = self.message(x_i=x_i, x_j=x_j, edge_attr=edge_attr)
        else:
            msg = self.message(x_i=x_i, x_j=x_j)
        n = xs[1].shape[0]
        dst = edge_index[1]
        if self.aggr == "add":
            out = msg.new_zeros(n, msg.shape[1])
            out.index_add_(0, dst, msg)
            return out
        if self.aggr == "max":
            out = msg.new_full((n, msg.shape[1]), float("-inf"))
            out = out.scatter_reduce(0, dst.unsqueeze(-1).expand_as(msg), msg, "amax", include_self=True)
            return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)
        raise NotImplementedError(self.aggr)


class Data:
    """torch_geometric.data.Data stand-in: attribute bag with .to() and .to_dict()."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    def to_dict(self):
        return dict(self.__dict__)


def _unused(*a, **k):
    raise RuntimeError("stubbed PyG symbol that the reference hot path never calls")


def install():
    """Register the stub modules in ``sys.modules`` (idempotent)."""
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torch_scatter", scatter_mean=_unused, scatter_max=_unused, scatter_add=_unused)
    mod("torch_sparse", coalesce=coalesce)
    mod("torch_cluster", knn=knn, knn_graph=knn_graph)
    tg = mod("torch_geometric")
    tg.nn = mod("torch_geometric.nn", voxel_grid=_unused, radius_graph=_unused, knn_graph=knn_graph,
                GraphConv=_unused, knn=knn)
    tg.nn.pool = mod("torch_geometric.nn.pool", knn=knn)
    mod("torch_geometric.nn.pool.consecutive", consecutive_cluster=_unused)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
    tg.utils = mod("torch_geometric.utils", grid=_unused, add_self_loops=add_self_loops,
                   remove_self_loops=_unused, softmax=_unused)
    tg.data = mod("torch_geometric.data", Data=Data)
    # reference nets.py is a dead layer zoo whose names are imported but never instantiated
    mod("nets", GATConv=_unused, EdgePooling=_unused, ASAPooling=_unused, SAModule=_unused,
        FPModule=_unused, MLP=_unused)
