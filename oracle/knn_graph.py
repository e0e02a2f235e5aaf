"""ORACLE (test infrastructure, never the product path): CPU restatement of the reference
k-NN random-geometric-graph construction ``create_data`` -- reference ``eval_gnn.py:150-165``.

    v  = float32(cat(free, collided));   N = F + C                         (:152-153)
    k1 = int(ceil(k * ln(F) / ln(100)))                                    (:159, float64 math)
    S  = kNN_k1(v[0:N]) U kNN_k1(v[0:F])      (knn_graph(..., loop=True))  (:160,:162)
    edge_index = sorted-unique(S U reverse(S)) by key src*N + dst, int64   (:161,:163-164)

``knn_graph`` / ``coalesce`` live in torch_cluster / torch_sparse (absent from /root/reference,
not installable here, versions unpinned by the reference README).  Their published semantics are
restated; the distance rule is made canonical: fp32, squared L2 accumulated
left-to-right over the dims (no FMA contraction), neighbours ordered by (distance, index) so ties
go to the lower index.  This coincides with torch_cluster except on exact fp32 ties at the k-th
neighbour.  "Bit-exact edge indices" is tested against THIS rule: parity unpinned vs torch_cluster.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import math

import numpy as np


def k1_of(k, n_free):
    """eval_gnn.py:159."""
    return int(np.ceil(k * np.log(n_free) / np.log(100)))


def sqdist_f32(x):
    """[n,n] canonical fp32 squared distances: ((0 + d0*d0) + d1*d1) + ... , each op rounded to fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    d = np.zeros((len(x), len(x)), dtype=np.float32)
    for c in range(x.shape[1]):
        diff = (x[:, c:c + 1] - x[None, :, c]).astype(np.float32)
        d = (d + (diff * diff).astype(np.float32)).astype(np.float32)
    return d


def knn_self(x, k):
    """For every centre i the k nearest j (self included), ascending (distance, index). [n, min(k,n)]."""
    d = sqdist_f32(x)
    k = min(k, len(x))
    return np.argsort(d, axis=1, kind="stable")[:, :k]


def knn_graph_edges(v, n_free, k1):
    """The symmetrised, coalesced edge set of create_data as int64 [2,E]; row0 = src, row1 = dst."""
    n = len(v)
    keys = []
    for cnt in (n, n_free):
        nb = knn_self(v[:cnt], k1)                         # nb[i] = neighbours j of centre i
        ctr = np.repeat(np.arange(cnt, dtype=np.int64), nb.shape[1])
        nbr = nb.reshape(-1).astype(np.int64)
        keys.append(nbr * n + ctr)                         # (j -> i): row0 = neighbour, row1 = centre
        keys.append(ctr * n + nbr)                         # flipped
    key = np.unique(np.concatenate(keys))
    return np.stack([key // n, key % n]).astype(np.int64)


def create_data(free, collided, goal_state, k):
    """Restates eval_gnn.create_data; returns a dict with goal, v, labels, edge_index (numpy)."""
    free = np.asarray(free, dtype=np.float64)
    collided = np.asarray(collided, dtype=np.float64).reshape(-1, free.shape[1])
    v = np.concatenate([free, collided], axis=0).astype(np.float32)
    labels = np.zeros((len(v), 3), dtype=np.float32)
    labels[:len(free), 0] = 1
    labels[len(free):, 1] = 1
    labels[1, 2] = 1
    k1 = k1_of(k, len(free))
    return dict(goal=np.asarray(goal_state, dtype=np.float32), v=v, labels=labels,
                edge_index=knn_graph_edges(v, len(free), k1), k1=k1)
