"""Host-side mirror of ``construct_graph`` -- reference ``algorithm/dijkstra.py:15-31`` (the PRM dataset builder; the same body
is ``LazySP.construct_graph``, ``algorithm/lazy_sp.py:123-137``, and ``eval_bit.py:19-35``): k-NN graph over a point set, then
EVERY edge collision-checked.  SURVEY.md 8(f)-4: exactly "k-NN kernel + batched edge-check kernel".

The reference checks one edge per Python iteration; here the graph comes from ``gmp_knn_graph`` and all E checks are ONE
launch (``env.edge_fp_batch`` -> ``gmp_maze_edge_fp`` / ``gmp_arm_edge_fp`` in the points' own dtype, float64 for the
reference's callers).  Return value, dtypes, edge order and the ``collision_check_count`` side effect are the reference's.
"""
from collections import defaultdict

import numpy as np
import torch

from .. import graph

INFINITY = float('inf')


def construct_graph(env, points, check_collision=True, k=5):
    """-> (edge_cost: dict node -> [cost of each in-edge, INFINITY if blocked], neighbors: dict node -> [source of each in-edge],
    edge_index: int64 ndarray [E,2] (source, target) sorted by source*N+target, edge_free: list of bool)."""
    points = np.asarray(points)
    n = len(points)
    dev = getattr(env, "device", torch.device("cuda", torch.cuda.current_device()))
    v = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(dev)        # knn_graph(torch.FloatTensor(points), k=5, loop=True)
    ei = graph.knn_graph_edges(v, n, k)                                                 # + flip + coalesce   (:16-18)
    edge_index = ei.cpu().numpy().T.copy()
    a, b = points[edge_index[:, 0]], points[edge_index[:, 1]]
    free = env.edge_fp_batch(a, b)                                                      # env._edge_fp(points[e0], points[e1])  (:24)
    edge_cost, neighbors = defaultdict(list), defaultdict(list)
    for (s, t), f in zip(edge_index, free):
        # (per-edge 1-D norm, as the reference: NumPy's vector norm goes through dot() and can differ from an axis-wise
        #  sqrt(sum(d*d)) in the last bit)
        edge_cost[t].append(np.linalg.norm(points[t] - points[s]) if f else INFINITY)
        neighbors[t].append(s)
    return edge_cost, neighbors, edge_index, [bool(f) for f in free]
