"""GPU: k-NN RGG construction vs golden (reference create_data) and vs the oracle; bit-exact edge indices."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["maze2", "kuka7", "kuka14", "dup"])
def test_create_data_golden(cuda_device, tag):
    from types import SimpleNamespace
    from gnn_motion_planning_b200.eval_gnn import create_data
    cd = np.load(os.path.join(G, "create_data.npz"))
    free, coll = list(cd[tag + "_free"]), list(cd[tag + "_collided"])
    d = create_data(free, coll, SimpleNamespace(goal_state=free[1]), int(cd[tag + "_k"]))
    assert np.array_equal(d.edge_index.cpu().numpy(), cd[tag + "_edge_index"])
    assert np.array_equal(d.v.cpu().numpy(), cd[tag + "_v"])
    assert np.array_equal(d.labels.cpu().numpy(), cd[tag + "_labels"])
    assert np.array_equal(d.goal.cpu().numpy(), cd[tag + "_goal"])
    dd = d.to(cuda_device).to_dict()
    assert set(dd) == {"goal", "v", "labels", "edge_index"}


def test_batched_ragged_vs_oracle(cuda_device):
    from gnn_motion_planning_b200 import graph
    from oracle import knn_graph as o_knn
    rng = np.random.default_rng(5)
    sizes = [(257, 257, 7, 2), (64, 40, 70, 3), (1000, 1000, 50, 2), (1, 1, 3, 7), (333, 100, 12, 14), (2, 2, 1, 2)]
    vs, node_ptr, nf, k1 = [], [0], [], []
    c = 7
    for n, f, k, _ in sizes:
        vs.append(rng.uniform(-2, 2, (n, c)).astype(np.float32))
        node_ptr.append(node_ptr[-1] + n); nf.append(f); k1.append(k)
    v = np.concatenate(vs)
    ei, ep = graph.knn_graph_batch(torch.from_numpy(v).to(cuda_device), node_ptr, nf, k1)
    # the same call with edge_ptr posted into pinned host memory by a kernel (gmp_post_to_host) instead of a copy-engine read-back
    pin = torch.full((len(sizes) + 1,), -1, dtype=torch.int32).pin_memory()
    ei2, ep2 = graph.knn_graph_batch(torch.from_numpy(v).to(cuda_device), node_ptr, nf, k1, edge_ptr_host=pin)
    assert np.array_equal(ep, ep2) and np.array_equal(pin.numpy(), ep) and torch.equal(ei[:, :ep[-1]], ei2[:, :ep[-1]])
    with pytest.raises(ValueError):
        graph.knn_graph_batch(torch.from_numpy(v).to(cuda_device), node_ptr, nf, k1, edge_ptr_host=torch.zeros(len(sizes) + 1, dtype=torch.int32))
    ei = ei.cpu().numpy()
    for g, (n, f, k, _) in enumerate(sizes):
        want = o_knn.knn_graph_edges(vs[g], f, k)
        got = ei[:, ep[g]:ep[g + 1]]
        assert got.shape == want.shape, (g, got.shape, want.shape)
        assert np.array_equal(got, want), g


@pytest.mark.parametrize("c,n,k", [(2, 1000, 75), (7, 1000, 75), (14, 2000, 83)])
def test_full_size_properties(cuda_device, c, n, k):
    """BASELINE sizes: sorted+unique, symmetric, self loops, every node has >= k1 neighbours; equals the oracle."""
    from gnn_motion_planning_b200 import graph
    from oracle import knn_graph as o_knn
    rng = np.random.default_rng(11)
    v = rng.uniform(-1, 1, (n, c)).astype(np.float32)
    ei = graph.knn_graph_edges(torch.from_numpy(v).to(cuda_device), n, k).cpu().numpy()
    key = ei[0] * n + ei[1]
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(np.sort(ei[1] * n + ei[0]), key)
    assert np.all(np.isin(np.arange(n) * (n + 1), key))
    assert np.bincount(ei[1], minlength=n).min() >= k
    assert np.array_equal(ei, o_knn.knn_graph_edges(v, n, k))


def test_errors(cuda_device):
    from gnn_motion_planning_b200 import _lib, graph
    v = torch.zeros(10, 2, device=cuda_device)
    with pytest.raises(_lib.GnnmpError):
        graph.knn_graph_batch(v, [0, 10], [11], [3])       # n_free > n
    with pytest.raises(_lib.GnnmpError):
        graph.knn_graph_batch(torch.zeros(10, 2), [0, 10], [10], [3])   # CPU tensor: no fallback
