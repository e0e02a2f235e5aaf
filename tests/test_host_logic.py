"""CPU: host-side logic of the batched planner loop that needs no GPU -- the read-ahead RNG stream (must hand out exactly the
draws np.random.uniform would, one state at a time, like maze_env.py:127-135), the k1 rule of create_data (eval_gnn.py:159), the
chain graph of model_smooth (smoother.py:238-241) and the reference's explored-edge mask quirk as the search kernel reads it."""
import numpy as np
import torch

from gnn_motion_planning_b200 import graph, smoother
from gnn_motion_planning_b200.environment.env_config import LIMITS
from gnn_motion_planning_b200.search import _Stream


def test_stream_equals_sequential_numpy_draws():
    seed = 1234
    np.random.seed(seed)
    want = np.array([np.random.uniform(-LIMITS[:2], LIMITS[:2], (1, 2)).reshape(-1) for _ in range(500)])   # uniform_sample(), one call per draw
    s = _Stream(seed)
    got = []
    for n_peek, n_use in ((64, 10), (64, 64), (7, 3), (250, 123), (400, 300)):   # peeks larger than what is consumed: nothing is lost
        d = s.peek(n_peek)
        assert len(d) == n_peek
        got.append(d[:n_use].copy())
        s.consume(n_use)
    got = np.concatenate(got)
    assert np.array_equal(got, want[:len(got)])


def test_k1_rule():
    # k1 = ceil(k * ln(len(free)) / ln(100)) in float64 (eval_gnn.py:159): 502 free samples, k = 30 -> 41
    assert graph.k1_of(30, 502) == 41
    assert graph.k1_of(10, 102) == int(np.ceil(10 * np.log(102) / np.log(100)))
    assert graph.k1_of(50, 100) == 50


def test_chain_edge_index_matches_reference_construction():
    # smoother.py:238-241: (i+1 -> i), (i -> i+1), then add_self_loops appends (i, i)
    p = 5
    e = torch.cat((torch.arange(1, p).reshape(1, -1), torch.arange(0, p - 1).reshape(1, -1)), dim=0)
    e = torch.cat((e, e.flip(0)), dim=-1)
    loops = torch.arange(p)
    want = torch.cat((e, torch.stack((loops, loops))), dim=-1)
    assert torch.equal(smoother.chain_edge_index(p), want)
    assert smoother.chain_edge_index(1).shape == (2, 1)


def test_explored_edge_mask_quirk_semantics():
    """eval_gnn.py:202 `policy[np.array(explored_edges).reshape(2, -1)] = 0` under the author's torch: with the flat list
    L = [0,0, a1,b1,b1,a1, ...] and M = len(L)/2 the zeroed entries are (L[i], L[M+i]) -- what gmp_maze_tree_search applies."""
    explored_edges = [[0, 0], [0, 5], [5, 0], [5, 7], [7, 5]]
    ee = np.array(explored_edges).reshape(2, -1)
    flat = np.array(explored_edges).reshape(-1)
    m = len(flat) // 2
    assert np.array_equal(ee[0], flat[:m]) and np.array_equal(ee[1], flat[m:])
    pairs = set(zip(ee[0].tolist(), ee[1].tolist()))
    assert pairs == {(0, 0), (0, 5), (0, 7), (5, 7), (5, 5)}           # (0,7), (5,5) were never explored; (5,0), (7,5) were
    assert (5, 0) not in pairs and (7, 5) not in pairs                 # -- the quirk the device search reproduces
