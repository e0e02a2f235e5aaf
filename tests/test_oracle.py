"""CPU: the oracle restatements against the golden vectors produced by the reference's own python
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import explorer, knn_graph, maze, smoother

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def maze_golden():
    return np.load(os.path.join(G, "maze_collision.npz")), np.load(os.path.join(G, "maze_problems.npz"))


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_maze_state_oracle_matches_reference(maze_golden, t):
    mc, mp = maze_golden
    free, counted = maze.state_fp(mc["states_" + t], mp["maps"], mc["state_problem_" + t])
    assert np.array_equal(free, mc["state_free_" + t])
    assert np.array_equal(counted, mc["state_counted_" + t])


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_maze_edge_oracle_matches_reference(maze_golden, t):
    mc, mp = maze_golden
    free, checks = maze.edge_fp(mc["edge_a_" + t], mc["edge_b_" + t], mp["maps"], mc["edge_problem_" + t])
    assert np.array_equal(free, mc["edge_free_" + t])
    assert np.array_equal(checks, mc["edge_checks_" + t])          # collision_check_count increments
    ok = mc["edge_free_" + t] == 1
    assert np.array_equal((checks - 2)[ok], mc["edge_k_" + t][ok])  # env.k = midpoints, on accepted edges


def test_maze_edge_symmetry_property():
    """_edge_fp(a,b) == _edge_fp(b,a) in the 2-D maze: midpoint arithmetic is commutative in IEEE."""
    rng = np.random.default_rng(0)
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    a = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    b = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    prob = rng.integers(0, len(mp["maps"]), 4000).astype(np.int32)
    f1, _ = maze.edge_fp(a, b, mp["maps"], prob)
    f2, _ = maze.edge_fp(b, a, mp["maps"], prob)
    assert np.array_equal(f1, f2)


def test_maze_empty():
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    f, c = maze.edge_fp(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), mp["maps"])
    assert f.shape == (0,) and c.shape == (0,)


@pytest.mark.parametrize("tag", ["maze2", "kuka7", "kuka14", "dup"])
def test_create_data_oracle_matches_reference(tag):
    cd = np.load(os.path.join(G, "create_data.npz"))
    d = knn_graph.create_data(cd[tag + "_free"], cd[tag + "_collided"], cd[tag + "_free"][1], int(cd[tag + "_k"]))
    assert np.array_equal(d["edge_index"], cd[tag + "_edge_index"])
    assert np.array_equal(d["v"], cd[tag + "_v"])
    assert np.array_equal(d["labels"], cd[tag + "_labels"])
    assert np.array_equal(d["goal"], cd[tag + "_goal"])


def test_knn_graph_properties():
    rng = np.random.default_rng(1)
    v = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    ei = knn_graph.knn_graph_edges(v, 200, 9)
    key = ei[0] * 300 + ei[1]
    assert np.all(np.diff(key) > 0)                                    # sorted, unique (coalesce idempotent)
    assert set(map(tuple, ei.T)) == set(map(tuple, ei[::-1].T))        # symmetric
    assert np.all(np.isin(np.arange(300) * 301, key))                  # self loops (loop=True)


# every (config, embed, obs) combination of reference str2name.py:12-66; the last four live in explorer_more.npz (round 2)
EXPLORER_CASES = [("maze2", "weights_maze.pt"), ("kuka7", "weights_kuka.pt"), ("kuka14", "kuka_14.pt"),
                  ("snake7", "weights_snake.pt"), ("ur5", "weights_ur5.pt"), ("kuka13", "weights_kuka_13.pt"),
                  ("maze3", "weights_maze_3.pt")]
EXPLORER_MORE = ("snake7", "ur5", "kuka13", "maze3")


@pytest.mark.parametrize("tag,wfile", EXPLORER_CASES)
def test_explorer_oracle_matches_reference(tag, wfile):
    ex = np.load(os.path.join(G, "explorer_more.npz" if tag in EXPLORER_MORE else "explorer.npz"))
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    v, ei = torch.from_numpy(ex[tag + "_v"]), torch.from_numpy(ex[tag + "_edge_index"])
    obs, goal = torch.from_numpy(ex[tag + "_obstacles"]), torch.from_numpy(ex[tag + "_goal"])
    for loop in (1, 5):
        got = explorer.explorer_forward(sd, v, ei, goal, obs, loop=loop, dense=False).numpy()
        assert np.abs(got - ex["%s_logits_loop%d" % (tag, loop)]).max() < 1e-4   # north_star tolerance
    got = explorer.explorer_forward(sd, v, ei, goal, obs, loop=5, dense=False, use_obstacles=False).numpy()
    assert np.abs(got - ex[tag + "_logits_noobs"]).max() < 1e-4
    dense = explorer.explorer_forward(sd, v, ei, goal, obs, loop=5, dense=True)
    assert dense.shape == (len(v), len(v)) and int((dense != 0).sum()) <= ei.shape[1]
    assert torch.equal(dense[ei[1], ei[0]], torch.from_numpy(
        explorer.explorer_forward(sd, v, ei, goal, obs, loop=5, dense=False).numpy()))


def test_explorer_permutation_equivariance():
    """Relabelling the nodes permutes the logits (no dependence on node order beyond the goal arg-min)."""
    ex = np.load(os.path.join(G, "explorer.npz"))
    sd = torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu")
    v, ei = torch.from_numpy(ex["maze2_v"]), torch.from_numpy(ex["maze2_edge_index"])
    obs, goal = torch.from_numpy(ex["maze2_obstacles"]), torch.from_numpy(ex["maze2_goal"])
    perm = torch.randperm(len(v), generator=torch.Generator().manual_seed(0))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(len(v))
    a = explorer.explorer_forward(sd, v, ei, goal, obs, loop=3, dense=True, dtype=torch.float64)
    b = explorer.explorer_forward(sd, v[perm], inv[ei], goal, obs, loop=3, dense=True, dtype=torch.float64)
    assert torch.allclose(a, b[inv][:, inv], atol=1e-9)


SMOOTHER_CASES = [("2d", "smooth_2d_attv3.pt"), ("7d", "smooth_7d_attv3.pt"), ("2d_short", "smooth_2d_attv3.pt"),
                  ("14d", "smooth_14d_attv3.pt"), ("13d", "smooth_13d_attv3.pt"), ("ur5", "smooth_ur5_attv3.pt"),
                  ("snake", "smooth_snake_attv3.pt")]
SMOOTHER_MORE = ("14d", "13d", "ur5", "snake")


@pytest.mark.parametrize("tag,wfile", SMOOTHER_CASES)
def test_smoother_oracle_matches_reference(tag, wfile):
    sm = np.load(os.path.join(G, "smoother_more.npz" if tag in SMOOTHER_MORE else "smoother.npz"))
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    scale = float(sm[tag + "_scale"]) if tag in SMOOTHER_MORE else 1.0          # ur5: 2*pi (str2name.py:40)
    for loop in (1, 3):
        got = smoother.smoother_forward(sd, torch.from_numpy(sm[tag + "_path"]), torch.from_numpy(sm[tag + "_free"]),
                                        torch.from_numpy(sm[tag + "_collided"]), torch.from_numpy(sm[tag + "_edge_index"]),
                                        loop=loop, scale=scale).numpy()
        want = sm["%s_out_loop%d" % (tag, loop)]
        assert np.abs(got - want).max() < 1e-5
        assert np.array_equal(got[0], sm[tag + "_path"][0]) and np.array_equal(got[-1], sm[tag + "_path"][-1])
    assert torch.equal(smoother.chain_edge_index(len(sm[tag + "_path"])), torch.from_numpy(sm[tag + "_edge_index"]))


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_maze3_stick_oracle_matches_reference(t):
    """3-D stick maze (MazeEnv(dim=3), maze_env.py:245-264,279-291,327-347): booleans, collision_check_count and env.k of the
    reference module itself, float32 and float64 states (theta wrap-around and out-of-range states included)."""
    g = np.load(os.path.join(G, "maze3_collision.npz"))
    f, c, k = maze.stick_state_fp(g["states_" + t], g["maps"], g["state_problem_" + t])
    assert np.array_equal(f, g["state_free_" + t]) and np.array_equal(c, g["state_checks_" + t]) and np.array_equal(k, g["state_k_" + t])
    f, c, k = maze.stick_edge_fp(g["edge_a_" + t], g["edge_b_" + t], g["maps"], g["edge_problem_" + t])
    assert np.array_equal(f, g["edge_free_" + t]) and np.array_equal(c, g["edge_checks_" + t]) and np.array_equal(k, g["edge_k_" + t])
    assert 0.3 < f.mean() < 0.7 and c.max() > 100
