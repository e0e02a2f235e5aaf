"""GPU: maze collision kernels vs golden vectors (reference's own NumPy code) and vs the C oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def data():
    return np.load(os.path.join(G, "maze_collision.npz")), np.load(os.path.join(G, "maze_problems.npz"))


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_state_fp_golden(cuda_device, data, t):
    from gnn_motion_planning_b200 import collision
    mc, mp = data
    maps = torch.from_numpy(mp["maps"]).to(cuda_device)
    free, counted = collision.maze_state_fp(torch.from_numpy(mc["states_" + t]).to(cuda_device), maps,
                                            torch.from_numpy(mc["state_problem_" + t]).to(cuda_device), want_counted=True)
    assert np.array_equal(free.cpu().numpy(), mc["state_free_" + t])
    assert np.array_equal(counted.cpu().numpy(), mc["state_counted_" + t])


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_edge_fp_golden(cuda_device, data, t):
    from gnn_motion_planning_b200 import collision
    mc, mp = data
    maps = torch.from_numpy(mp["maps"]).to(cuda_device)
    free, checks = collision.maze_edge_fp(torch.from_numpy(mc["edge_a_" + t]).to(cuda_device),
                                          torch.from_numpy(mc["edge_b_" + t]).to(cuda_device), maps,
                                          torch.from_numpy(mc["edge_problem_" + t]).to(cuda_device), want_checks=True)
    assert np.array_equal(free.cpu().numpy(), mc["edge_free_" + t])
    assert np.array_equal(checks.cpu().numpy(), mc["edge_checks_" + t])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_edge_fp_vs_oracle_1M(cuda_device, dt):
    """>= 10^6 random edges (incl. out-of-range endpoints), booleans and check counts bit-exact vs the C oracle."""
    from gnn_motion_planning_b200 import collision
    from oracle import maze as o_maze
    maps_np = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    rng = np.random.default_rng(7)
    n = 1 << 20
    a = rng.uniform(-1.02, 1.02, (n, 2)).astype(dt)
    b = (a + rng.normal(0, 0.25, (n, 2))).astype(dt)
    prob = rng.integers(0, len(maps_np), n).astype(np.int32)
    free, checks = collision.maze_edge_fp(torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device),
                                          torch.from_numpy(maps_np).to(cuda_device), torch.from_numpy(prob).to(cuda_device),
                                          want_checks=True)
    of, oc = o_maze.edge_fp(a, b, maps_np, prob)
    assert np.array_equal(free.cpu().numpy(), of)
    assert np.array_equal(checks.cpu().numpy(), oc)
    assert 0.05 < of.mean() < 0.95


def test_edge_fp_graph_matches_explicit(cuda_device):
    from gnn_motion_planning_b200 import collision
    maps_np = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    rng = np.random.default_rng(3)
    B, n = 5, 200
    v = rng.uniform(-1, 1, (B * n, 2)).astype(np.float32)
    es = [rng.integers(0, n, (2, 700 + 13 * g)) for g in range(B)]
    edge_ptr = np.cumsum([0] + [e.shape[1] for e in es]).astype(np.int32)
    node_ptr = (np.arange(B + 1) * n).astype(np.int32)
    ei = np.concatenate(es, 1).astype(np.int64)
    vd = torch.from_numpy(v).to(cuda_device)
    pg = torch.tensor([3, 1, 4, 1, 5], dtype=torch.int32, device=cuda_device)
    free, checks = collision.maze_edge_fp_graph(vd, torch.from_numpy(ei).to(cuda_device), torch.from_numpy(node_ptr).to(cuda_device),
                                                torch.from_numpy(edge_ptr).to(cuda_device), torch.from_numpy(maps_np).to(cuda_device),
                                                int(edge_ptr[-1]), problem_of_graph=pg, want_checks=True)
    gid = np.repeat(np.arange(B), np.diff(edge_ptr))
    a = v[ei[0] + node_ptr[gid]]
    b = v[ei[1] + node_ptr[gid]]
    prob = pg.cpu().numpy()[gid]
    f2, c2 = collision.maze_edge_fp(torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device),
                                    torch.from_numpy(maps_np).to(cuda_device), torch.from_numpy(prob).to(cuda_device), want_checks=True)
    assert torch.equal(free, f2) and torch.equal(checks, c2)


def test_empty_and_bad_dtype(cuda_device):
    from gnn_motion_planning_b200 import collision
    maps = torch.zeros(1, 15, 15, dtype=torch.uint8, device=cuda_device)
    e = torch.zeros(0, 2, device=cuda_device)
    assert collision.maze_edge_fp(e, e, maps).shape == (0,)
    with pytest.raises(TypeError):
        collision.maze_state_fp(torch.zeros(4, 2, dtype=torch.float16, device=cuda_device), maps)


def test_maze_env_protocol(cuda_device):
    """Drop-in env: scalar _state_fp/_edge_fp + collision_check_count side effects equal the golden (reference) values."""
    from gnn_motion_planning_b200.environment import MazeEnv
    mc, mp = np.load(os.path.join(G, "maze_collision.npz")), np.load(os.path.join(G, "maze_problems.npz"))
    env = MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz"))
    assert str(env) == "maze2"
    env.init_new_problem(6)
    sel = np.flatnonzero(mc["edge_problem_f32"] == 6)[:60]
    c0 = env.collision_check_count
    for i in sel:
        before = env.collision_check_count
        got = env._edge_fp(mc["edge_a_f32"][i], mc["edge_b_f32"][i])
        assert got == bool(mc["edge_free_f32"][i])
        assert env.collision_check_count - before == mc["edge_checks_f32"][i]
        if got:
            assert env.k == mc["edge_k_f32"][i]
    assert env.collision_check_count - c0 == mc["edge_checks_f32"][sel].sum()
    got = env.edge_fp_batch(mc["edge_a_f64"][sel], mc["edge_b_f64"][sel])
    sel64 = sel
    assert np.array_equal(got, mc["edge_free_f64"][sel64].astype(bool))
    # obstacles tokens as the reference builds them (maze_env.py:73-79)
    occ = np.argwhere(mp["maps"][6] == 1)
    assert np.allclose(env.obstacles, occ / 15 - 0.5)


def test_construct_graph_golden(cuda_device):
    """algorithm/dijkstra.py:15-31 (k-NN(5) graph + EVERY edge checked) against the reference's own construct_graph +
    MazeEnv run on float64 points: edge list, free flags, per-node neighbour / cost lists and collision_check_count."""
    from gnn_motion_planning_b200.algorithm import construct_graph
    from gnn_motion_planning_b200.environment import MazeEnv
    gold = np.load(os.path.join(G, "construct_graph.npz"))
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    env = MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz"))
    for pid in gold["ids"]:
        tag = "p%d" % pid
        env.init_new_problem(int(np.flatnonzero(mp["ids"] == pid)[0]))
        pts = gold[tag + "_points"]
        c0 = env.collision_check_count
        edge_cost, neighbors, edge_index, edge_free = construct_graph(env, pts)
        assert np.array_equal(edge_index, gold[tag + "_edge_index"]) and edge_index.dtype == np.int64
        assert np.array_equal(np.array(edge_free), gold[tag + "_edge_free"])
        assert env.collision_check_count - c0 == int(gold[tag + "_checks"])
        assert np.array_equal(np.array([len(neighbors[i]) for i in range(len(pts))]), gold[tag + "_deg"])
        assert np.array_equal(np.concatenate([np.asarray(neighbors[i], np.int64) for i in range(len(pts))]), gold[tag + "_nbr_flat"])
        assert np.array_equal(np.concatenate([np.asarray(edge_cost[i], np.float64) for i in range(len(pts))]), gold[tag + "_cost_flat"])


@pytest.mark.parametrize("t", ["f32", "f64"])
def test_maze3_stick_golden(cuda_device, t):
    """gmp_maze3_state_fp / gmp_maze3_edge_fp against the reference's own MazeEnv(dim=3): booleans, check counts and env.k."""
    from gnn_motion_planning_b200 import collision
    g = np.load(os.path.join(G, "maze3_collision.npz"))
    dev = cuda_device
    maps = torch.from_numpy(g["maps"]).to(dev)
    f, c, k = collision.maze3_state_fp(torch.from_numpy(g["states_" + t]).to(dev), maps, torch.from_numpy(g["state_problem_" + t]).to(dev))
    assert np.array_equal(f.cpu().numpy(), g["state_free_" + t])
    assert np.array_equal(c.cpu().numpy(), g["state_checks_" + t]) and np.array_equal(k.cpu().numpy(), g["state_k_" + t])
    f, c, k = collision.maze3_edge_fp(torch.from_numpy(g["edge_a_" + t]).to(dev), torch.from_numpy(g["edge_b_" + t]).to(dev), maps,
                                      torch.from_numpy(g["edge_problem_" + t]).to(dev))
    assert np.array_equal(f.cpu().numpy(), g["edge_free_" + t])
    assert np.array_equal(c.cpu().numpy(), g["edge_checks_" + t]) and np.array_equal(k.cpu().numpy(), g["edge_k_" + t])


def test_maze3_stick_vs_oracle_large(cuda_device):
    """200 000 random edges / states per dtype against the C oracle (device cos / sin vs libm: the stick end points may differ in
    the last bit, which can only matter on a cell boundary -- none may show up here)."""
    from gnn_motion_planning_b200 import collision
    from oracle import maze as o_maze
    g = np.load(os.path.join(G, "maze3_collision.npz"))
    rng = np.random.default_rng(12)
    n = 200000
    for dt in (np.float32, np.float64):
        a = np.concatenate([rng.uniform(-1, 1, (n, 2)), rng.uniform(-0.4, 0.4, (n, 1))], 1).astype(dt)
        b = (a + np.concatenate([rng.normal(0, 0.08, (n, 2)), rng.normal(0, 0.15, (n, 1))], 1)).astype(dt)
        b[:, 2] = np.clip(b[:, 2], -0.4, 0.4)
        pr = rng.integers(0, len(g["maps"]), n).astype(np.int32)
        dev = cuda_device
        maps = torch.from_numpy(g["maps"]).to(dev)
        f, c, k = collision.maze3_edge_fp(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), maps, torch.from_numpy(pr).to(dev))
        of, oc, ok = o_maze.stick_edge_fp(a, b, g["maps"], pr)
        assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(c.cpu().numpy(), oc) and np.array_equal(k.cpu().numpy(), ok)
        f, c, k = collision.maze3_state_fp(torch.from_numpy(a).to(dev), maps, torch.from_numpy(pr).to(dev))
        of, oc, ok = o_maze.stick_state_fp(a, g["maps"], pr)
        assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(c.cpu().numpy(), oc)
