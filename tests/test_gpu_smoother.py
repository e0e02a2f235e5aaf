"""GPU: smoother forward through the C ABI vs golden (reference model_smoother.py) and vs the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4

CASES = [("2d", "smooth_2d_attv3.pt", 2), ("7d", "smooth_7d_attv3.pt", 7), ("2d_short", "smooth_2d_attv3.pt", 2)]


def make(wfile, c, dev, scale=1.0):
    from gnn_motion_planning_b200.model_smoother import ModelSmoother
    m = ModelSmoother(workspace_size=3, config_size=c, embed_size=128, obs_size=6, scale=scale).to(dev)
    m.load_state_dict(torch.load(os.path.join(G, "weights", wfile), map_location="cpu"))
    return m.eval()


@pytest.mark.parametrize("tag,wfile,c", CASES)
def test_forward_golden(cuda_device, tag, wfile, c):
    sm = np.load(os.path.join(G, "smoother.npz"))
    m = make(wfile, c, cuda_device)
    path = torch.from_numpy(sm[tag + "_path"]).to(cuda_device)
    keep = path.clone()
    for loop in (1, 3):
        out = m(path=path, free=torch.from_numpy(sm[tag + "_free"]).to(cuda_device),
                collided=torch.from_numpy(sm[tag + "_collided"]).to(cuda_device), obstacles=None,
                edge_index=torch.from_numpy(sm[tag + "_edge_index"]).to(cuda_device), loop=loop)
        want = sm["%s_out_loop%d" % (tag, loop)]
        assert out.shape == want.shape and out.is_cuda
        assert np.abs(out.cpu().numpy() - want).max() < TOL, (tag, loop, np.abs(out.cpu().numpy() - want).max())
        assert torch.equal(out[0], path[0]) and torch.equal(out[-1], path[-1])      # endpoints pinned (model_smoother.py:139)
    assert torch.equal(path, keep)                                                  # caller's path not mutated


def test_batched_vs_oracle_with_scale_and_duplicates(cuda_device):
    """Ragged batch; scale != 1 (UR5, str2name.py:40); duplicate + sample-sourced caller edges (coalesce dedups)."""
    from oracle import smoother as o_sm
    sd = torch.load(os.path.join(G, "weights", "smooth_7d_attv3.pt"), map_location="cpu")
    scale = 2 * np.pi
    m = make("smooth_7d_attv3.pt", 7, cuda_device, scale=scale)
    rng = np.random.default_rng(9)
    probs = []
    for p, f, c_ in [(12, 300, 200), (3, 5, 0), (40, 500, 500), (7, 9, 4)]:
        path = np.cumsum(rng.uniform(-0.2, 0.3, (p, 7)), 0).astype(np.float32)
        free = rng.uniform(-3, 3, (f, 7)).astype(np.float32)
        coll = rng.uniform(-3, 3, (c_, 7)).astype(np.float32)
        e = o_sm.chain_edge_index(p).numpy()
        if p == 12:   # duplicates and an explicit sample -> path edge
            e = np.concatenate([e, e[:, :5], np.array([[p + 3], [2]])], 1)
        probs.append((path, free, coll, e))
    path_ptr = np.cumsum([0] + [len(x[0]) for x in probs])
    sample_ptr = np.cumsum([0] + [len(x[1]) + len(x[2]) for x in probs])
    edge_ptr = np.cumsum([0] + [x[3].shape[1] for x in probs])
    out = m.forward_batch(torch.from_numpy(np.concatenate([x[0] for x in probs])).to(cuda_device),
                          torch.from_numpy(np.concatenate([np.concatenate([x[1], x[2]]) for x in probs])).to(cuda_device),
                          torch.from_numpy(np.concatenate([x[3] for x in probs], 1)).to(cuda_device),
                          path_ptr, sample_ptr, [len(x[1]) for x in probs], edge_ptr, loop=2).cpu().numpy()
    for g, (path, free, coll, e) in enumerate(probs):
        want = o_sm.smoother_forward(sd, torch.from_numpy(path), torch.from_numpy(free), torch.from_numpy(coll).reshape(-1, 7),
                                     torch.from_numpy(e), loop=2, scale=scale).numpy()
        got = out[path_ptr[g]:path_ptr[g + 1]]
        assert np.abs(got - want).max() < TOL, (g, np.abs(got - want).max())


def test_model_smooth_host_loop(cuda_device):
    """smoother.model_smooth (the caller, smoother.py:233-246) runs end to end on the drop-in env + model."""
    from gnn_motion_planning_b200.environment import MazeEnv
    from gnn_motion_planning_b200.smoother import model_smooth
    env = MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz"))
    env.init_new_problem(6)
    np.random.seed(0)
    free, coll = env.sample_n_points(120, need_negative=True)
    m = make("smooth_2d_attv3.pt", 2, cuda_device)
    # a collision-free straight segment chopped into waypoints
    a = free[0]
    b = next(f for f in free[1:] if env._edge_fp(a, f) and np.linalg.norm(a - f) > 0.2)
    path = [a + (b - a) * t for t in np.linspace(0, 1, 6)]
    c0 = env.collision_check_count
    new = model_smooth(m, list(free), list(coll), path, env)
    assert len(new) == len(path) and np.allclose(new[0], path[0]) and np.allclose(new[-1], path[-1])
    assert env.collision_check_count > c0
    for p, q in zip(new[:-1], new[1:]):
        assert env._edge_fp(np.asarray(p), np.asarray(q))
