"""GPU: smoother forward through the C ABI vs golden (reference model_smoother.py) and vs the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4

CASES = [("2d", "smooth_2d_attv3.pt", 2), ("7d", "smooth_7d_attv3.pt", 7), ("2d_short", "smooth_2d_attv3.pt", 2),
         ("14d", "smooth_14d_attv3.pt", 14), ("13d", "smooth_13d_attv3.pt", 13), ("ur5", "smooth_ur5_attv3.pt", 6),
         ("snake", "smooth_snake_attv3.pt", 7)]
MORE = ("14d", "13d", "ur5", "snake")                # fixtures of round 2: tests/golden/smoother_more.npz (ur5: scale = 2*pi)


def make(wfile, c, dev, scale=1.0):
    from gnn_motion_planning_b200.model_smoother import ModelSmoother
    m = ModelSmoother(workspace_size=3, config_size=c, embed_size=128, obs_size=6, scale=scale).to(dev)
    m.load_state_dict(torch.load(os.path.join(G, "weights", wfile), map_location="cpu"))
    return m.eval()


@pytest.mark.parametrize("tag,wfile,c", CASES)
def test_forward_golden(cuda_device, tag, wfile, c):
    sm = np.load(os.path.join(G, "smoother_more.npz" if tag in MORE else "smoother.npz"))
    m = make(wfile, c, cuda_device, scale=float(sm[tag + "_scale"]) if tag in MORE else 1.0)
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


# ---------------------------------------------------------------------------------------------------------------
# a5: the CALLER of the smoother forward -- reference smoother.py:194-246 -- against fixtures recorded while
# the reference's own model_smooth / proposed_path_smootherv2 / MazeEnv / model_smoother.py ran (make_golden.more_goldens)
def _a5_env():
    from gnn_motion_planning_b200.environment import MazeEnv
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    return MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz")), mp


def test_steering_rounds_golden(cuda_device):
    """proposed_path_smootherv2 (smoother.py:194-216) has no model in it: same inputs => the same path bit for bit and
    the same collision_check_count, for every steering call the reference made (5 per problem)."""
    from gnn_motion_planning_b200.smoother import proposed_path_smootherv2
    gold = np.load(os.path.join(G, "model_smooth.npz"))
    env, mp = _a5_env()
    n = 0
    for pid in gold["ids"]:
        env.init_new_problem(int(np.flatnonzero(mp["ids"] == pid)[0]))
        for j in range(int(gold["p%d_n_steer" % pid])):
            old, new = gold["p%d_steer%d_old" % (pid, j)], gold["p%d_steer%d_new" % (pid, j)]
            c0 = env.collision_check_count
            out = proposed_path_smootherv2(list(old), list(new), env)
            assert np.array_equal(np.array(out), gold["p%d_steer%d_out" % (pid, j)]), (pid, j)
            assert env.collision_check_count - c0 == int(gold["p%d_steer%d_checks" % (pid, j)]), (pid, j)
            n += 1
    assert n == 30


def test_model_smooth_golden(cuda_device):
    """model_smooth (smoother.py:233-246): 5 x (smoother forward on the GPU -> steering rounds).  The forward is within 1e-4 of
    the reference's, the steering thresholds (dist < RRT_EPS, collision booleans) amplify nothing on these problems: the
    accepted path matches to 2e-4 and the collision-check count exactly."""
    from gnn_motion_planning_b200.smoother import model_smooth
    gold = np.load(os.path.join(G, "model_smooth.npz"))
    env, mp = _a5_env()
    m = make("smooth_2d_attv3.pt", 2, cuda_device)
    for pid in gold["ids"]:
        env.init_new_problem(int(np.flatnonzero(mp["ids"] == pid)[0]))
        free = list(gold["p%d_ms_free" % pid])
        coll = list(gold["p%d_ms_collided" % pid])
        path = list(gold["p%d_ms_path" % pid])
        c0 = env.collision_check_count
        out = np.array(model_smooth(m, free, coll, path, env))
        want = gold["p%d_ms_out" % pid]
        assert out.shape == want.shape
        assert np.abs(out - want).max() < 2e-4, (pid, np.abs(out - want).max())
        assert env.collision_check_count - c0 == int(gold["p%d_ms_checks" % pid]), pid


def test_explore_with_model_smoother_golden(cuda_device):
    """explore(..., smoother='model') end to end (eval_gnn.py:168-276 -> smoother.py:233): success, explore / smooth check
    counts and the smoothed path against the reference run with the same NumPy seed."""
    from gnn_motion_planning_b200.eval_gnn import explore, path_cost
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    gold = np.load(os.path.join(G, "model_smooth.npz"))
    env, mp = _a5_env()
    model = EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2).to(cuda_device)
    model.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu"))
    ms = make("smooth_2d_attv3.pt", 2, cuda_device)
    for pid in gold["ids"]:
        np.random.seed(4321 + int(pid))
        env.init_new_problem(int(np.flatnonzero(mp["ids"] == pid)[0]))
        r = explore(env, model, ms, smooth=True, batch=100, t_max=100, k=10, smoother="model")
        assert r["success"] == bool(gold["p%d_success" % pid])
        assert r["c_explore"] == int(gold["p%d_c_explore" % pid])
        assert np.allclose(np.array(r["path"]), gold["p%d_path" % pid])
        assert r["c_smooth"] == int(gold["p%d_c_smooth" % pid]), pid
        assert np.abs(np.array(r["smooth_path"]) - gold["p%d_smooth_path" % pid]).max() < 2e-4
        assert path_cost(r["smooth_path"]) <= path_cost(r["path"]) + 1e-6


def test_steering_rounds_on_device_golden(cuda_device):
    """gmp_maze_steer_rounds: all 30 steering calls of the reference run as ONE batch -- paths bit for bit,
    collision_check_count exact, path cost as eval_gnn.path_cost gives it."""
    from gnn_motion_planning_b200.eval_gnn import path_cost
    from gnn_motion_planning_b200.smoother import steer_rounds_batch
    gold = np.load(os.path.join(G, "model_smooth.npz"))
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    olds, news, outs, checks, probs = [], [], [], [], []
    for pid in gold["ids"]:
        for j in range(int(gold["p%d_n_steer" % pid])):
            olds.append(gold["p%d_steer%d_old" % (pid, j)]); news.append(gold["p%d_steer%d_new" % (pid, j)])
            outs.append(gold["p%d_steer%d_out" % (pid, j)]); checks.append(int(gold["p%d_steer%d_checks" % (pid, j)]))
            probs.append(int(np.flatnonzero(mp["ids"] == pid)[0]))
    assert all(o.dtype == np.float32 for o in olds)
    ptr = np.concatenate([[0], np.cumsum([len(o) for o in olds])]).astype(np.int32)
    dev = cuda_device
    out, chk, rounds, cost = steer_rounds_batch(torch.from_numpy(np.concatenate(olds)).to(dev), torch.from_numpy(np.concatenate(news)).to(dev),
                                                torch.from_numpy(ptr).to(dev), torch.from_numpy(mp["maps"]).to(dev),
                                                torch.tensor(probs, dtype=torch.int32, device=dev), 0.05, want_cost=True)
    out, chk, cost = out.cpu().numpy(), chk.cpu().numpy(), cost.cpu().numpy()
    for i, want in enumerate(outs):
        assert np.array_equal(out[ptr[i]:ptr[i + 1]], want), i
        assert chk[i] == checks[i], i
        assert abs(cost[i] - path_cost(list(want))) < 1e-6
    assert len(outs) == 30


def test_model_smooth_batch_golden(cuda_device):
    """model_smooth for all six problems in one batch, paths device-resident across the five iterations."""
    from gnn_motion_planning_b200.smoother import model_smooth_batch
    gold = np.load(os.path.join(G, "model_smooth.npz"))
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    m = make("smooth_2d_attv3.pt", 2, cuda_device)
    pids = [int(p) for p in gold["ids"]]
    paths, checks, cost = model_smooth_batch(m, [gold["p%d_ms_free" % p] for p in pids], [gold["p%d_ms_collided" % p] for p in pids],
                                             [gold["p%d_ms_path" % p] for p in pids], torch.from_numpy(mp["maps"]).to(cuda_device),
                                             [int(np.flatnonzero(mp["ids"] == p)[0]) for p in pids])
    for i, p in enumerate(pids):
        want = gold["p%d_ms_out" % p]
        assert np.abs(paths[i] - want).max() < 2e-4, (p, np.abs(paths[i] - want).max())
        assert int(checks[i]) == int(gold["p%d_ms_checks" % p]), p
