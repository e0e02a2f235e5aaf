"""GPU: the batched planner loop with the lazy tree search on the device against what the reference's own
explore() / MazeEnv / model.py produced for the same seeds (tests/golden/make_golden.py): explore_c1.npz (BASELINE config C1,
one graph per problem) and explore_rounds.npz (small batches: several resampling rounds, the explored-edge quirk of
eval_gnn.py:202 carried from graph to graph).  Every problem of a fixture runs in ONE batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def setup(cuda_device):
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    model = EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2).to(cuda_device)
    model.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu"))
    return mp, model.eval()


def _rows(mp, pids):
    return [int(np.flatnonzero(mp["ids"] == p)[0]) for p in pids]


@pytest.mark.parametrize("spec_k", [1, 8])
def test_c1_batch_matches_reference(cuda_device, setup, spec_k):
    from gnn_motion_planning_b200.search import explore_batch, path_cost
    mp, model = setup
    gold = np.load(os.path.join(G, "explore_c1.npz"))
    pids = [int(p) for p in gold["ids"]]
    res = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], _rows(mp, pids), [1234 + p for p in pids],
                        batch=100, t_max=100, k=10, spec_k=spec_k, device=cuda_device)
    wasted = 0
    for pid, r in zip(pids, res):
        assert r["success"] == bool(gold["p%d_success" % pid]), pid
        assert r["n_nodes"] == int(gold["p%d_n_nodes" % pid]), pid                # same RNG stream, same rejections
        assert r["explored"] == list(gold["p%d_explored" % pid]), pid             # same search order
        assert r["c_explore"] == int(gold["p%d_c_explore" % pid]), pid            # same collision_check_count
        assert np.allclose(np.array(r["path"]), gold["p%d_path" % pid]), pid
        assert abs(r["path_cost"] - path_cost(gold["p%d_path" % pid])) < 1e-5     # the device-side result row (eval_gnn.py:120-122)
        assert r["row"][1] == 1.0 and r["row"][3] == r["c_search"] and r["row"][5] == len(r["explored"])
        assert r["rounds"] == 1
        wasted += r["spec_checks"]
    assert (wasted == 0) if spec_k == 1 else (wasted > 0)
    print("spec_k=%d: speculative checks never committed: %d (committed search checks %d)" % (spec_k, wasted, sum(r["c_search"] for r in res)))


@pytest.mark.parametrize("b,t,kk", [(25, 200, 6), (40, 300, 8)])
@pytest.mark.parametrize("spec_k", [1, 4])
def test_resampling_rounds_match_reference(cuda_device, setup, b, t, kk, spec_k):
    from gnn_motion_planning_b200.search import explore_batch
    mp, model = setup
    gold = np.load(os.path.join(G, "explore_rounds.npz"))
    pids = sorted({int(c[0]) for c in gold["cases"]})
    res = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], _rows(mp, pids), [777 + p for p in pids],
                        batch=b, t_max=t, k=kk, spec_k=spec_k, device=cuda_device)
    multi = 0
    for pid, r in zip(pids, res):
        tag = "p%d_b%d" % (pid, b)
        assert r["success"] == bool(gold[tag + "_success"]), tag
        assert r["n_nodes"] == int(gold[tag + "_n_nodes"]), tag
        assert r["explored"] == list(gold[tag + "_explored"]), tag
        assert r["c_explore"] == int(gold[tag + "_c_explore"]), tag
        if r["success"]:
            assert np.allclose(np.array(r["path"]), gold[tag + "_path"]), tag
        multi += r["rounds"] > 1
    assert multi >= 2                                                           # the fixture does exercise the carry-over


def test_batch_equals_host_loop(cuda_device, setup):
    """The device search against this repo's own host mirror of explore() (the reference's loop with GPU calls) on problems and
    seeds that are in no fixture."""
    from gnn_motion_planning_b200.environment import MazeEnv
    from gnn_motion_planning_b200.eval_gnn import explore
    from gnn_motion_planning_b200.search import explore_batch
    mp, model = setup
    rows = [0, 3, 5, 12, 13]
    res = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], rows, [99 + r for r in rows], batch=60, t_max=240, k=8,
                        spec_k=2, device=cuda_device)
    env = MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz"))
    for row, r in zip(rows, res):
        np.random.seed(99 + row)
        env.init_new_problem(row)
        h = explore(env, model, None, smooth=True, batch=60, t_max=240, k=8, smoother="none")
        assert r["success"] == h["success"] and r["explored"] == h["explored"] and r["c_explore"] == h["c_explore"], row
        if h["success"]:
            assert np.allclose(np.array(r["path"]), np.array(h["path"]))


def test_device_sampler_semantics(cuda_device, setup):
    """gmp_maze_sample_points: a counter-based stream with the reference's rejection-sampling semantics --
    every returned free state is free and every rejected one is not (bit-exact state checks), draws = free + rejected, the
    result is a PREFIX of one fixed sequence (sampling 40 then 60 more == sampling 100), streams differ, and the accepted
    fraction matches the free area of the map."""
    from gnn_motion_planning_b200 import collision
    mp, _ = setup
    dev = cuda_device
    maps_d = torch.from_numpy(mp["maps"]).to(dev)
    probs = [0, 3, 6, 6]
    streams = [11, 12, 13, 14]
    free, coll, n_coll, n_draws = collision.maze_sample_points(maps_d, probs, streams, 100, seed=7)
    free_h, coll_h, n_coll, n_draws = free.cpu().numpy(), coll.cpu().numpy(), n_coll.cpu().numpy(), n_draws.cpu().numpy()
    assert free_h.dtype == np.float64 and np.all(np.abs(free_h) <= 1.0)
    for i, p in enumerate(probs):
        assert n_draws[i] == 100 + n_coll[i]
        pr = torch.full((100,), p, dtype=torch.int32, device=dev)
        assert bool(collision.maze_state_fp(free[i], maps_d, pr).all())
        nc = int(n_coll[i])
        prc = torch.full((nc,), p, dtype=torch.int32, device=dev)
        assert not bool(collision.maze_state_fp(coll[i, :nc].contiguous(), maps_d, prc).any())
        area = float((mp["maps"][p] == 0).mean())
        assert abs(100.0 / n_draws[i] - area) < 0.15                                   # accepted fraction ~ free area
    assert not np.array_equal(free_h[2], free_h[3])                                   # same map, different streams
    # prefix property + continuation (first_draw)
    f40, _, c40, d40 = collision.maze_sample_points(maps_d, probs, streams, 40, seed=7)
    assert np.array_equal(f40.cpu().numpy(), free_h[:, :40])
    f60, _, c60, d60 = collision.maze_sample_points(maps_d, probs, streams, 60, seed=7, first_draw=d40)
    assert np.array_equal(f60.cpu().numpy(), free_h[:, 40:])
    assert np.array_equal((d40 + d60).cpu().numpy(), n_draws)


def test_explore_batch_with_device_sampler(cuda_device, setup):
    """The whole planner round on the device sampler: different samples than the NumPy stream, same planner quality band."""
    from gnn_motion_planning_b200.search import explore_batch
    mp, model = setup
    rows = list(range(len(mp["ids"])))
    a = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], rows, [500 + r for r in rows], batch=100, t_max=300, k=10,
                      device=cuda_device, sampler="device")
    b = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], rows, [500 + r for r in rows], batch=100, t_max=300, k=10,
                      device=cuda_device, sampler="device")
    assert [r["explored"] for r in a] == [r["explored"] for r in b]                    # deterministic
    n = explore_batch(model, mp["maps"], mp["init_states"], mp["goal_states"], rows, [500 + r for r in rows], batch=100, t_max=300, k=10,
                      device=cuda_device)
    assert sum(r["success"] for r in a) >= sum(r["success"] for r in n) - 2
    for r in a:
        if r["success"]:
            assert r["path_nodes"][0] == 0 and r["c_explore"] > 0


@pytest.mark.parametrize("tag,model_id,dims,wfile", [("kuka7", 0, (3, 7, 64, 6), "weights_kuka.pt"), ("kuka14", 1, (3, 14, 32, 6), "kuka_14.pt")])
def test_arm_tree_search_equals_host_loop(cuda_device, tag, model_id, dims, wfile):
    """gmp_arm_tree_search (the same search kernel instantiated on the arm edge check) against this repo's host mirror of the
    reference loop on KukaEnv / Kuka2Env problems: explored order, collision_check_count, path -- for spec_k = 1 and 4.  (Arm
    collision itself is pinned to oracle/arm.c, not to PyBullet.)"""
    from gnn_motion_planning_b200.environment import Kuka2Env, KukaEnv
    from gnn_motion_planning_b200.eval_gnn import explore
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    from gnn_motion_planning_b200.search import explore_batch_arm
    probs = np.load(os.path.join(G, "arm_problems.npz"))
    boxes, ptr = probs[tag + "_boxes"], probs[tag + "_box_ptr"]
    problems = []
    for i in range(6):
        obs = [(boxes[j, :3], boxes[j, 3:]) for j in range(ptr[i], ptr[i + 1])]
        problems.append((obs, probs[tag + "_start"][i], probs[tag + "_goal"][i], []))
    model = EncoderProcessDecoder(*dims).to(cuda_device)
    model.load_state_dict(torch.load(os.path.join(G, "weights", wfile), map_location="cpu"))
    seeds = [7000 + i for i in range(len(problems))]
    kw = dict(batch=60, t_max=180, k=8)
    res = {sk: explore_batch_arm(model, model_id, [p[:3] for p in problems], seeds, rrt_eps=0.5, spec_k=sk, device=cuda_device, **kw)
           for sk in (1, 4)}
    env = (KukaEnv if tag == "kuka7" else Kuka2Env)(problems=problems)
    n_ok = 0
    for i, seed in enumerate(seeds):
        np.random.seed(seed)
        env.init_new_problem(i)
        h = explore(env, model, None, smooth=True, smoother="none", **kw)
        for sk in (1, 4):
            r = res[sk][i]
            assert r["success"] == h["success"] and r["n_nodes"] == len(h["data"].v), (tag, i, sk)
            assert r["explored"] == h["explored"], (tag, i, sk)
            assert r["c_explore"] == h["c_explore"], (tag, i, sk, r["c_explore"], h["c_explore"])
            if h["success"]:
                assert np.allclose(np.array(r["path"]), np.array(h["path"]))
        n_ok += h["success"]
    assert res[1][0]["spec_checks"] == 0
    print("%s: %d / %d solved; uncommitted speculative checks at spec_k=4: %d" % (tag, n_ok, len(seeds), sum(r["spec_checks"] for r in res[4])))
