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
