"""GPU: BASELINE config C1 -- the reference planner loop explore(batch=100, t_max=100, k=10, smoother='none') on real
maze problems, driven through the drop-in surface (MazeEnv, create_data, EncoderProcessDecoder), against the
end-to-end golden produced by the reference's own explore()/MazeEnv/model.py with the same NumPy seed."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_explore_matches_reference(cuda_device):
    from gnn_motion_planning_b200.environment import MazeEnv
    from gnn_motion_planning_b200.eval_gnn import explore
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    gold = np.load(os.path.join(G, "explore_c1.npz"))
    mp = np.load(os.path.join(G, "maze_problems.npz"))
    env = MazeEnv(dim=2, map_file=os.path.join(G, "maze_problems.npz"))
    model = EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2).to(cuda_device)
    model.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu"))
    model.eval()
    checked = 0
    for pid in gold["ids"]:
        where = np.flatnonzero(mp["ids"] == pid)
        if len(where) == 0:
            continue
        np.random.seed(1234 + int(pid))
        env.init_new_problem(int(where[0]))
        r = explore(env, model, None, smooth=True, batch=100, t_max=100, k=10, smoother="none")
        assert r["success"] == bool(gold["p%d_success" % pid])
        assert len(r["data"].v) == int(gold["p%d_n_nodes" % pid])          # same RNG stream, same rejections
        assert r["explored"] == list(gold["p%d_explored" % pid])           # same edge order => same logits ranking
        assert r["c_explore"] == int(gold["p%d_c_explore" % pid])          # same collision_check_count
        assert np.allclose(np.array(r["path"]), gold["p%d_path" % pid])
        checked += 1
    assert checked >= 2


def test_explore_maze3_matches_reference(cuda_device):
    """The 3-D stick maze (str2name 'maze3': explorer (3, 32, 2), weights_maze_3.pt; MazeEnv(dim=3) on the gmp_maze3_* kernels)
    through the planner loop, smoother='none' (the reference ships no smoother weights for it), against the reference's own run:
    multi-round problems included (three graphs of up to 604 nodes), two of the four problems fail in the reference as well."""
    from gnn_motion_planning_b200.environment import MazeEnv
    from gnn_motion_planning_b200.eval_gnn import explore
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    gold = np.load(os.path.join(G, "explore_maze3.npz"))
    m3 = np.load(os.path.join(G, "maze3_collision.npz"))
    env = MazeEnv(dim=3, map_file=os.path.join(G, "maze3_collision.npz"))
    assert str(env) == "maze3" and env.bound == (-1, -1, -0.4, 1, 1, 0.4)
    model = EncoderProcessDecoder(workspace_size=2, config_size=3, embed_size=32, obs_size=2).to(cuda_device)
    model.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze_3.pt"), map_location="cpu"))
    for pid in gold["ids"]:
        np.random.seed(31 + int(pid))
        env.init_new_problem(int(np.flatnonzero(m3["ids"] == pid)[0]))
        r = explore(env, model, None, smooth=True, batch=100, t_max=300, k=10, smoother="none")
        assert r["success"] == bool(gold["p%d_success" % pid]), pid
        assert len(r["data"].v) == int(gold["p%d_n_nodes" % pid]), pid
        assert r["explored"] == list(gold["p%d_explored" % pid]), pid
        assert r["c_explore"] == int(gold["p%d_c_explore" % pid]), pid
        if r["success"]:
            assert np.allclose(np.array(r["path"]), gold["p%d_path" % pid])
