"""GPU: explorer forward through the C ABI vs golden (reference model.py) and vs the oracle; 1e-4 on logits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4  # BASELINE.json north_star: "edge logits within 1e-4 fp32"

# all seven (config, embed, obs) combinations of reference str2name.py:12-66 == model.SUPPORTED_DIMS
CASES = [("maze2", "weights_maze.pt", (2, 2, 32, 2)), ("kuka7", "weights_kuka.pt", (3, 7, 64, 6)),
         ("kuka14", "kuka_14.pt", (3, 14, 32, 6)), ("snake7", "weights_snake.pt", (3, 7, 32, 2)),
         ("ur5", "weights_ur5.pt", (3, 6, 32, 6)), ("kuka13", "weights_kuka_13.pt", (3, 13, 32, 6)),
         ("maze3", "weights_maze_3.pt", (2, 3, 32, 2))]
MORE = ("snake7", "ur5", "kuka13", "maze3")          # fixtures of round 2: tests/golden/explorer_more.npz


def test_cases_cover_every_supported_instantiation():
    from gnn_motion_planning_b200.model import SUPPORTED_DIMS
    assert {(d[1], d[2], d[3]) for _, _, d in CASES} == set(SUPPORTED_DIMS)


def make_model(wfile, dims, dev):
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    m = EncoderProcessDecoder(workspace_size=dims[0], config_size=dims[1], embed_size=dims[2], obs_size=dims[3]).to(dev)
    m.load_state_dict(torch.load(os.path.join(G, "weights", wfile), map_location="cpu"))
    return m.eval()


@pytest.mark.parametrize("tag,wfile,dims", CASES)
def test_forward_golden(cuda_device, tag, wfile, dims):
    ex = np.load(os.path.join(G, "explorer_more.npz" if tag in MORE else "explorer.npz"))
    m = make_model(wfile, dims, cuda_device)
    v = torch.from_numpy(ex[tag + "_v"]).to(cuda_device)
    ei = torch.from_numpy(ex[tag + "_edge_index"]).to(cuda_device)
    obs = torch.from_numpy(ex[tag + "_obstacles"]).to(cuda_device)
    goal = torch.from_numpy(ex[tag + "_goal"]).to(cuda_device)
    for loop in (1, 5):
        dense = m(goal=goal, loop=loop, v=v, obstacles=obs, free=None, collided=None, edge_index=ei, labels=None)
        assert dense.shape == (len(v), len(v)) and dense.dtype == torch.float32 and dense.is_cuda
        got = dense[ei[1], ei[0]].cpu().numpy()
        want = ex["%s_logits_loop%d" % (tag, loop)]
        assert np.abs(got - want).max() < TOL, (tag, loop, np.abs(got - want).max())
        mask = torch.zeros_like(dense, dtype=torch.bool)
        mask[ei[1], ei[0]] = True
        assert float(dense[~mask].abs().max()) == 0.0                  # zeros off the edge set (model.py:148)
        sp = m.forward_sparse(goal=goal, loop=loop, v=v, obstacles=obs, edge_index=ei)
        assert torch.equal(sp, dense[ei[1], ei[0]])
    m.use_obstacles = False                                            # poked from outside, eval_gnn.py:88
    got = m.forward_sparse(goal=goal, loop=5, v=v, obstacles=obs, edge_index=ei).cpu().numpy()
    assert np.abs(got - ex[tag + "_logits_noobs"]).max() < TOL


@pytest.mark.parametrize("tag,wfile,dims", CASES)
def test_batched_ragged_vs_oracle(cuda_device, tag, wfile, dims):
    """A ragged packed batch (different N, E, O per graph, one graph without obstacles, unsorted non-symmetric edges)."""
    from oracle import explorer as o_explorer
    from oracle import knn_graph as o_knn
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    m = make_model(wfile, dims, cuda_device)
    c, s = dims[1], dims[3]
    rng = np.random.default_rng(42)
    graphs = []
    for n, k, o in [(300, 12, 70 if s == 2 else 9), (40, 5, 0), (513, 9, 33 if s == 2 else 1), (129, 20, 5)]:
        v = rng.uniform(-1, 1, (n, c)).astype(np.float32)
        ei = o_knn.knn_graph_edges(v, n, k)
        if n == 129:   # arbitrary COO: shuffled order, a few edges dropped (not symmetric)
            keep = rng.permutation(ei.shape[1])[: ei.shape[1] - 17]
            ei = ei[:, keep]
        obs = rng.uniform(-0.5, 0.5, (o, s)).astype(np.float32)
        graphs.append((v, ei, v[1].copy(), obs))
    node_ptr = np.cumsum([0] + [len(g[0]) for g in graphs])
    edge_ptr = np.cumsum([0] + [g[1].shape[1] for g in graphs])
    obs_ptr = np.cumsum([0] + [len(g[3]) for g in graphs])
    V = torch.from_numpy(np.concatenate([g[0] for g in graphs])).to(cuda_device)
    EI = torch.from_numpy(np.concatenate([g[1] for g in graphs], 1)).to(cuda_device)
    GOAL = torch.from_numpy(np.stack([g[2] for g in graphs])).to(cuda_device)
    OBS = torch.from_numpy(np.concatenate([g[3] for g in graphs])).to(cuda_device)
    logits, dense = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5, dense=True)
    logits = logits.cpu().numpy()
    dense = dense.cpu().numpy()
    doff = np.cumsum([0] + [len(g[0]) ** 2 for g in graphs])
    for g, (v, ei, goal, obs) in enumerate(graphs):
        want = o_explorer.explorer_forward(sd, torch.from_numpy(v), torch.from_numpy(ei), torch.from_numpy(goal),
                                           torch.from_numpy(obs), loop=5, dense=False).numpy()
        got = logits[edge_ptr[g]:edge_ptr[g + 1]]
        assert np.abs(got - want).max() < TOL, (tag, g, np.abs(got - want).max())
        d = dense[doff[g]:doff[g + 1]].reshape(len(v), len(v))
        assert np.array_equal(d[ei[1], ei[0]], got)
        assert np.count_nonzero(d) <= ei.shape[1]


def test_error_budget_vs_fp64(cuda_device):
    """The kernel's error against the fp64 arbiter is of the same order as the reference fp32 path's own error."""
    from oracle import explorer as o_explorer
    ex = np.load(os.path.join(G, "explorer.npz"))
    sd = torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu")
    m = make_model("weights_maze.pt", (2, 2, 32, 2), cuda_device)
    v, ei = torch.from_numpy(ex["maze2_v"]), torch.from_numpy(ex["maze2_edge_index"])
    obs, goal = torch.from_numpy(ex["maze2_obstacles"]), torch.from_numpy(ex["maze2_goal"])
    f64 = o_explorer.explorer_forward(sd, v, ei, goal, obs, loop=5, dense=False, dtype=torch.float64).numpy()
    got = m.forward_sparse(goal=goal.to(cuda_device), loop=5, v=v.to(cuda_device), obstacles=obs.to(cuda_device),
                           edge_index=ei.to(cuda_device)).cpu().numpy()
    ref_err = np.abs(ex["maze2_logits_loop5"] - f64).max()
    our_err = np.abs(got - f64).max()
    print("max |err| vs fp64: reference fp32 %.3g, kernel %.3g" % (ref_err, our_err))
    assert our_err < 5e-5


def test_errors(cuda_device):
    from gnn_motion_planning_b200 import _lib
    m = make_model("weights_maze.pt", (2, 2, 32, 2), cuda_device)
    v = torch.zeros(10, 2, device=cuda_device)
    ei = torch.tensor([[0, 1], [1, 12]], device=cuda_device)
    with pytest.raises(IndexError):
        m(goal=v[0], loop=1, v=v, obstacles=torch.zeros(3, 2, device=cuda_device), edge_index=ei)
    with pytest.raises(_lib.GnnmpError):
        m.forward_batch(torch.zeros(10, 2), ei.cpu(), v[:1], None, [0, 10], [0, 2], [0, 0])
    # the batched entry point cannot raise without a device round trip: bad ids are clamped (memory safe) and counted
    out = m.forward_batch(v, ei, v[:1].contiguous(), torch.zeros(3, 2, device=cuda_device), [0, 10], [0, 2], [0, 3])
    assert torch.isfinite(out).all() and m.last_bad_edges() == 1
    good = torch.tensor([[0, 1], [1, 2]], device=cuda_device)
    m.forward_batch(v, good, v[:1].contiguous(), torch.zeros(3, 2, device=cuda_device), [0, 10], [0, 2], [0, 3])
    assert m.last_bad_edges() == 0
    # empty edge set / single node
    out = m(goal=v[0], loop=2, v=v[:1], obstacles=torch.zeros(0, 2, device=cuda_device),
            edge_index=torch.zeros(2, 0, dtype=torch.int64, device=cuda_device))
    assert out.shape == (1, 1) and float(out.abs().sum()) == 0.0


@pytest.mark.parametrize("tag,wfile,dims", CASES)
def test_tensor_core_edge_stage_vs_simt_and_oracle(cuda_device, tag, wfile, dims):
    """The tcgen05 3xTF32 kernels (default) against the fp32 SIMT kernels and the oracle on a ragged batch.  embed 32: obstacle
    counts cover 0, one partial chunk, exactly 96, two chunks (97, 130) and three chunks (200).  embed 64 (kuka7): the
    phase-split stage handles up to 32 obstacles per graph (0, 1, 5, 12, 17, 31, 32 here)."""
    from oracle import explorer as o_explorer
    from oracle import knn_graph as o_knn
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    m = make_model(wfile, dims, cuda_device)
    c, s = dims[1], dims[3]
    rng = np.random.default_rng(7)
    graphs = []
    obs_counts = [96, 0, 97, 5, 130, 200, 57] if dims[2] == 32 else [32, 0, 17, 5, 1, 31, 12]
    for (n, k), o in zip([(300, 12), (40, 5), (513, 9), (129, 20), (260, 8), (64, 6), (700, 10)], obs_counts):
        v = rng.uniform(-1, 1, (n, c)).astype(np.float32)
        ei = o_knn.knn_graph_edges(v, n, k)
        obs = rng.uniform(-0.5, 0.5, (o, s)).astype(np.float32)
        graphs.append((v, ei, v[1].copy(), obs))
    node_ptr = np.cumsum([0] + [len(g[0]) for g in graphs])
    edge_ptr = np.cumsum([0] + [g[1].shape[1] for g in graphs])
    obs_ptr = np.cumsum([0] + [len(g[3]) for g in graphs])
    V = torch.from_numpy(np.concatenate([g[0] for g in graphs])).to(cuda_device)
    EI = torch.from_numpy(np.concatenate([g[1] for g in graphs], 1)).to(cuda_device)
    GOAL = torch.from_numpy(np.stack([g[2] for g in graphs])).to(cuda_device)
    OBS = torch.from_numpy(np.concatenate([g[3] for g in graphs])).to(cuda_device)
    out = {}
    modes = ("tc", "simt", "tc4") if dims[2] == 32 else ("tc", "simt")
    for mode in modes:
        m.set_edge_feature_mode(mode)
        out[mode] = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5).cpu().numpy()
    m.use_obstacles = False
    m.set_edge_feature_mode("tc")
    noobs_tc = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5).cpu().numpy()
    m.set_edge_feature_mode("simt")
    noobs_simt = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5).cpu().numpy()
    assert np.abs(noobs_tc - noobs_simt).max() < TOL
    worst = 0.0
    for g, (v, ei, goal, obs) in enumerate(graphs):
        want = o_explorer.explorer_forward(sd, torch.from_numpy(v), torch.from_numpy(ei), torch.from_numpy(goal),
                                           torch.from_numpy(obs), loop=5, dense=False, dtype=torch.float64).numpy()
        for mode in modes:
            err = np.abs(out[mode][edge_ptr[g]:edge_ptr[g + 1]] - want).max()
            worst = max(worst, err)
            assert err < TOL, (tag, g, mode, err)
    print("tc vs simt max |diff| %.3g; worst |err| vs fp64 oracle %.3g" % (np.abs(out["tc"] - out["simt"]).max(), worst))
    assert np.abs(out["tc"] - out["simt"]).max() < TOL


def test_forward_is_stable_under_concurrent_streams(cuda_device):
    """The tensor-core kernels synchronise through mbarrier phases; a protocol slip shows up as a hang (trap) or a wrong
    result only when warps are delayed.  Run the forward repeatedly while a second stream keeps the SMs busy with k-NN
    graph builds: every repetition must reproduce the first result bit for bit."""
    from gnn_motion_planning_b200 import graph
    from oracle import knn_graph as o_knn
    m = make_model("weights_maze.pt", (2, 2, 32, 2), cuda_device)
    rng = np.random.default_rng(3)
    B, n, k = 12, 700, 20
    vs, eis, obs = [], [], []
    for g in range(B):
        v = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        vs.append(v)
        eis.append(o_knn.knn_graph_edges(v, n, k))
        obs.append(rng.uniform(-0.5, 0.5, (60 + 7 * g, 2)).astype(np.float32))     # 60 .. 137 obstacles: lone chunks and pairs
    node_ptr = np.arange(B + 1) * n
    edge_ptr = np.cumsum([0] + [e.shape[1] for e in eis])
    obs_ptr = np.cumsum([0] + [len(o) for o in obs])
    V = torch.from_numpy(np.concatenate(vs)).to(cuda_device)
    EI = torch.from_numpy(np.concatenate(eis, 1)).to(cuda_device)
    GOAL = torch.from_numpy(np.stack([v[1] for v in vs])).to(cuda_device)
    OBS = torch.from_numpy(np.concatenate(obs)).to(cuda_device)
    first = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5).clone()
    side = torch.cuda.Stream(device=cuda_device)
    npt = np.arange(B + 1, dtype=np.int32) * n
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    for rep in range(12):
        out = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5)     # asynchronous on the current stream
        with torch.cuda.stream(side):                                                   # ... while the side stream builds graphs
            for _ in range(3):
                graph.knn_graph_batch(V, npt, np.full(B, n, np.int32), np.full(B, k, np.int32))
        assert torch.equal(out, first), rep
    torch.cuda.synchronize()


@pytest.mark.parametrize("wfile,dims,lockstep", [("weights_maze.pt", (2, 2, 32, 2), "tc"), ("weights_maze_3.pt", (2, 3, 32, 2), "tc")])
def test_ready_driven_issuer_is_bit_identical(cuda_device, wfile, dims, lockstep):
    """mode "tcrd" (what auto picks when every graph has 1..128 obstacles): the edge-feature kernel with ONE ISSUER WARP PER TILE
    computes exactly what the lockstep issuer (modes "tc" / "tc4": tile 0 then tile 1, stage by stage) computes -- same MMAs, same
    epilogues, another interleaving -- on ragged batches (57, 96, 97, 128 obstacles cover lone table chunks and pairs), repeatedly
    and under a concurrent stream.  (Eight-warps-per-tile organisation only: wide-input models keep the lockstep issuer.)"""
    from gnn_motion_planning_b200 import graph
    from oracle import knn_graph as o_knn
    m = make_model(wfile, dims, cuda_device)
    c, s_obs = dims[1], dims[3]
    rng = np.random.default_rng(21)
    vs, eis, obs = [], [], []
    for (n, k), o in zip([(700, 14), (300, 12), (513, 9), (129, 20), (900, 10), (64, 6)], [57, 96, 97, 128, 1, 70]):
        v = rng.uniform(-1, 1, (n, c)).astype(np.float32)
        vs.append(v)
        eis.append(o_knn.knn_graph_edges(v, n, k))
        obs.append(rng.uniform(-0.5, 0.5, (o, s_obs)).astype(np.float32))
    node_ptr = np.cumsum([0] + [len(v) for v in vs])
    edge_ptr = np.cumsum([0] + [e.shape[1] for e in eis])
    obs_ptr = np.cumsum([0] + [len(o) for o in obs])
    V = torch.from_numpy(np.concatenate(vs)).to(cuda_device)
    EI = torch.from_numpy(np.concatenate(eis, 1)).to(cuda_device)
    GOAL = torch.from_numpy(np.stack([v[1] for v in vs])).to(cuda_device)
    OBS = torch.from_numpy(np.concatenate(obs)).to(cuda_device)
    m.set_edge_feature_mode(lockstep)
    want = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5).clone()
    m.set_edge_feature_mode("tcrd")
    side = torch.cuda.Stream(device=cuda_device)
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    npt = node_ptr.astype(np.int32)
    for rep in range(8):
        got = m.forward_batch(V, EI, GOAL, OBS, node_ptr, edge_ptr, obs_ptr, loop=5)
        with torch.cuda.stream(side):
            graph.knn_graph_batch(V, npt, np.diff(npt).astype(np.int32), np.full(len(vs), 8, np.int32))
        assert torch.equal(got, want), rep
    torch.cuda.synchronize()
