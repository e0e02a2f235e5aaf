"""GPU: BASELINE-size checks through size-independent properties: at N=1000, k=50 the oracle is too
slow to run per graph for a whole batch, so the batched CUDA path is checked by (i) batch invariance -- a graph scores
bit-identically alone and inside a 24-graph packed batch, (ii) dense == sparse scatter, (iii) node-permutation
equivariance, (iv) one full-size graph against the oracle at the 1e-4 gate, (v) collision symmetry + check-count
bounds over every edge of the batch, (vi) the planner-facing factory."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _batch(dev, B=24, n=1000, k=50, c=2):
    from gnn_motion_planning_b200 import graph
    maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    vs, obss = [], []
    for g in range(B):
        rng = np.random.default_rng(1234 + g)
        vs.append(rng.uniform(-1, 1, (n, c)).astype(np.float32))
        obss.append((np.argwhere(maps[g] == 1) / 15.0 - 0.5).astype(np.float32))
    v = torch.from_numpy(np.concatenate(vs)).to(dev)
    node_ptr = (np.arange(B + 1) * n).astype(np.int32)
    ei, edge_ptr = graph.knn_graph_batch(v, node_ptr, np.full(B, n), np.full(B, k))
    obs = torch.from_numpy(np.concatenate(obss)).to(dev)
    obs_ptr = np.cumsum([0] + [len(o) for o in obss]).astype(np.int32)
    goal = torch.from_numpy(np.stack([x[1] for x in vs])).to(dev)
    return vs, obss, v, ei, edge_ptr, node_ptr, obs, obs_ptr, goal, maps


def test_c2_batch_properties(cuda_device):
    from gnn_motion_planning_b200 import collision
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    from oracle import explorer as o_explorer
    dev = cuda_device
    sd = torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu")
    m = EncoderProcessDecoder(2, 2, 32, 2).to(dev)
    m.load_state_dict(sd)
    vs, obss, v, ei, edge_ptr, node_ptr, obs, obs_ptr, goal, maps = _batch(dev)
    et = int(edge_ptr[-1])
    logits = m.forward_batch(v, ei, goal, obs, node_ptr, edge_ptr, obs_ptr, loop=5).clone()
    assert torch.isfinite(logits[:et]).all()
    # (i) batch invariance + (ii) dense == sparse, on three graphs of the batch
    for g in (0, 7, 23):
        e0, e1 = int(edge_ptr[g]), int(edge_ptr[g + 1])
        eg = ei[:, e0:e1].contiguous()
        vg = v[node_ptr[g]:node_ptr[g + 1]]
        og = obs[obs_ptr[g]:obs_ptr[g + 1]]
        alone = m.forward_sparse(goal=goal[g], loop=5, v=vg, obstacles=og, edge_index=eg)
        assert torch.equal(alone, logits[e0:e1]), g
        dense = m(goal=goal[g], loop=5, v=vg, obstacles=og, edge_index=eg)
        assert torch.equal(dense[eg[1], eg[0]], alone)
        assert int((dense != 0).sum()) <= e1 - e0
    # (iii) node-permutation equivariance on one full-size graph
    g = 3
    e0, e1 = int(edge_ptr[g]), int(edge_ptr[g + 1])
    eg = ei[:, e0:e1]
    n = 1000
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).to(dev)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=dev)
    vg = v[node_ptr[g]:node_ptr[g + 1]]
    og = obs[obs_ptr[g]:obs_ptr[g + 1]]
    a = m.forward_sparse(goal=goal[g], loop=5, v=vg, obstacles=og, edge_index=eg.contiguous())
    b = m.forward_sparse(goal=goal[g], loop=5, v=vg[perm].contiguous(), obstacles=og, edge_index=inv[eg].contiguous())
    assert float((a - b).abs().max()) < 1e-4
    # (iv) one full-size graph against the oracle
    want = o_explorer.explorer_forward(sd, torch.from_numpy(vs[g]), eg.cpu(), torch.from_numpy(vs[g][1]), torch.from_numpy(obss[g]),
                                       loop=5, dense=False)
    assert float((a.cpu() - want).abs().max()) < 1e-4
    # (v) collision over every edge of the batch: symmetric, counts within [0 | 1 | 2 .. 2+127]
    maps_d = torch.from_numpy(maps).to(dev)
    node_ptr_d, edge_ptr_d = torch.from_numpy(node_ptr).to(dev), torch.from_numpy(edge_ptr).to(dev)
    free, checks = collision.maze_edge_fp_graph(v, ei, node_ptr_d, edge_ptr_d, maps_d, et, want_checks=True)
    flipped = torch.stack([ei[1, :et], ei[0, :et]]).contiguous()
    free_r, _ = collision.maze_edge_fp_graph(v, flipped, node_ptr_d, edge_ptr_d, maps_d, et, want_checks=True)
    assert torch.equal(free, free_r)                                   # _edge_fp(a,b) == _edge_fp(b,a) in the 2-D maze
    assert int(checks.min()) >= 1 and int(checks.max()) <= 129
    self_loops = ei[0, :et] == ei[1, :et]
    gid = torch.repeat_interleave(torch.arange(len(edge_ptr) - 1, device=dev), torch.from_numpy(np.diff(edge_ptr)).to(dev))
    pts = v[(ei[0, :et] + node_ptr_d[gid])[self_loops]]
    st = collision.maze_state_fp(pts, maps_d, gid[self_loops].to(torch.int32))
    assert torch.equal(st, free[self_loops])                            # a self loop is free iff its endpoint is
    rows = collision.result_rows(logits, free, edge_ptr_d, 100)
    assert torch.equal(rows[:, 0].cpu(), torch.arange(100, 124, dtype=torch.float32))
    assert torch.equal(rows[:, 1].cpu(), torch.from_numpy(np.diff(edge_ptr).astype(np.float32)))
    for g in (0, 11):
        e0, e1 = int(edge_ptr[g]), int(edge_ptr[g + 1])
        assert float(rows[g, 2]) == float(free[e0:e1].sum()) and float(rows[g, 3]) == float(logits[e0:e1].max())


def test_str2name_factory(cuda_device):
    from gnn_motion_planning_b200.str2name import TABLE, str2name
    for name, (ws, c, e, s, *_rest) in TABLE.items():
        env, model, mp, model_s, sp = str2name(name, make_env=False)
        assert env is None and model.config_size == c and model.embed_size == e and model.obs_size == s
        assert model_s.config_size == c and model_s.embed_size == 128
        assert mp.startswith("data/weights/") and sp.startswith("data/weights/")
    assert str2name("ur5", make_env=False)[3].scale == pytest.approx(2 * np.pi)
    assert str2name("maze3", make_env=False)[1].config_size == 3                  # runnable with smoother='none' (round 2)
    with pytest.raises(KeyError):
        str2name("maze4", make_env=False)


def test_hotpath_submit_wait_matches_compute(cuda_device):
    """The public batched API: pinned-host submit/wait (double buffered, overlapped read-back) returns exactly what the
    device-resident compute produces, batch after batch."""
    from gnn_motion_planning_b200.batch import HotPath
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    dev = cuda_device
    m = EncoderProcessDecoder(2, 2, 32, 2).to(dev)
    m.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu"))
    maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    B, n, k = 6, 300, 12
    hp = HotPath(m, B, n, k, kind="maze", maps=torch.from_numpy(maps).to(dev), first_problem_id=40, device=dev)
    tickets, wants = [], []
    for it in range(3):
        rng = np.random.default_rng(it)
        v = rng.uniform(-1, 1, (B * n, 2)).astype(np.float32)
        obss = [(np.argwhere(maps[(it * B + g) % len(maps)] == 1) / 15.0 - 0.5).astype(np.float32) for g in range(B)]
        obs_ptr = np.cumsum([0] + [len(o) for o in obss]).astype(np.int32)
        v_h = torch.from_numpy(v).pin_memory()
        goal_h = torch.from_numpy(v.reshape(B, n, 2)[:, 1].copy()).pin_memory()
        obs_h = torch.from_numpy(np.concatenate(obss)).pin_memory()
        prob_h = torch.from_numpy(((it * B + np.arange(B)) % len(maps)).astype(np.int32)).pin_memory()
        bufs = hp.compute(v_h.to(dev), goal_h.to(dev), obs_h.to(dev), obs_ptr, prob_h.to(dev), bufs=hp._alloc_set())
        et = bufs["et"]
        wants.append((bufs["edge_ptr"].copy(), bufs["ei"][:, :et].cpu(), bufs["logits"][:et].cpu(), bufs["free"][:et].cpu(), bufs["rows"].cpu()))
        tickets.append(hp.submit(v_h, goal_h, obs_h, obs_ptr, prob_h))
        if it == 1:   # results of an older batch stay valid while newer ones are in flight (two buffer sets)
            r0 = HotPath.wait(tickets[0])
            assert torch.equal(r0["logits"], wants[0][2])
    with pytest.raises(RuntimeError):                                     # the ticket of batch 0 was consumed: its buffers are reused
        HotPath.wait(tickets[0])
    for t, (ep, ei, lg, fr, rows) in list(zip(tickets, wants))[1:]:
        r = HotPath.wait(t)
        assert r["edge_index"].dtype == torch.int16                       # local ids travel narrow (N <= 32767): 4 B/edge, not 16
        assert np.array_equal(r["edge_ptr"], ep) and torch.equal(r["edge_index"].to(torch.int64), ei)
        assert torch.equal(r["logits"], lg) and torch.equal(r["free"], fr) and torch.equal(r["rows"], rows)
        assert float(rows[0, 0]) == 40.0


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] / configs[3] sizes: ONE graph of each against the oracle at the 1e-4 gate
KUKA_LIM = np.array([2.967, 2.094, 2.967, 2.094, 2.967, 2.094, 3.054])      # kuka_iiwa/model_0.urdf joint limits


def _one_graph_vs_oracle(dev, wfile, dims, n, k, lim, boxes, seed):
    from gnn_motion_planning_b200 import graph
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    from oracle import explorer as o_explorer
    from oracle import knn_graph as o_knn
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    m = EncoderProcessDecoder(*dims).to(dev)
    m.load_state_dict(sd)
    rng = np.random.default_rng(seed)
    v = rng.uniform(-lim, lim, (n, len(lim))).astype(np.float32)
    ei_want = o_knn.knn_graph_edges(v, n, k)
    vd = torch.from_numpy(v).to(dev)
    ei = graph.knn_graph_edges(vd, n, k)
    assert np.array_equal(ei.cpu().numpy(), ei_want)                               # bit-exact edge indices at full size
    obs = torch.from_numpy(boxes.astype(np.float32))
    got = {}
    for mode in ("tc", "simt"):
        m.set_edge_feature_mode(mode)
        got[mode] = m.forward_sparse(goal=vd[1], loop=5, v=vd, obstacles=obs.to(dev), edge_index=ei).cpu()
    want = o_explorer.explorer_forward(sd, torch.from_numpy(v), torch.from_numpy(ei_want), torch.from_numpy(v[1]), obs, loop=5,
                                       dense=False)
    for mode in ("tc", "simt"):
        err = float((got[mode] - want).abs().max())
        assert err < 1e-4, (mode, err)
    return ei_want, got


def test_c3_size_graph_vs_oracle(cuda_device):
    """C3: kuka7, embed 64 (phase-split tcgen05 stage), N=1000, k=50, the boxes of a real kukas_7 problem."""
    arm = np.load(os.path.join(G, "arm_problems.npz"))
    bp = arm["kuka7_box_ptr"]
    g = int(np.argmax(np.diff(bp)))                                                # the problem with the most boxes
    ei, _ = _one_graph_vs_oracle(cuda_device, "weights_kuka.pt", (3, 7, 64, 6), 1000, 50, KUKA_LIM,
                                 arm["kuka7_boxes"][bp[g]:bp[g + 1]], 1234)
    assert ei.shape[1] > 55000


def test_c4_size_graph_vs_oracle(cuda_device):
    """C4: kuka14, N=2000, k=50: hub rows (in-degree > 200) go through the segmented max of the message kernel."""
    arm = np.load(os.path.join(G, "arm_problems.npz"))
    bp = arm["kuka14_box_ptr"]
    ei, _ = _one_graph_vs_oracle(cuda_device, "kuka_14.pt", (3, 14, 32, 6), 2000, 50, np.concatenate([KUKA_LIM, KUKA_LIM]),
                                 arm["kuka14_boxes"][bp[3]:bp[4]], 1234)
    assert ei.shape[1] > 120000 and int(np.bincount(ei[1]).max()) > 200
