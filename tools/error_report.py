"""GPU box: report explorer logit error of the CUDA path vs the fp32 oracle and the fp64 arbiter at full graph size
(N=1000, k=50), for the shipped maze2 / kuka7 / kuka14 weights.  Used to judge precision-affecting kernel changes
against the 1e-4 gate.   python tools/error_report.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnn_motion_planning_b200.model import EncoderProcessDecoder  # noqa: E402
from oracle import explorer as o_explorer, knn_graph as o_knn  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda", 0)
maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
for tag, wfile, dims, n, k in (("maze2", "weights_maze.pt", (2, 2, 32, 2), 1000, 50), ("kuka7", "weights_kuka.pt", (3, 7, 64, 6), 1000, 50),
                               ("kuka14", "kuka_14.pt", (3, 14, 32, 6), 1000, 50)):
    sd = torch.load(os.path.join(G, "weights", wfile), map_location="cpu")
    m = EncoderProcessDecoder(*dims).to(dev)
    m.load_state_dict(sd)
    worst32 = worst64 = ref64 = 0.0
    for g in range(3):
        rng = np.random.default_rng(100 + g)
        lo, hi = (-1, 1) if tag == "maze2" else (-2.9, 2.9)
        v = rng.uniform(lo, hi, (n, dims[1])).astype(np.float32)
        ei = o_knn.knn_graph_edges(v, n, k)
        if tag == "maze2":
            obs = (np.argwhere(maps[g] == 1) / 15.0 - 0.5).astype(np.float32)
        else:
            obs = np.concatenate([rng.uniform(0.1, 0.3, (5, 3)), rng.uniform(-0.8, 0.8, (5, 3))], 1).astype(np.float32)
        vt, et, ot = torch.from_numpy(v), torch.from_numpy(ei), torch.from_numpy(obs)
        got = m.forward_sparse(goal=vt[1].to(dev), loop=5, v=vt.to(dev), obstacles=ot.to(dev), edge_index=et.to(dev)).cpu().double()
        f32 = o_explorer.explorer_forward(sd, vt, et, vt[1], ot, loop=5, dense=False).double()
        f64 = o_explorer.explorer_forward(sd, vt, et, vt[1], ot, loop=5, dense=False, dtype=torch.float64)
        worst32 = max(worst32, float((got - f32).abs().max()))
        worst64 = max(worst64, float((got - f64).abs().max()))
        ref64 = max(ref64, float((f32 - f64).abs().max()))
    print("%-7s N=%d E=%d |logit|max=%.1f : kernel-vs-fp32-oracle %.2e  kernel-vs-fp64 %.2e  fp32-oracle-vs-fp64 %.2e" % (
        tag, n, ei.shape[1], float(f64.abs().max()), worst32, worst64, ref64))
