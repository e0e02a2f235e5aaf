"""CPU experiment: would a 3xTF32 (hi/lo split) tensor-core formulation of the explorer's dense layers stay inside the
1e-4 logit gate?  Emulates every Linear / attention product with operands rounded to TF32 (10 explicit mantissa bits)
and fp32 accumulation, in four variants: 1xTF32, 3xTF32 (hi*hi + hi*lo + lo*hi), 2xTF32 on the weight products (the
weights' lo plane dropped), and plain fp32, each compared with the
fp64 evaluation of the same graph.   python tools/tf32_feasibility.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import explorer as ox, knn_graph as ok  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
MODE = {"m": "fp32"}


def tf32(x):
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)            # round-to-nearest (ties away) on the 13 dropped bits
    return r.view(torch.float32)


def mm(a, b, b_is_weight=False):
    if a.dtype != torch.float32 or MODE["m"] == "fp32":
        return a @ b
    ah, bh = tf32(a), tf32(b)
    if MODE["m"] == "tf32x1":
        return ah @ bh
    al, bl = tf32(a - ah), tf32(b - bh)
    if MODE["m"] == "tf32x2w" and b_is_weight:   # drop A_hi . B_lo where B is a weight plane (two MMAs per product)
        return al @ bh + ah @ bh
    return (al @ bh + ah @ bl) + ah @ bh      # small terms first


def _lin(x, sd, name, bias=True):
    y = mm(x, sd[name + ".weight"].t(), b_is_weight=True)
    return y + sd[name + ".bias"] if bias else y


ox._lin = _lin


def _attention(map_code, obs_code, sd, name, embed):
    mv = _lin(map_code, sd, name + ".value", False); ov = _lin(obs_code, sd, name + ".value", False)
    q = _lin(map_code, sd, name + ".query", False); k = _lin(map_code, sd, name + ".key", False); okk = _lin(obs_code, sd, name + ".key", False)
    att = torch.cat(((q * k).sum(-1, keepdim=True), mm(q, okk.t())), -1) / embed ** 0.5
    att = att.softmax(-1)
    new = att[:, :1] * mv + mm(att[:, 1:], ov)
    return ox._layer_norm(new + map_code, sd, name + ".layer_norm")


ox._attention = _attention

sd = torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu")
maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
torch.set_num_threads(8)
for g in range(2):
    rng = np.random.default_rng(100 + g)
    v = rng.uniform(-1, 1, (1000, 2)).astype(np.float32)
    ei = torch.from_numpy(ok.knn_graph_edges(v, 1000, 50)); vt = torch.from_numpy(v)
    obs = torch.from_numpy((np.argwhere(maps[g] == 1) / 15.0 - 0.5).astype(np.float32))
    MODE["m"] = "fp32"
    f64 = ox.explorer_forward(sd, vt, ei, vt[1], obs, loop=5, dense=False, dtype=torch.float64)
    for mode in ("fp32", "tf32x3", "tf32x2w", "tf32x1"):
        MODE["m"] = mode
        got = ox.explorer_forward(sd, vt, ei, vt[1], obs, loop=5, dense=False).double()
        print("graph %d  %-7s max|logit - fp64| = %.3e   (|logit|max %.1f)" % (g, mode, float((got - f64).abs().max()), float(f64.abs().max())))
