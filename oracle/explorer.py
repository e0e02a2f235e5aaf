"""ORACLE (test infrastructure, never the product path): CPU restatement of the reference
GNN path explorer forward, ``EncoderProcessDecoder.forward`` -- reference ``model.py:115-150``
with ``MPNN`` (``model.py:22-41``), ``Attention`` (:153-181), ``FeedForward`` (:184-201) and
``Block`` (:204-218).

Plain torch tensor ops on the CPU, functional over the reference ``state_dict`` (the 200-tensor
dict of ``data/weights/weights_*.pt``; dead tensors are ignored).  ``dtype=torch.float64`` gives
the high-precision arbiter used to judge fp32 error budgets.

Pinning: ``tests/golden/make_golden.py`` runs the reference's own ``model.py`` (PyG primitives
stubbed per their published semantics, see ``tests/golden/_pyg_stubs.py``) on shipped weights and
commits its outputs; ``tests/test_oracle_explorer.py`` checks this file against those vectors.
PyG itself (torch_geometric / torch_scatter / torch_cluster, versions unpinned by the reference)
is not installable here, so parity is pinned to the reference's python code but UNPINNED at the
PyG-primitive boundary.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
import torch


def _lin(x, sd, name, bias=True):
    w = sd[name + ".weight"]
    y = x @ w.t()
    if bias:
        y = y + sd[name + ".bias"]
    return y


def _mlp2(x, sd, name):
    """Seq(Lin, ReLU, Lin) -- model.py:59-60."""
    return _lin(torch.relu(_lin(x, sd, name + ".0")), sd, name + ".2")


def _layer_norm(x, sd, name, eps=1e-6):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * sd[name + ".weight"] + sd[name + ".bias"]


def _attention(map_code, obs_code, sd, name, embed):
    """model.py:164-181: one shared K/Q/V; softmax over [self, obstacles] / sqrt(e)."""
    map_value = _lin(map_code, sd, name + ".value", bias=False)
    obs_value = _lin(obs_code, sd, name + ".value", bias=False)
    map_query = _lin(map_code, sd, name + ".query", bias=False)
    map_key = _lin(map_code, sd, name + ".key", bias=False)
    obs_key = _lin(obs_code, sd, name + ".key", bias=False)
    obs_att = map_query @ obs_key.t()                                  # [M, O]
    self_att = (map_query * map_key).sum(dim=-1, keepdim=True)        # [M, 1]
    att = torch.cat((self_att, obs_att), dim=-1) / (embed ** 0.5)
    att = att.softmax(dim=-1)
    new = att[:, :1] * map_value + att[:, 1:] @ obs_value              # never materialise [M,1+O,e]
    return _layer_norm(new + map_code, sd, name + ".layer_norm")


def _feed(x, sd, name):
    """model.py:193-201."""
    y = _lin(torch.relu(_lin(x, sd, name + ".w_1")), sd, name + ".w_2") + x
    return _layer_norm(y, sd, name + ".layer_norm")


def _block(map_code, obs_code, sd, name, embed):
    """model.py:212-218."""
    map_code = _attention(map_code, obs_code, sd, name + ".attention", embed)
    map_code = _feed(map_code, sd, name + ".map_feed")
    obs_code = _feed(obs_code, sd, name + ".obs_feed")
    return map_code, obs_code


def goal_index_of(v, goal):
    """knn(v, goal, k=1)[1] -- model.py:132; canonical fp32 left-to-right distance, first min."""
    v32 = v.to(torch.float32)
    g32 = goal.to(torch.float32).view(1, -1)
    d = torch.zeros(len(v32), dtype=torch.float32)
    for c in range(v32.shape[1]):
        diff = g32[:, c] - v32[:, c]
        d = d + diff * diff
    return int(torch.argmin(d))  # torch.argmin returns the first minimal index


@torch.no_grad()
def explorer_forward(sd, v, edge_index, goal, obstacles, loop=5, use_obstacles=True,
                     dtype=torch.float32, dense=True, return_intermediates=False):
    """Returns the dense ``[N,N]`` policy (``out[dst,src] = logit``) or the ``[E]`` logits.

    sd: reference state_dict; v [N,c]; edge_index [2,E] int64 (row0 = src j, row1 = dst i);
    goal [c]; obstacles [O,...] viewed as [-1, obs_size].
    """
    sd = {k: t.to(dtype) for k, t in sd.items() if t.is_floating_point()}
    embed = sd["goal_encoder"].numel()
    c = v.shape[1]
    obs_size = sd["obs_node_code.0.weight"].shape[1]
    v = v.to(dtype)
    goal = goal.to(dtype).view(-1, c)
    src, dst = edge_index[0].long(), edge_index[1].long()
    n = len(v)

    node_code = _mlp2(torch.cat((v, goal.repeat(n, 1), (v - goal) ** 2, v - goal), dim=-1), sd, "node_code")
    vv = torch.cat((v[src], v[dst]), dim=-1)
    edge_code = _mlp2(vv, sd, "edge_code")
    node_free = _mlp2(v, sd, "node_free_code")
    edge_free = _mlp2(vv, sd, "edge_free_code")

    if use_obstacles:
        obs = obstacles.to(dtype).reshape(-1, obs_size)
        obs_n = _mlp2(obs, sd, "obs_node_code")
        obs_e = _mlp2(obs, sd, "obs_edge_code")
        for i in range(3):
            node_free, obs_n = _block(node_free, obs_n, sd, "node_attentions.%d" % i, embed)
            edge_free, obs_e = _block(edge_free, obs_e, sd, "edge_attentions.%d" % i, embed)

    gi = goal_index_of(v, goal)
    h0 = torch.zeros(n, embed, dtype=dtype)
    h0[gi] = h0[gi] + sd["goal_encoder"]
    h = h0
    edge_attr = torch.cat((edge_free, edge_code), dim=-1)
    decode = None
    for _ in range(loop):
        x = _lin(torch.cat((node_code, node_free, h0, h), dim=-1), sd, "encoder")
        x_j, x_i = x[src], x[dst]
        msg = _mlp2(torch.cat((x_j - x_i, x_j, x_i, edge_attr), dim=-1), sd, "process.lin_0")
        agg = torch.full((n, embed), float("-inf"), dtype=dtype)
        agg = agg.scatter_reduce(0, dst.unsqueeze(-1).expand_as(msg), msg, "amax", include_self=True)
        agg = torch.where(torch.isinf(agg) & (agg < 0), torch.zeros_like(agg), agg)
        h = _lin(torch.cat((x, agg), dim=-1), sd, "process.lin_1")
        decode = _lin(torch.cat((node_code, h), dim=-1), sd, "decoder")

    z = torch.cat((decode[src], decode[src] - decode[dst], edge_free), dim=-1)
    z = torch.relu(_lin(z, sd, "policy.0"))
    z = torch.relu(_lin(z, sd, "policy.2"))
    logits = _lin(z, sd, "policy.4", bias=False).squeeze(-1)

    if return_intermediates:
        return dict(node_code=node_code, node_free=node_free, edge_code=edge_code, edge_free=edge_free,
                    h=h, decode=decode, logits=logits, goal_index=gi)
    if not dense:
        return logits
    out = torch.zeros(n, n, dtype=dtype)
    out[dst, src] = logits
    return out


def flops_per_graph(n, e_cnt, o, c, e, s, loop=5):
    """Algorithmic FLOPs (2*MAC) of the REFERENCE op list, SURVEY.md section 8(d)."""
    enc = 2 * n * (4 * c * e + e * e) + 2 * n * (c * e + e * e) + 2 * 2 * e_cnt * (2 * c * e + e * e) \
        + 2 * 2 * o * (s * e + e * e)

    def block(m):
        return 2 * (3 * m * e * e + 2 * o * e * e + 2 * m * (o + 1) * e + 2 * m * e * e + 2 * o * e * e)
    att = 3 * (block(n) + block(e_cnt))
    lp = loop * 2 * (n * 4 * e * e + e_cnt * (5 * e * e + e * e) + n * 2 * e * e + n * 2 * e * e)
    pol = 2 * e_cnt * (3 * e * e + e * e + e)
    return enc + att + lp + pol
