"""k-NN random-geometric-graph construction on the GPU (C ABI: ``gmp_knn_graph``).

Replaces the graph build inside reference ``eval_gnn.create_data`` (eval_gnn.py:159-164) for one
graph or a packed batch of graphs.
"""
import numpy as np
import torch

from . import _lib


def k1_of(k, n_free):
    """eval_gnn.py:159 -- evaluated in float64 exactly as the reference does."""
    return int(np.ceil(k * np.log(n_free) / np.log(100)))


_ws_cache = {}


def _workspace(device, nbytes):
    ws = _ws_cache.get(device)
    if ws is None or ws.numel() < nbytes:
        _ws_cache[device] = None
        ws = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)
        _ws_cache[device] = ws
    return ws


@torch.no_grad()
def knn_graph_batch(v, node_ptr, n_free, k1, edge_index_out=None, edge_ptr_host=None):
    """Symmetrised, coalesced k-NN graphs of a packed batch.

    v [N_total,c] f32 cuda; node_ptr [B+1], n_free [B], k1 [B] host ints.
    Returns (edge_index [2, capacity] i64 cuda -- valid columns are [0, edge_ptr[-1]) --, edge_ptr [B+1] host int32).
    edge_ptr_host: optional PINNED int32 [B+1] tensor; edge_ptr is then written into it by a kernel (``gmp_post_to_host``)
    instead of a device->host copy, so that this -- the only host synchronisation of a batch -- does not queue behind a large
    read-back in flight on the copy engine (pipelined callers: ``batch.HotPath``).
    """
    _lib.require_cuda(v, "v")
    lib = _lib.load()
    dev = v.device
    _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
    if v.dtype != torch.float32 or not v.is_contiguous():
        raise ValueError("v must be contiguous float32")
    node_ptr = np.ascontiguousarray(node_ptr, dtype=np.int32)
    n_free = np.ascontiguousarray(n_free, dtype=np.int32)
    k1 = np.ascontiguousarray(k1, dtype=np.int32)
    B = len(node_ptr) - 1
    n = np.diff(node_ptr).astype(np.int64)
    cap = int(sum(lib.gmp_knn_graph_max_edges(int(a), int(b)) for a, b in zip(n, k1))) if B else 0
    cap = max(cap, 1)
    if edge_index_out is None:
        edge_index_out = torch.empty((2, cap), dtype=torch.int64, device=dev)
    elif edge_index_out.shape[1] < cap or edge_index_out.dtype != torch.int64 or not edge_index_out.is_contiguous():
        raise ValueError("edge_index_out must be contiguous int64 [2, >=%d]" % cap)
    edge_ptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
    nbytes = lib.gmp_knn_graph_workspace_bytes(B, int(node_ptr[-1]), int(n.max()) if B else 0, int(k1.max()) if B else 0)
    ws = _workspace(dev, nbytes)
    _lib.check(lib.gmp_knn_graph(None, B, _lib.ptr(v), v.shape[1], _lib.ptr(node_ptr), _lib.ptr(n_free), _lib.ptr(k1),
                                 _lib.ptr(edge_index_out), edge_index_out.shape[1], _lib.ptr(edge_ptr), _lib.ptr(ws),
                                 ws.numel(), _lib.stream_ptr(dev)))
    if edge_ptr_host is None:
        return edge_index_out, edge_ptr.cpu().numpy()
    if (not edge_ptr_host.is_pinned() or edge_ptr_host.dtype != torch.int32 or edge_ptr_host.numel() != B + 1
            or not edge_ptr_host.is_contiguous()):
        raise ValueError("edge_ptr_host must be a pinned contiguous int32 [B+1] tensor")
    _lib.check(lib.gmp_post_to_host(_lib.ptr(edge_ptr), 4 * (B + 1), edge_ptr_host.data_ptr(), _lib.stream_ptr(dev)))
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(dev))
    done.synchronize()
    return edge_index_out, edge_ptr_host.numpy().copy()


@torch.no_grad()
def knn_graph_edges(v, n_free, k1):
    """Single graph: returns the [2,E] int64 cuda edge_index of create_data."""
    ei, ep = knn_graph_batch(v, np.array([0, v.shape[0]]), np.array([n_free]), np.array([k1]))
    return ei[:, :int(ep[1])]
