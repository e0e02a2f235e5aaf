"""Host-side mirror of the reference ``model.py``: the GNN path explorer.

``EncoderProcessDecoder`` keeps the reference constructor, ``.to() / .eval() / .load_state_dict()``
and ``forward(**kwargs)`` call surface (reference ``model.py:48-150``; called from
``eval_gnn.py:194``), but holds no ``torch.nn`` modules: the weights go to ``libgnnmp.so`` once, and
``forward`` is one C-ABI call that runs the hand-written sm_100a kernels of ``csrc/explorer.cu``.
PyTorch tensors are only the container for device memory.  There is no CPU fallback.

Beyond the drop-in single-graph call the class exposes what the reference cannot do:
``forward_batch`` scores a packed batch of independent planning problems in one call.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

# (config_size, embed_size, obs_size) combinations of reference str2name.py:12-66
SUPPORTED_DIMS = {(2, 32, 2), (3, 32, 2), (7, 64, 6), (6, 32, 6), (7, 32, 2), (13, 32, 6), (14, 32, 6)}


def _live_explorer_shapes(c, e, s):
    """name -> shape of every tensor EncoderProcessDecoder.forward touches."""
    sh = OrderedDict()
    sh["goal_encoder"] = (e,)

    def lin(name, i, o=e, bias=True):
        sh[name + ".weight"] = (o, i)
        if bias:
            sh[name + ".bias"] = (o,)

    def seq(name, i):
        lin(name + ".0", i)
        lin(name + ".2", e)

    seq("node_code", 4 * c)
    seq("edge_code", 2 * c)
    seq("obs_node_code", s)
    seq("obs_edge_code", s)
    seq("node_free_code", c)
    seq("edge_free_code", 2 * c)
    for st in ("node_attentions", "edge_attentions"):
        for i in range(3):
            p = "%s.%d." % (st, i)
            for q in ("key", "query", "value"):
                lin(p + "attention." + q, e, bias=False)
            sh[p + "attention.layer_norm.weight"] = (e,)
            sh[p + "attention.layer_norm.bias"] = (e,)
            for f in ("map_feed", "obs_feed"):
                lin(p + f + ".w_1", e)
                lin(p + f + ".w_2", e)
                sh[p + f + ".layer_norm.weight"] = (e,)
                sh[p + f + ".layer_norm.bias"] = (e,)
    lin("encoder", 4 * e)
    lin("process.lin_0.0", 5 * e)
    lin("process.lin_0.2", e)
    lin("process.lin_1", 2 * e)
    lin("decoder", 2 * e)
    lin("policy.0", 3 * e)
    lin("policy.2", e)
    lin("policy.4", e, o=1, bias=False)
    return sh


class EncoderProcessDecoder:
    """Drop-in for reference ``model.EncoderProcessDecoder`` (model.py:48).  Inference only."""

    def __init__(self, workspace_size, config_size, embed_size, obs_size, use_obstacles=True):
        if (config_size, embed_size, obs_size) not in SUPPORTED_DIMS:
            raise ValueError("no sm_100a kernel instantiated for (config_size, embed_size, obs_size)=%r; supported: %r"
                             % ((config_size, embed_size, obs_size), sorted(SUPPORTED_DIMS)))
        self.workspace = workspace_size
        self.config_size = config_size
        self.obs_size = obs_size
        self.embed_size = embed_size
        self.use_obstacles = use_obstacles      # poked from outside by eval_gnn.py:88
        self.training = False
        self._device = None
        self._handle = None
        self._uploaded = False
        self._ws = None
        self._state = OrderedDict()
        self._extra_state = OrderedDict()       # dead tensors of the reference state_dict, kept for state_dict()
        self.reset_parameters()

    # ------------------------------------------------------------------ nn.Module-like surface
    def reset_parameters(self, seed=None):
        """torch.nn.Linear-style init (U(-1/sqrt(in), 1/sqrt(in))), LayerNorm (1, 0), goal_encoder U(0,1) (model.py:77)."""
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        for name, shape in _live_explorer_shapes(self.config_size, self.embed_size, self.obs_size).items():
            if name == "goal_encoder":
                t = torch.rand(shape, generator=gen)
            elif "layer_norm.weight" in name:
                t = torch.ones(shape)
            elif "layer_norm.bias" in name:
                t = torch.zeros(shape)
            else:
                fan_in = shape[1] if len(shape) == 2 else _live_explorer_shapes(
                    self.config_size, self.embed_size, self.obs_size)[name.replace(".bias", ".weight")][1]
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            self._state[name] = t.to(torch.float32)
        self._uploaded = False

    def state_dict(self):
        sd = OrderedDict(self._state)
        sd.update(self._extra_state)
        return sd

    def load_state_dict(self, state_dict, strict=True):
        """Accepts the reference's 200-tensor dict (dead tensors included)."""
        live = _live_explorer_shapes(self.config_size, self.embed_size, self.obs_size)
        missing = [k for k in live if k not in state_dict]
        if missing and strict:
            raise RuntimeError("Error(s) in loading state_dict for EncoderProcessDecoder: missing keys %r" % missing)
        for k, shape in live.items():
            if k in state_dict:
                t = torch.as_tensor(state_dict[k]).detach().to("cpu", torch.float32)
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError("size mismatch for %s: copying a param with shape %r, the model expects %r"
                                       % (k, tuple(t.shape), tuple(shape)))
                self._state[k] = t.contiguous().clone()
        self._extra_state = OrderedDict((k, v) for k, v in state_dict.items() if k not in live)
        self._uploaded = False
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.GnnmpError("EncoderProcessDecoder runs on CUDA (sm_100a) only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._device is None or self._device.index != idx:
            self._device = torch.device("cuda", idx)
            self._uploaded = False
            self._handle = None
            self._ws = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise _lib.GnnmpError("the B200 explorer is inference-only (training is out of scope, SURVEY.md section 2)")
        return self

    def parameters(self):
        return iter(self._state.values())

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    # ------------------------------------------------------------------ device plumbing
    def _ensure_uploaded(self):
        if self._device is None:
            self.to("cuda")
        lib = _lib.load()
        if self._handle is None:
            h = lib.gmp_create(self._device.index)
            if not h:
                raise _lib.GnnmpError(lib.gmp_last_error().decode())
            self._handle = h
        if not self._uploaded:
            _lib.check(lib.gmp_explorer_init(self._handle, self.config_size, self.embed_size, self.obs_size))
            for name, t in self._state.items():
                a = np.ascontiguousarray(t.numpy(), dtype=np.float32)
                _lib.check(lib.gmp_explorer_set_tensor(self._handle, name.encode(), a.ctypes.data, a.size))
            _lib.check(lib.gmp_explorer_finalize(self._handle))
            self._uploaded = True

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().gmp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self._device)
        return self._ws

    PHASES = ("csr_build", "goal_index", "obstacle_stream", "node_pre", "edge_feature", "node_loop", "edge_msg", "policy")

    def set_edge_feature_mode(self, mode):
        """Arithmetic of the edge-feature stage: "auto" (tcgen05 3xTF32 tensor cores), "simt" (fp32 FMA), "tc" (eight epilogue
        warps per 128-edge tile, columns split between warp pairs, ONE lockstep MMA issuer), "tc4" (embed 32 only: the round-1
        organisation, four warps per tile) or "tcrd" (= what auto picks for narrow inputs when every graph has 1..128 obstacles:
        eight warps per tile and one issuer warp PER TILE).  All meet the 1e-4 logit tolerance and "tc" / "tcrd" are bit-identical;
        the switch exists for A/B parity tests and profiling."""
        self._ensure_uploaded()
        code = {"auto": -1, "simt": 0, "tc": 1, "tc4": 2, "tcrd": 3}[mode]
        _lib.check(_lib.load().gmp_explorer_set_edge_feature_mode(self._handle, code))

    def last_bad_edges(self):
        """Number of edge_index entries outside [0, N_g) in the last ``forward_batch`` / ``forward_sparse`` (they are clamped on
        the device so that nothing is written out of bounds; the drop-in ``forward`` raises IndexError up front).  Synchronises."""
        rc = _lib.load().gmp_explorer_bad_edges(self._handle, _lib.stream_ptr(self._device))
        if rc < 0:
            _lib.check(rc)
        return rc

    def set_timing(self, enable=True):
        """Record CUDA events around every phase of subsequent forwards (see ``last_timings``)."""
        self._ensure_uploaded()
        _lib.check(_lib.load().gmp_set_timing(self._handle, int(bool(enable))))

    def last_timings(self):
        """dict phase -> milliseconds of the last forward (synchronises on its events)."""
        ms = np.zeros(len(self.PHASES), np.float32)
        rc = _lib.load().gmp_get_timings(self._handle, ms.ctypes.data, len(ms))
        if rc < 0:
            _lib.check(rc)
        return dict(zip(self.PHASES, ms.tolist()))

    # ------------------------------------------------------------------ batched entry point (new capability)
    @torch.no_grad()
    def forward_batch(self, v, edge_index, goal, obstacles, node_ptr, edge_ptr, obs_ptr, loop=5, dense=False,
                      out=None, dense_out=None):
        """Score every edge of a packed batch of graphs.

        v [N_total,c] f32 cuda; edge_index [2,E_total] i64 cuda (local ids); goal [B,c]; obstacles [O_total,s];
        node_ptr/edge_ptr/obs_ptr: host int32 [B+1].  Returns logits [E_total] (and the packed dense
        matrices [sum N_g^2] when ``dense``).
        """
        self._ensure_uploaded()
        lib = _lib.load()
        for name, t in (("v", v), ("edge_index", edge_index), ("goal", goal)):
            _lib.require_cuda(t, name)
        node_ptr = np.ascontiguousarray(node_ptr, dtype=np.int32)
        edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int32)
        B = len(node_ptr) - 1
        nt, et = int(node_ptr[-1]), int(edge_ptr[-1])
        if v.dtype != torch.float32 or not v.is_contiguous() or v.shape != (nt, self.config_size):
            raise ValueError("v must be contiguous float32 [N_total=%d, c=%d], got %r %r" % (nt, self.config_size, v.dtype, tuple(v.shape)))
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2 or edge_index.shape[1] < et:
            raise ValueError("edge_index must be int64 [2, >=E_total=%d]" % et)
        if edge_index.stride(1) != 1:
            edge_index = edge_index.contiguous()
        goal = goal.reshape(B, self.config_size).to(torch.float32).contiguous()
        use_obs = bool(self.use_obstacles)
        if use_obs:
            obs_ptr = np.ascontiguousarray(obs_ptr, dtype=np.int32)
            _lib.require_cuda(obstacles, "obstacles")
            obstacles = obstacles.reshape(-1, self.obs_size).to(torch.float32).contiguous()
            if obstacles.shape[0] != int(obs_ptr[-1]):
                raise ValueError("obstacles rows (%d) != obs_ptr[-1] (%d)" % (obstacles.shape[0], int(obs_ptr[-1])))
            n_obs = int(obs_ptr[-1])
        else:
            obs_ptr, n_obs = None, 0
        nbytes = lib.gmp_explorer_workspace_bytes(self._handle, B, nt, et, n_obs)
        ws = self._workspace(nbytes)
        logits = out if out is not None else torch.empty(et, dtype=torch.float32, device=self._device)
        dn = None
        if dense:
            n_dense = int((np.diff(node_ptr).astype(np.int64) ** 2).sum())
            dn = dense_out if dense_out is not None else torch.empty(n_dense, dtype=torch.float32, device=self._device)
        _lib.check(lib.gmp_explorer_forward(
            self._handle, B, _lib.ptr(v), _lib.ptr(edge_index), edge_index.stride(0), _lib.ptr(goal),
            _lib.ptr(obstacles) if use_obs else None, _lib.ptr(node_ptr), _lib.ptr(edge_ptr),
            _lib.ptr(obs_ptr) if use_obs else None, int(loop), int(use_obs), _lib.ptr(logits), _lib.ptr(dn),
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)))
        return (logits, dn) if dense else logits

    # ------------------------------------------------------------------ drop-in single-graph call (model.py:115)
    @torch.no_grad()
    def forward(self, goal, loop, v, obstacles, free=None, collided=None, edge_index=None, k=10, **kwargs):
        """Same arguments and return value as the reference: dense fp32 ``[N,N]`` with ``out[dst,src] = logit``."""
        if self._device is None:
            self.to(v.device if torch.is_tensor(v) and v.is_cuda else "cuda")
        dev = self._device
        v = torch.as_tensor(v).to(dev, torch.float32).contiguous()
        edge_index = torch.as_tensor(edge_index).to(dev, torch.int64).contiguous()
        goal = torch.as_tensor(goal).to(dev, torch.float32).reshape(1, -1)
        n, e = v.shape[0], edge_index.shape[1]
        if e and (int(edge_index.min()) < 0 or int(edge_index.max()) >= n):
            raise IndexError("edge_index out of range for %d nodes" % n)
        if self.use_obstacles:
            obstacles = torch.as_tensor(obstacles).to(dev, torch.float32).reshape(-1, self.obs_size).contiguous()
            obs_ptr = np.array([0, obstacles.shape[0]], np.int32)
        else:
            obs_ptr = None
        _, dense = self.forward_batch(v, edge_index, goal, obstacles, np.array([0, n], np.int32), np.array([0, e], np.int32),
                                      obs_ptr, loop=loop, dense=True)
        return dense.view(n, n)

    @torch.no_grad()
    def forward_sparse(self, goal, loop, v, obstacles, edge_index, **kwargs):
        """Logits ``[E]`` aligned with ``edge_index`` (the `policy` vector of model.py:145 before the dense scatter)."""
        if self._device is None:
            self.to(v.device if torch.is_tensor(v) and v.is_cuda else "cuda")
        dev = self._device
        v = torch.as_tensor(v).to(dev, torch.float32).contiguous()
        edge_index = torch.as_tensor(edge_index).to(dev, torch.int64).contiguous()
        goal = torch.as_tensor(goal).to(dev, torch.float32).reshape(1, -1)
        n, e = v.shape[0], edge_index.shape[1]
        if self.use_obstacles:
            obstacles = torch.as_tensor(obstacles).to(dev, torch.float32).reshape(-1, self.obs_size).contiguous()
            obs_ptr = np.array([0, obstacles.shape[0]], np.int32)
        else:
            obs_ptr = None
        return self.forward_batch(v, edge_index, goal, obstacles, np.array([0, n], np.int32), np.array([0, e], np.int32),
                                  obs_ptr, loop=loop, dense=False)
