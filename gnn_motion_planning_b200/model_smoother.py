"""Host-side mirror of the reference ``model_smoother.py``: the GNN path smoother.

``ModelSmoother`` keeps the reference constructor and ``forward(**kwargs)`` surface (reference
``model_smoother.py:46-142``; called from ``smoother.py:243``) with the weights held by
``libgnnmp.so`` and the forward running in ``csrc/smoother.cu``.  ``forward_batch`` smooths many
problems' paths in one call.  Inference only (BatchNorm in eval mode), no CPU fallback.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

SUPPORTED_CONFIG_SIZES = (2, 3, 6, 7, 13, 14)


def _live_smoother_shapes(c, e):
    sh = OrderedDict()
    sh["node_code.0.weight"] = (e, c + 3)
    sh["node_code.0.bias"] = (e,)
    for k in ("weight", "bias", "running_mean", "running_var"):
        sh["node_code.1." + k] = (e,)
    sh["node_code.3.weight"] = (e, e)
    sh["node_code.3.bias"] = (e,)
    sh["process.lin_0.0.weight"] = (e, 3 * e)
    sh["process.lin_0.0.bias"] = (e,)
    for n in ("process.lin_0.2", "process.lin_1.0", "process.lin_1.2"):
        sh[n + ".weight"] = (e, e)
        sh[n + ".bias"] = (e,)
    sh["smooth_node.weight"] = (c, e)
    sh["smooth_node.bias"] = (c,)
    return sh


class ModelSmoother:
    """Drop-in for reference ``model_smoother.ModelSmoother`` (model_smoother.py:46)."""

    def __init__(self, workspace_size, config_size, obs_size, embed_size, scale=1.):
        if embed_size != 128 or config_size not in SUPPORTED_CONFIG_SIZES:
            raise ValueError("no sm_100a smoother kernel for config_size=%r embed_size=%r (supported: embed 128, config %r)"
                             % (config_size, embed_size, SUPPORTED_CONFIG_SIZES))
        self.workspace = workspace_size
        self.config_size = config_size
        self.obs_size = obs_size
        self.latent_dim = workspace_size
        self.scale = scale
        self.embed_size = embed_size
        self.training = False
        self._device = None
        self._handle = None
        self._uploaded = False
        self._ws = None
        self._state = OrderedDict()
        self._extra_state = OrderedDict()
        self.reset_parameters()

    def reset_parameters(self, seed=None):
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        shapes = _live_smoother_shapes(self.config_size, self.embed_size)
        for name, shape in shapes.items():
            if name.startswith("node_code.1."):
                t = {"weight": torch.ones, "bias": torch.zeros, "running_mean": torch.zeros, "running_var": torch.ones}[
                    name.rsplit(".", 1)[1]](shape)
            else:
                fan_in = shape[1] if len(shape) == 2 else shapes[name.replace(".bias", ".weight")][1]
                t = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(fan_in)
            self._state[name] = t.to(torch.float32)
        self._uploaded = False

    def state_dict(self):
        sd = OrderedDict(self._state)
        sd.update(self._extra_state)
        if "node_code.1.weight" in sd:     # bn2 is the same module registered twice (model_smoother.py:63,65)
            for k in ("weight", "bias", "running_mean", "running_var"):
                sd.setdefault("bn2." + k, sd["node_code.1." + k])
        return sd

    def load_state_dict(self, state_dict, strict=True):
        live = _live_smoother_shapes(self.config_size, self.embed_size)
        missing = [k for k in live if k not in state_dict]
        if missing and strict:
            raise RuntimeError("Error(s) in loading state_dict for ModelSmoother: missing keys %r" % missing)
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
            raise _lib.GnnmpError("ModelSmoother runs on CUDA (sm_100a) only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._device is None or self._device.index != idx:
            self._device = torch.device("cuda", idx)
            self._uploaded = False
            self._handle = None
            self._ws = None
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise _lib.GnnmpError("the B200 smoother is inference-only (BatchNorm runs in eval mode)")
        return self

    def parameters(self):
        return iter(self._state.values())

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

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
            _lib.check(lib.gmp_smoother_init(self._handle, self.config_size, self.embed_size))
            for name, t in self._state.items():
                a = np.ascontiguousarray(t.numpy(), dtype=np.float32)
                _lib.check(lib.gmp_smoother_set_tensor(self._handle, name.encode(), a.ctypes.data, a.size))
            _lib.check(lib.gmp_smoother_finalize(self._handle))
            self._uploaded = True

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().gmp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    @torch.no_grad()
    def forward_batch(self, path, samples, edge_index, path_ptr, sample_ptr, n_free, edge_ptr, loop=1):
        """path [P_total,c], samples [S_total,c] (= cat(free, collided) per problem), edge_index [2,E_total] i64 local ids,
        host int arrays path_ptr/sample_ptr/edge_ptr [B+1], n_free [B].  Returns the new paths [P_total,c]."""
        self._ensure_uploaded()
        lib = _lib.load()
        _lib.require_cuda(path, "path")
        dev = self._device
        path = path.to(dev, torch.float32).contiguous()
        samples = samples.to(dev, torch.float32).contiguous()
        edge_index = edge_index.to(dev, torch.int64)
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must be [2, E]")
        if edge_index.stride(1) != 1:
            edge_index = edge_index.contiguous()
        path_ptr = np.ascontiguousarray(path_ptr, dtype=np.int32)
        sample_ptr = np.ascontiguousarray(sample_ptr, dtype=np.int32)
        edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int32)
        n_free = np.ascontiguousarray(n_free, dtype=np.int32)
        B = len(path_ptr) - 1
        if path.shape != (int(path_ptr[-1]), self.config_size) or samples.shape[0] != int(sample_ptr[-1]):
            raise ValueError("path / samples shapes do not match the offset arrays")
        nbytes = lib.gmp_smoother_workspace_bytes(self._handle, B, int(path_ptr[-1]), int(sample_ptr[-1]), int(edge_ptr[-1]))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=dev)
        out = torch.empty_like(path)
        _lib.check(lib.gmp_smoother_forward(
            self._handle, B, _lib.ptr(path), _lib.ptr(samples), _lib.ptr(edge_index), edge_index.stride(0),
            _lib.ptr(path_ptr), _lib.ptr(sample_ptr), _lib.ptr(n_free), _lib.ptr(edge_ptr), float(self.scale), int(loop),
            _lib.ptr(out), _lib.ptr(self._ws), self._ws.numel(), _lib.stream_ptr(dev)))
        return out

    @torch.no_grad()
    def forward(self, path, free, collided, obstacles=None, edge_index=None, loop=10, **kwargs):
        """Same arguments / return as the reference (model_smoother.py:104): new path ``[P,c]`` fp32 on the device."""
        if self._device is None:
            self.to(path.device if torch.is_tensor(path) and path.is_cuda else "cuda")
        dev = self._device
        path = torch.as_tensor(path).to(dev, torch.float32)
        free = torch.as_tensor(free).to(dev, torch.float32).reshape(-1, self.config_size)
        collided = torch.as_tensor(collided).to(dev, torch.float32).reshape(-1, self.config_size)
        samples = torch.cat((free, collided), dim=0)
        edge_index = torch.as_tensor(edge_index).to(dev, torch.int64)
        return self.forward_batch(path, samples, edge_index, [0, path.shape[0]], [0, samples.shape[0]], [free.shape[0]],
                                  [0, edge_index.shape[1]], loop=loop)
