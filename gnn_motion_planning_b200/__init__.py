"""gnn_motion_planning_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of
rainorangelemon/gnn-motion-planning, behind the reference's own Python call surface.

    model.EncoderProcessDecoder      <- reference model.py            (explorer forward)
    model_smoother.ModelSmoother     <- reference model_smoother.py   (smoother forward)
    eval_gnn.create_data / Data / obs_data / explore   <- reference eval_gnn.py
    environment.MazeEnv / KukaEnv / Kuka2Env            <- reference environment/*_env.py
    batch                             packed-batch entry points (new capability: many problems per call)

All compute goes through the C ABI of ``libgnnmp.so`` (``include/gnnmp.h``); PyTorch is only the
container for device memory, streams and ``torch.distributed``.  No CPU fallback exists.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
