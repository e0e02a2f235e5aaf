"""CPU: the C-ABI shared library loads and exports every symbol include/gnnmp.h declares; the ctypes
binding declares exactly those symbols; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gnnmp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__
    __graft_entry__.build()
    from gnn_motion_planning_b200 import _lib
    return _lib


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libgnnmp.so does not export %s" % n


def test_binding_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == header_functions()
    built_lib.load()


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    lib = built_lib.load()
    assert lib.gmp_device_ok(0) == 0
    assert not lib.gmp_create(0)
    assert b"no CPU fallback" in lib.gmp_last_error()
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    m = EncoderProcessDecoder(2, 2, 32, 2)
    with pytest.raises(built_lib.GnnmpError):
        m.to("cpu")


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (only tests, smoke() and bench's baseline legs may)."""
    pkg = os.path.join(ROOT, "gnn_motion_planning_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
                assert "liboracle" not in txt


def test_explorer_state_dict_surface():
    """load_state_dict takes the reference's 200-tensor dict (dead tensors included) and rejects bad shapes."""
    import torch
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "weights_maze.pt"), map_location="cpu")
    assert len(sd) == 200
    m = EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2)
    m.load_state_dict(sd)
    out = m.state_dict()
    assert set(out) == set(sd)
    assert torch.equal(out["process.lin_0.0.weight"], sd["process.lin_0.0.weight"])
    bad = dict(sd)
    bad["encoder.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)
    del bad["encoder.weight"]
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)
    with pytest.raises(ValueError):
        EncoderProcessDecoder(2, 5, 48, 2)
