"""ORACLE (test infrastructure): ctypes front-end of ``oracle/maze.c`` -- the plain-C restatement
of reference ``environment/maze_env.py:236-325`` (2-D maze state / edge collision check)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("maze.c", "arm.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _prep(states, maps, problem):
    dt = np.asarray(states).dtype
    if dt not in (np.float32, np.float64):
        raise TypeError("states must be float32 or float64")
    maps = np.ascontiguousarray(maps, dtype=np.uint8).reshape(-1, 15, 15)
    if problem is not None:
        problem = np.ascontiguousarray(problem, dtype=np.int32)
    return ("f32" if dt == np.float32 else "f64"), maps, problem


def state_fp(states, maps, problem=None):
    """-> (free uint8[n], counted uint8[n])."""
    suf, maps, problem = _prep(states, maps, problem)
    s = np.ascontiguousarray(states).reshape(-1, 2)
    n = len(s)
    free = np.zeros(n, np.uint8)
    counted = np.zeros(n, np.uint8)
    getattr(lib(), "oracle_maze_state_fp_" + suf)(_ptr(s), _ptr(maps), _ptr(problem), ctypes.c_int64(n),
                                                   _ptr(free), _ptr(counted))
    return free, counted


def edge_fp(a, b, maps, problem=None):
    """-> (free uint8[n], n_checks int32[n])."""
    suf, maps, problem = _prep(a, maps, problem)
    a = np.ascontiguousarray(a).reshape(-1, 2)
    b = np.ascontiguousarray(b, dtype=a.dtype).reshape(-1, 2)
    n = len(a)
    free = np.zeros(n, np.uint8)
    cnt = np.zeros(n, np.int32)
    getattr(lib(), "oracle_maze_edge_fp_" + suf)(_ptr(a), _ptr(b), _ptr(maps), _ptr(problem), ctypes.c_int64(n),
                                                  _ptr(free), _ptr(cnt))
    return free, cnt


def _prep3(states, maps, problem):
    dt = np.asarray(states).dtype
    if dt not in (np.float32, np.float64):
        raise TypeError("states must be float32 or float64")
    maps = np.ascontiguousarray(maps, dtype=np.uint8).reshape(-1, 15, 15)
    if problem is not None:
        problem = np.ascontiguousarray(problem, dtype=np.int32)
    return ("f32" if dt == np.float32 else "f64"), maps, problem


def stick_state_fp(states, maps, problem=None):
    """3-D stick maze, MazeEnv(dim=3)._state_fp (maze_env.py:279-291) -> (free uint8[n], n_checks int32[n], k int32[n])."""
    suf, maps, problem = _prep3(states, maps, problem)
    s = np.ascontiguousarray(states).reshape(-1, 3)
    n = len(s)
    free, cnt, k = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
    getattr(lib(), "oracle_maze3_state_fp_" + suf)(_ptr(s), _ptr(maps), _ptr(problem), ctypes.c_int64(n), _ptr(free), _ptr(cnt), _ptr(k))
    return free, cnt, k


def stick_edge_fp(a, b, maps, problem=None):
    """3-D stick maze, MazeEnv(dim=3)._edge_fp (maze_env.py:316-347) -> (free uint8[n], n_checks int32[n], k int32[n])."""
    suf, maps, problem = _prep3(a, maps, problem)
    a = np.ascontiguousarray(a).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=a.dtype).reshape(-1, 3)
    n = len(a)
    free, cnt, k = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
    getattr(lib(), "oracle_maze3_edge_fp_" + suf)(_ptr(a), _ptr(b), _ptr(maps), _ptr(problem), ctypes.c_int64(n), _ptr(free), _ptr(cnt),
                                                  _ptr(k))
    return free, cnt, k
