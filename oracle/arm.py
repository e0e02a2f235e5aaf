"""ORACLE (test infrastructure): ctypes front-end of ``oracle/arm.c`` -- the CPU restatement of the arm
state / edge collision check (reference ``environment/kuka_env.py:350-411``, ``kuka_2arm_env.py:352-402``) on
the sphere model this repository specifies (PyBullet parity is unpinned, see arm.c)."""
import ctypes

import numpy as np

from .maze import lib

KUKA7, KUKA14, KUKA13, UR5, SNAKE7 = 0, 1, 2, 3, 4


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def dof(model):
    return lib().oracle_arm_dof(model)


def limits(model):
    n = dof(model)
    lo, hi = np.zeros(n), np.zeros(n)
    lib().oracle_arm_limits(model, _ptr(lo), _ptr(hi))
    return lo, hi


def pack_boxes(problems_obstacles):
    """list (per problem) of [(halfExtents[3], basePosition[3]), ...] -> (boxes [O_total,6] f64, box_ptr [P+1] i32)."""
    rows, ptr = [], [0]
    for obs in problems_obstacles:
        for h, p in obs:   # ur5s_6_3000.pkl has ragged entries such as [0.01, 0.01, array([0.84])]
            rows.append(np.array([float(np.ravel(x)[0]) for x in list(h) + list(p)], np.float64))
        ptr.append(len(rows))
    return np.array(rows, np.float64).reshape(-1, 6), np.array(ptr, np.int32)


def state_fp(model, states, boxes, box_ptr, problem=None):
    states = np.ascontiguousarray(states)
    suf = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[states.dtype]
    n = states.reshape(-1, dof(model)).shape[0]
    free, counted = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    problem = None if problem is None else np.ascontiguousarray(problem, np.int32)
    getattr(lib(), "oracle_arm_state_fp_" + suf)(model, _ptr(states), _ptr(np.ascontiguousarray(boxes, np.float64)),
                                                  _ptr(np.ascontiguousarray(box_ptr, np.int32)), _ptr(problem),
                                                  ctypes.c_int64(n), _ptr(free), _ptr(counted))
    return free, counted


def edge_fp(model, a, b, boxes, box_ptr, problem=None, rrt_eps=0.5):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    suf = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[a.dtype]
    n = a.reshape(-1, dof(model)).shape[0]
    free, cnt = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    problem = None if problem is None else np.ascontiguousarray(problem, np.int32)
    getattr(lib(), "oracle_arm_edge_fp_" + suf)(model, _ptr(a), _ptr(b), _ptr(np.ascontiguousarray(boxes, np.float64)),
                                                 _ptr(np.ascontiguousarray(box_ptr, np.int32)), _ptr(problem),
                                                 ctypes.c_int64(n), ctypes.c_double(rrt_eps), _ptr(free), _ptr(cnt))
    return free, cnt


def snake_boxes(maps):
    """SnakeEnv.create_maze (snake_env.py:63-71): a box of half extents (0.7, 0.7, 1) at (1.4 i - 10.5, 1.4 j - 10.5, 0) for
    every occupied cell map[i, j], in the reference's loop order (j outer, i inner)."""
    obs = []
    for m in np.asarray(maps):
        obs.append([((0.7, 0.7, 1.0), (1.4 * i - 10.5, 1.4 * j - 10.5, 0.0)) for j in range(m.shape[0]) for i in range(m.shape[1]) if m[i, j]])
    return pack_boxes(obs)
