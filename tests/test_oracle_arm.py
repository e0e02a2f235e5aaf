"""CPU: the arm-collision oracle (oracle/arm.c) -- model sanity against the PyBullet-verified free states the
reference's problem files ship, and the specification-level pieces (sincos, limits, counters)."""
import math
import os

import numpy as np
import pytest

from oracle import arm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = [("kuka7", arm.KUKA7, 7), ("kuka14", arm.KUKA14, 14), ("kuka13", arm.KUKA13, 13), ("ur5", arm.UR5, 6)]


@pytest.fixture(scope="module")
def probs():
    return np.load(os.path.join(G, "arm_problems.npz"))


@pytest.mark.parametrize("tag,model,dof", MODELS)
def test_known_free_states(probs, tag, model, dof):
    """start / goal / solution waypoints of the dataset were collision-free in PyBullet; the inscribed-sphere model is a
    subset of the link hulls, so it must call (nearly) all of them free.  kuka14 uses model_0.urdf as a stand-in for
    pybullet_data's model.urdf, hence the looser band."""
    assert arm.dof(model) == dof
    free, counted = arm.state_fp(model, probs[tag + "_known_free"], probs[tag + "_boxes"], probs[tag + "_box_ptr"],
                                 probs[tag + "_known_free_problem"])
    assert counted.all()
    assert free.mean() >= {"kuka7": 0.999, "kuka14": 0.95, "kuka13": 0.97, "ur5": 0.999}[tag], free.mean()
    fe, ce = arm.edge_fp(model, probs[tag + "_path_a"], probs[tag + "_path_b"], probs[tag + "_boxes"], probs[tag + "_box_ptr"],
                         probs[tag + "_path_problem"], rrt_eps=0.1 if tag == "ur5" else 0.5)
    assert fe.mean() >= {"kuka7": 0.999, "kuka14": 0.95, "kuka13": 0.97, "ur5": 0.999}[tag], fe.mean()
    assert (ce[fe == 1] >= 2).all()


def test_model_blocks_obvious_collisions(probs):
    """A box swallowing the whole arm makes every in-limit state collide; out-of-limit states are not counted."""
    lo, hi = arm.limits(arm.KUKA7)
    rng = np.random.default_rng(0)
    q = rng.uniform(lo, hi, (200, 7))
    boxes = np.array([[2.0, 2.0, 2.0, 0.0, 0.0, 0.5]])
    free, counted = arm.state_fp(arm.KUKA7, q, boxes, np.array([0, 1], np.int32))
    assert not free.any() and counted.all()
    free, counted = arm.state_fp(arm.KUKA7, q, np.zeros((0, 6)), np.array([0, 0], np.int32))
    assert free.all()
    q[:, 3] = hi[3] + 1e-9
    free, counted = arm.state_fp(arm.KUKA7, q, np.zeros((0, 6)), np.array([0, 0], np.int32))
    assert not free.any() and not counted.any()


def test_edge_counts_follow_reference_loop(probs):
    """kuka_env.py:397-409: 2 endpoint checks + K = int(d / 0.5) interpolation states when everything is free."""
    lo, hi = arm.limits(arm.KUKA7)
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        a = rng.uniform(lo * 0.9, hi * 0.9, (300, 7)).astype(dt)
        b = rng.uniform(lo * 0.9, hi * 0.9, (300, 7)).astype(dt)
        free, cnt = arm.edge_fp(arm.KUKA7, a, b, np.zeros((0, 6)), np.array([0, 0], np.int32), rrt_eps=0.5)
        d = np.sqrt(np.sum(np.abs(b - a) ** 2, axis=-1))
        assert free.all()
        assert np.array_equal(cnt, 2 + (d / dt(0.5)).astype(int))


def test_two_arm_self_collision():
    """Kuka2Env: both arms bent towards each other collide with no obstacle present (kuka_2arm_env.py:363)."""
    q = np.zeros((2, 14))
    q[1, 1] = 1.4    # arm at x=-0.5 leans towards +x ...
    q[1, 8] = -1.4   # ... arm at x=+0.5 leans towards -x
    free, _ = arm.state_fp(arm.KUKA14, q, np.zeros((0, 6)), np.array([0, 0], np.int32))
    assert free[0] == 1 and free[1] == 0


def test_sincos_spec():
    import ctypes
    import subprocess
    import tempfile
    src = '#include "%s"\nvoid sc(double x, double* s, double* c) { gmp_sincos(x, s, c); }\n' % os.path.join(
        os.path.dirname(G), "..", "include", "gmp_arm_math.h")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", os.path.join(td, "t.so"), os.path.join(td, "t.c")])
        lib = ctypes.CDLL(os.path.join(td, "t.so"))
        s, c = ctypes.c_double(), ctypes.c_double()
        for x in np.linspace(-7, 7, 2001):
            lib.sc(ctypes.c_double(x), ctypes.byref(s), ctypes.byref(c))
            assert abs(s.value - math.sin(x)) < 4e-16 and abs(c.value - math.cos(x)) < 4e-16


def test_ur5_self_collision_and_plane():
    """UR5Env (ur5_env.py:107-111): self collision between links that are not directly connected, ground plane z = 0."""
    none = (np.zeros((0, 6)), np.array([0, 0], np.int32))
    up = np.array([[0.0, -np.pi / 2, 0.0, -np.pi / 2, 0.0, 0.0]])          # arm pointing straight up: free
    assert arm.state_fp(arm.UR5, up, *none)[0][0] == 1
    down = np.array([[0.0, np.pi / 2, 0.0, 0.0, 0.0, 0.0]])                # upper arm pointing into the ground
    assert arm.state_fp(arm.UR5, down, *none)[0][0] == 0
    folded = np.array([[0.0, -np.pi / 2, np.pi, 0.0, 0.0, 0.0]])           # forearm folded back through the upper arm / shoulder
    assert arm.state_fp(arm.UR5, folded, *none)[0][0] == 0


def test_snake_model(probs):
    """SnakeEnv: all dataset init / goal states (free under PyBullet) are free; the config[3] double use and the unused
    config[6] of snake_env.py:124-128 are reproduced; a folded snake collides with itself."""
    boxes, ptr = arm.snake_boxes(probs["snake7_maps"])
    n = len(probs["snake7_start"])
    S = np.concatenate([probs["snake7_start"], probs["snake7_goal"]])
    free, counted = arm.state_fp(arm.SNAKE7, S, boxes, ptr, np.concatenate([np.arange(n), np.arange(n)]))
    assert counted.all() and free.all()
    none = (np.zeros((0, 6)), np.array([0, 0], np.int32))
    q = np.array([[0.0, 0.0, 0.3, -0.2, 0.4, 0.1, 0.0]])
    q2 = q.copy(); q2[0, 6] = 2.5                        # config[6] is never read
    assert arm.state_fp(arm.SNAKE7, q, *none)[0][0] == 1 and arm.state_fp(arm.SNAKE7, q2, *none)[0][0] == 1
    fold = np.array([[0.0, 0.0, 3.0, 0.0, 0.0, 0.0, 0.0]])   # first joint folded back by ~172 degrees: link 1 lies on the shoulder link
    assert arm.state_fp(arm.SNAKE7, fold, *none)[0][0] == 0
    lo, hi = arm.limits(arm.SNAKE7)
    assert lo[0] == -9 and hi[1] == 9 and abs(hi[2] - np.pi) < 1e-15
