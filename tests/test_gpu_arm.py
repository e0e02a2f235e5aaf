"""GPU: arm collision kernels vs the C oracle -- bit-exact booleans and check counts (model-level parity; PyBullet
parity is unpinned)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = [("kuka7", 0, 7), ("kuka14", 1, 14), ("kuka13", 2, 13), ("ur5", 3, 6)]


@pytest.fixture(scope="module")
def probs():
    return np.load(os.path.join(G, "arm_problems.npz"))


@pytest.mark.parametrize("tag,model,dof", MODELS)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_state_and_edge_vs_oracle(cuda_device, probs, tag, model, dof, dt):
    from gnn_motion_planning_b200 import collision
    from oracle import arm as o_arm
    d, lo, hi = collision.arm_model_info(model)
    assert d == dof
    olo, ohi = o_arm.limits(model)
    assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    boxes, ptr = probs[tag + "_boxes"], probs[tag + "_box_ptr"]
    P = len(ptr) - 1
    rng = np.random.default_rng(3)
    n = 60000
    q = rng.uniform(lo * 1.01, hi * 1.01, (n, dof)).astype(dt)        # a few out of limits
    prob = rng.integers(0, P, n).astype(np.int32)
    bd, pd = torch.from_numpy(boxes).to(cuda_device), torch.from_numpy(ptr).to(cuda_device)
    free, counted = collision.arm_state_fp(model, torch.from_numpy(q).to(cuda_device), bd, pd, torch.from_numpy(prob).to(cuda_device),
                                           want_counted=True)
    of, oc = o_arm.state_fp(model, q, boxes, ptr, prob)
    assert np.array_equal(free.cpu().numpy(), of) and np.array_equal(counted.cpu().numpy(), oc)
    assert 0.1 < of.mean() < 0.95
    m = 20000
    a = rng.uniform(lo, hi, (m, dof)).astype(dt)
    eps = 0.1 if tag == "ur5" else 0.5
    b = np.clip(a + rng.normal(0, 0.6 * eps / 0.5, (m, dof)), lo * 1.005, hi * 1.005).astype(dt)
    free, checks = collision.arm_edge_fp(model, torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device), bd, pd,
                                         torch.from_numpy(prob[:m]).to(cuda_device), rrt_eps=eps, want_checks=True)
    of, oc = o_arm.edge_fp(model, a, b, boxes, ptr, prob[:m], rrt_eps=eps)
    assert np.array_equal(free.cpu().numpy(), of)
    assert np.array_equal(checks.cpu().numpy(), oc)
    assert 0.02 < of.mean() < 0.95


def test_edge_graph_matches_explicit(cuda_device, probs):
    from gnn_motion_planning_b200 import collision
    boxes, ptr = probs["kuka7_boxes"], probs["kuka7_box_ptr"]
    _, lo, hi = collision.arm_model_info(0)
    rng = np.random.default_rng(5)
    B, n = 4, 150
    v = rng.uniform(lo, hi, (B * n, 7)).astype(np.float32)
    v[[5, 170, 171, 449]] *= 3.0          # a few nodes outside the joint limits (the graph form caches per-node validity / freeness)
    es = [rng.integers(0, n, (2, 500 + 7 * g)) for g in range(B)]
    edge_ptr = np.cumsum([0] + [e.shape[1] for e in es]).astype(np.int32)
    node_ptr = (np.arange(B + 1) * n).astype(np.int32)
    ei = np.concatenate(es, 1).astype(np.int64)
    dev = cuda_device
    bd, pd = torch.from_numpy(boxes).to(dev), torch.from_numpy(ptr).to(dev)
    pg = torch.tensor([3, 0, 7, 7], dtype=torch.int32, device=dev)
    free, checks = collision.arm_edge_fp_graph(0, torch.from_numpy(v).to(dev), torch.from_numpy(ei).to(dev), torch.from_numpy(node_ptr).to(dev),
                                               torch.from_numpy(edge_ptr).to(dev), bd, pd, int(edge_ptr[-1]), rrt_eps=0.5,
                                               problem_of_graph=pg, want_checks=True)
    gid = np.repeat(np.arange(B), np.diff(edge_ptr))
    a, b = v[ei[0] + node_ptr[gid]], v[ei[1] + node_ptr[gid]]
    f2, c2 = collision.arm_edge_fp(0, torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), bd, pd,
                                   torch.from_numpy(pg.cpu().numpy()[gid]).to(dev), rrt_eps=0.5, want_checks=True)
    assert torch.equal(free, f2) and torch.equal(checks, c2)


def test_kuka_env_protocol(cuda_device, probs):
    """Drop-in env: scalar _state_fp/_edge_fp + collision_check_count agree with the oracle; sampler keeps the RNG stream."""
    from gnn_motion_planning_b200.environment import Kuka2Env, KukaEnv
    from oracle import arm as o_arm
    for tag, cls, model in (("kuka7", KukaEnv, 0), ("kuka14", Kuka2Env, 1)):
        boxes, ptr = probs[tag + "_boxes"], probs[tag + "_box_ptr"]
        problems = []
        for i in range(len(ptr) - 1):
            obs = [(boxes[j, :3], boxes[j, 3:]) for j in range(ptr[i], ptr[i + 1])]
            problems.append((obs, probs[tag + "_start"][i], probs[tag + "_goal"][i], []))
        env = cls(problems=problems)
        assert str(env) == tag and env.config_dim == {"kuka7": 7, "kuka14": 14}[tag] and env.RRT_EPS == 0.5
        env.init_new_problem(2)
        np.random.seed(4)
        free, coll = env.sample_n_points(30, need_negative=True)
        assert len(free) == 30 and env.collision_check_count == len(free) + len(coll)
        np.random.seed(4)     # reference loop: one draw per check, stop at the 30th free sample
        pr = np.array(env.pose_range)
        ref_free = []
        while len(ref_free) < 30:
            s = np.random.uniform(pr[:, 0], pr[:, 1], size=(1, env.config_dim)).reshape(-1)
            if o_arm.state_fp(model, s[None], boxes, ptr, np.array([2], np.int32))[0][0]:
                ref_free.append(s)
        assert np.array_equal(np.array(free), np.array(ref_free))
        c0 = env.collision_check_count
        a, b = free[0].astype(np.float32), free[1].astype(np.float32)
        got = env._edge_fp(a, b)
        of, oc = o_arm.edge_fp(model, a[None], b[None], boxes, ptr, np.array([2], np.int32), rrt_eps=0.5)
        assert got == bool(of[0]) and env.collision_check_count - c0 == oc[0]
        assert env._state_fp(np.asarray(env.init_state)) in (True, False)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_snake_vs_oracle(cuda_device, probs, dt):
    from gnn_motion_planning_b200 import collision
    from oracle import arm as o_arm
    maps = probs["snake7_maps"]
    boxes, ptr = o_arm.snake_boxes(maps)
    bd, pd = collision.pack_boxes(collision.snake_obstacles(maps), cuda_device)
    assert np.array_equal(bd.cpu().numpy(), boxes) and np.array_equal(pd.cpu().numpy(), ptr)
    d, lo, hi = collision.arm_model_info(collision.ARM_SNAKE7)
    assert d == 7
    rng = np.random.default_rng(8)
    n = 40000
    q = rng.uniform(lo * 1.01, hi * 1.01, (n, 7)).astype(dt)
    prob = rng.integers(0, len(maps), n).astype(np.int32)
    free, counted = collision.arm_state_fp(collision.ARM_SNAKE7, torch.from_numpy(q).to(cuda_device), bd, pd,
                                           torch.from_numpy(prob).to(cuda_device), want_counted=True)
    of, oc = o_arm.state_fp(o_arm.SNAKE7, q, boxes, ptr, prob)
    assert np.array_equal(free.cpu().numpy(), of) and np.array_equal(counted.cpu().numpy(), oc)
    assert 0.03 < of.mean() < 0.5
    m = 15000
    a = rng.uniform(lo, hi, (m, 7)).astype(dt)
    b = np.clip(a + rng.normal(0, 0.15, (m, 7)), lo, hi).astype(dt)
    free, checks = collision.arm_edge_fp(collision.ARM_SNAKE7, torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device),
                                         bd, pd, torch.from_numpy(prob[:m]).to(cuda_device), rrt_eps=0.1, want_checks=True)
    of, oc = o_arm.edge_fp(o_arm.SNAKE7, a, b, boxes, ptr, prob[:m], rrt_eps=0.1)
    assert np.array_equal(free.cpu().numpy(), of) and np.array_equal(checks.cpu().numpy(), oc)


def test_snake_and_ur5_env(cuda_device, probs):
    from gnn_motion_planning_b200.environment import SnakeEnv, UR5Env
    env = SnakeEnv(maps=probs["snake7_maps"].astype(np.float64), init_states=probs["snake7_start"], goal_states=probs["snake7_goal"])
    assert str(env) == "snake7" and env.config_dim == 7 and env.RRT_EPS == 0.1
    env.init_new_problem(3)
    assert env._state_fp(np.asarray(env.init_state)) and env._state_fp(np.asarray(env.goal_state))
    assert env.collision_check_count == 2 and env.obstacles.shape[1] == 2
    boxes, ptr = probs["ur5_boxes"], probs["ur5_box_ptr"]
    problems = [([(boxes[j, :3], boxes[j, 3:]) for j in range(ptr[i], ptr[i + 1])], probs["ur5_start"][i], probs["ur5_goal"][i], [])
                for i in range(len(ptr) - 1)]
    u = UR5Env(problems=problems)
    assert str(u) == "ur5" and u.config_dim == 6 and u.RRT_EPS == 0.1 and abs(np.max(u.bound) - 2 * np.pi) < 1e-9
    u.init_new_problem(5)
    assert u._state_fp(np.asarray(u.init_state, dtype=np.float64)) and u.in_goal_region(np.asarray(u.goal_state, dtype=np.float64))


@pytest.mark.parametrize("tag,model,dof,eps", [("kuka7", 0, 7, 0.5), ("kuka14", 1, 14, 0.5), ("kuka13", 2, 13, 0.5), ("ur5", 3, 6, 0.1),
                                               ("snake7", 4, 7, 0.1)])
def test_fast_graph_form_is_bit_identical(cuda_device, probs, tag, model, dof, eps):
    """The fp32-filter + exact-fixup graph kernel (default) against the fp64 thread-per-edge form and the C oracle: the same
    booleans and the same collision_check_count increments on k-NN graphs of every arm model (boxes, ground plane,
    self collision, arm-arm, > 32 boxes = the clustered path for the snake mazes)."""
    from gnn_motion_planning_b200 import collision, graph
    from oracle import arm as o_arm
    dev = cuda_device
    if tag == "snake7":
        boxes, ptr = o_arm.snake_boxes(probs["snake7_maps"])
    else:
        boxes, ptr = probs[tag + "_boxes"], probs[tag + "_box_ptr"]
    _, lo, hi = collision.arm_model_info(model)
    rng = np.random.default_rng(11)
    B, n, k = 6, 400, 12
    v = rng.uniform(lo, hi, (B * n, dof)).astype(np.float32)
    if tag == "snake7":          # keep the snake near its maze so that edges are short enough to be interesting
        v[:, :2] = rng.uniform(-9, 9, (B * n, 2)).astype(np.float32)
    v[[3, 401, 999]] *= 1.5      # nodes outside the joint limits
    vd = torch.from_numpy(v).to(dev)
    node_ptr = (np.arange(B + 1) * n).astype(np.int32)
    ei, edge_ptr = graph.knn_graph_batch(vd, node_ptr, np.full(B, n), np.full(B, k))
    et = int(edge_ptr[-1])
    bd, pd = torch.from_numpy(boxes).to(dev), torch.from_numpy(ptr).to(dev)
    pg_np = rng.integers(0, len(ptr) - 1, B).astype(np.int32)
    pg = torch.from_numpy(pg_np).to(dev)
    npd, epd = torch.from_numpy(node_ptr).to(dev), torch.from_numpy(edge_ptr).to(dev)
    out = {}
    for mode in ("fast", "exact"):
        f, c = collision.arm_edge_fp_graph(model, vd, ei, npd, epd, bd, pd, et, rrt_eps=eps, problem_of_graph=pg, want_checks=True,
                                           mode=mode)
        out[mode] = (f.cpu().numpy().copy(), c.cpu().numpy().copy())
    assert np.array_equal(out["fast"][0], out["exact"][0])
    assert np.array_equal(out["fast"][1], out["exact"][1])
    # and against the oracle on a sample of edges
    ei_np = ei[:, :et].cpu().numpy()
    gid = np.repeat(np.arange(B), np.diff(edge_ptr))
    pick = rng.choice(et, 4000, replace=False)
    a, b = v[ei_np[0, pick] + node_ptr[gid[pick]]], v[ei_np[1, pick] + node_ptr[gid[pick]]]
    of, oc = o_arm.edge_fp(model, a, b, boxes, ptr, pg_np[gid[pick]], rrt_eps=eps)
    assert np.array_equal(out["fast"][0][pick], of) and np.array_equal(out["fast"][1][pick], oc)
    free_frac = out["fast"][0].mean()
    undecided = collision.arm_last_undecided(dev)
    print("%s: %d edges, free fraction %.3f, states checked %d, undecided by the fp32 filter %d" % (tag, et, free_frac, int(out["fast"][1].sum()), undecided))
    assert 0.0 < free_frac < 1.0
