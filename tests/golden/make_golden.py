"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE'S OWN PYTHON CODE.

Run in the build container only (``/root/reference`` must exist; it does not on the GPU box):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference (imported / compiled from the files where they lie,
nothing is copied into this repo):
  * environment/maze_env.py  MazeEnv (_state_fp, _edge_fp, collision_check_count)  -- real NumPy code
  * model.py                 EncoderProcessDecoder.forward with shipped weights
  * model_smoother.py        ModelSmoother.forward with shipped weights
  * eval_gnn.py              create_data, explore  (function bodies compiled out of the file via ast,
                             because importing the module pulls pybullet/matplotlib)
  * smoother.py              model_smooth, obs_data, proposed_path_smootherv2 (same way)
Third-party primitives that are not installable here (torch_geometric / torch_cluster /
torch_sparse / torch_scatter) are replaced by tests/golden/_pyg_stubs.py, which restates their
published semantics -- so these vectors pin our oracle + kernels to the reference's python, and
are "unpinned" only at the PyG-primitive boundary.

Also copies the shipped weight files the parity tests and bench need into tests/golden/weights/
(binary fixtures, not source).
"""
import ast
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GNNMP_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import _pyg_stubs  # noqa: E402

_pyg_stubs.install()


def load_ref_maze_env():
    """Import reference environment/maze_env.py without environment/__init__.py (which needs pybullet)."""
    import importlib.util
    pkg = types.ModuleType("environment")
    pkg.__path__ = [os.path.join(REF, "environment")]
    sys.modules["environment"] = pkg
    for name in ("env_config", "maze_env", "timer"):
        spec = importlib.util.spec_from_file_location("environment." + name,
                                                      os.path.join(REF, "environment", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["environment." + name] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
    return sys.modules["environment.maze_env"].MazeEnv


def ref_functions(pyfile, names, namespace):
    """Compile selected top-level functions/classes out of a reference file into ``namespace``."""
    src = open(os.path.join(REF, pyfile)).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(keep) == len(names), (pyfile, names)
    mod = ast.Module(body=keep, type_ignores=[])
    exec(compile(mod, os.path.join(REF, pyfile), "exec"), namespace)
    return namespace


def sym_knn_edges(v, k):
    ei = _pyg_stubs.knn_graph(v, k, loop=True)
    ei = torch.cat([ei, ei.flip(0)], 1)
    return _pyg_stubs.coalesce(ei, None, len(v), len(v))[0]


def main():
    np.random.seed(20211206)   # MazeEnv.sample_n_points draws from the global NumPy RNG: keep the fixtures reproducible
    os.chdir(REF)  # MazeEnv opens 'maze_files/...' relative to cwd
    sys.path.insert(0, REF)
    MazeEnv = load_ref_maze_env()
    import model as ref_model
    import model_smoother as ref_model_smoother

    # ---------------------------------------------------------------- weights (binary fixtures)
    wdir = os.path.join(HERE, "weights")
    os.makedirs(wdir, exist_ok=True)
    for w in ("weights_maze.pt", "weights_kuka.pt", "kuka_14.pt", "weights_snake.pt", "weights_ur5.pt", "smooth_2d_attv3.pt",
              "smooth_7d_attv3.pt"):
        shutil.copyfile(os.path.join(REF, "data/weights", w), os.path.join(wdir, w))

    # ---------------------------------------------------------------- maze problems subset
    env = MazeEnv(dim=2)
    prob_ids = np.array([0, 1, 2, 3, 7, 100, 2000, 2001, 2002, 2003, 2004, 2005, 2500, 2999])
    np.savez_compressed(os.path.join(HERE, "maze_problems.npz"),
                        ids=prob_ids, maps=env.maps[prob_ids].astype(np.uint8),
                        init_states=env.init_states[prob_ids], goal_states=env.goal_states[prob_ids])
    # a larger map pool (maps only, uint8) for the synthetic bench / full-size property tests
    np.savez_compressed(os.path.join(HERE, "maze_maps_256.npz"), maps=env.maps[:256].astype(np.uint8))

    # ---------------------------------------------------------------- maze collision golden
    rng = np.random.default_rng(20211206)
    out = {}
    for dt in (np.float32, np.float64):
        S, SP, SF, SC = [], [], [], []
        A, B, EP, EF, EC, EK = [], [], [], [], [], []
        for pi, pid in enumerate(prob_ids):
            env.init_new_problem(int(pid))
            # states: uniform in [-1.1,1.1] (some out of range), grid-boundary values, exact +-1
            st = rng.uniform(-1.1, 1.1, (300, 2))
            edges_of_grid = (np.arange(0, 16) * 2.0 / 15 - 1.0)
            bd = np.stack([rng.choice(edges_of_grid, 60), rng.uniform(-1, 1, 60)], 1)
            bd2 = bd[:, ::-1] + rng.choice([0, 1e-7, -1e-7, 1e-16], (60, 1))
            st = np.concatenate([st, bd, bd2, [[1, 1], [-1, -1], [1, -1], [0, 0], [1.0000001, 0]]]).astype(dt)
            for s in st:
                c0 = env.collision_check_count
                f = bool(env._state_fp(s))
                S.append(s); SP.append(pi); SF.append(f); SC.append(env.collision_check_count - c0)
            # edges: pairs among free samples (kNN-like short ones and long ones) + random pairs
            free = np.array(env.sample_n_points(150)).astype(dt)
            allp = rng.uniform(-1.05, 1.05, (150, 2)).astype(dt)
            pairs = []
            for _ in range(500):
                i, j = rng.integers(0, len(free), 2)
                pairs.append((free[i], free[j]))
            d = ((free[:, None, :] - free[None, :, :]) ** 2).sum(-1)
            nn = np.argsort(d, axis=1)[:, :5]
            for i in range(len(free)):
                for j in nn[i]:
                    pairs.append((free[i], free[j]))
            for _ in range(250):
                i, j = rng.integers(0, len(allp), 2)
                pairs.append((allp[i], free[j % len(free)]))
                pairs.append((free[i % len(free)], allp[j]))
            for a, b in pairs:
                c0 = env.collision_check_count
                f = bool(env._edge_fp(a, b))
                A.append(a); B.append(b); EP.append(pi); EF.append(f)
                EC.append(env.collision_check_count - c0); EK.append(env.k)
        tag = "f32" if dt == np.float32 else "f64"
        out.update({
            "states_" + tag: np.array(S, dt), "state_problem_" + tag: np.array(SP, np.int32),
            "state_free_" + tag: np.array(SF, np.uint8), "state_counted_" + tag: np.array(SC, np.uint8),
            "edge_a_" + tag: np.array(A, dt), "edge_b_" + tag: np.array(B, dt),
            "edge_problem_" + tag: np.array(EP, np.int32), "edge_free_" + tag: np.array(EF, np.uint8),
            "edge_checks_" + tag: np.array(EC, np.int32), "edge_k_" + tag: np.array(EK, np.int32),
        })
    np.savez_compressed(os.path.join(HERE, "maze_collision.npz"), **out)
    print("maze_collision:", {k: v.shape for k, v in out.items() if k.startswith("edge_free")},
          "free frac", out["edge_free_f32"].mean())

    # ---------------------------------------------------------------- create_data golden (eval_gnn.py:150-165)
    ns = dict(torch=torch, np=np, Data=_pyg_stubs.Data, knn_graph=_pyg_stubs.knn_graph, coalesce=_pyg_stubs.coalesce)
    ref_functions("eval_gnn.py", ["create_data"], ns)
    cd = {}
    for tag, (c, nf, ncol, k) in {"maze2": (2, 102, 60, 10), "kuka7": (7, 90, 90, 12), "kuka14": (14, 150, 40, 8),
                                  "dup": (2, 40, 10, 10)}.items():
        free = [rng.uniform(-1, 1, c) for _ in range(nf)]
        collided = [rng.uniform(-1, 1, c) for _ in range(ncol)]
        if tag == "dup":  # duplicated points => exact distance ties
            free[5] = free[4].copy(); collided[3] = free[9].copy()
        fake_env = types.SimpleNamespace(goal_state=free[1])
        d = ns["create_data"](free, collided, fake_env, k)
        cd.update({tag + "_free": np.array(free), tag + "_collided": np.array(collided), tag + "_k": np.array(k),
                   tag + "_v": d.v.numpy(), tag + "_labels": d.labels.numpy(),
                   tag + "_edge_index": d.edge_index.numpy(), tag + "_goal": d.goal.numpy()})
        print("create_data", tag, d.edge_index.shape)
    np.savez_compressed(os.path.join(HERE, "create_data.npz"), **cd)

    # ---------------------------------------------------------------- explorer golden (model.py:115-150)
    ex = {}
    cfgs = {
        "maze2": dict(w="weights_maze.pt", c=2, e=32, s=2, ws=2, n=200, k=10, lo=-1.0, hi=1.0),
        "kuka7": dict(w="weights_kuka.pt", c=7, e=64, s=6, ws=3, n=150, k=8, lo=-2.9, hi=2.9),
        "kuka14": dict(w="kuka_14.pt", c=14, e=32, s=6, ws=3, n=160, k=8, lo=-2.9, hi=2.9),
    }
    for tag, cf in cfgs.items():
        m = ref_model.EncoderProcessDecoder(workspace_size=cf["ws"], config_size=cf["c"], embed_size=cf["e"],
                                            obs_size=cf["s"])
        m.load_state_dict(torch.load(os.path.join(REF, "data/weights", cf["w"]), map_location="cpu"))
        m.eval()
        v = torch.from_numpy(rng.uniform(cf["lo"], cf["hi"], (cf["n"], cf["c"])).astype(np.float32))
        ei = sym_knn_edges(v, cf["k"])
        if tag == "maze2":
            env.init_new_problem(2000)
            obs = torch.FloatTensor(env.obstacles)                     # [O,2]
        else:
            nb = 5
            obs = torch.from_numpy(np.concatenate([rng.uniform(0.1, 0.3, (nb, 1, 3)),
                                                   rng.uniform(-0.8, 0.8, (nb, 1, 3))], 1).astype(np.float32))  # [O,2,3]
        for loop in (1, 5):
            with torch.no_grad():
                dense = m(goal=v[1], loop=loop, v=v, obstacles=obs, free=None, collided=None, edge_index=ei,
                          labels=None)
            ex["%s_logits_loop%d" % (tag, loop)] = dense[ei[1], ei[0]].numpy()
            assert int((dense != 0).sum()) <= ei.shape[1]
        m.use_obstacles = False
        with torch.no_grad():
            dense = m(goal=v[1], loop=5, v=v, obstacles=obs, free=None, collided=None, edge_index=ei)
        ex[tag + "_logits_noobs"] = dense[ei[1], ei[0]].numpy()
        ex.update({tag + "_v": v.numpy(), tag + "_edge_index": ei.numpy(), tag + "_obstacles": obs.numpy(),
                   tag + "_goal": v[1].numpy()})
        print("explorer", tag, ei.shape, float(ex[tag + "_logits_loop5"].mean()), float(ex[tag + "_logits_loop5"].std()))
    np.savez_compressed(os.path.join(HERE, "explorer.npz"), **ex)

    # ---------------------------------------------------------------- smoother golden (model_smoother.py:104-142)
    sm = {}
    for tag, (w, c, p, nf, ncol) in {"2d": ("smooth_2d_attv3.pt", 2, 9, 70, 40), "7d": ("smooth_7d_attv3.pt", 7, 14, 60, 50),
                                     "2d_short": ("smooth_2d_attv3.pt", 2, 3, 12, 1)}.items():
        ms = ref_model_smoother.ModelSmoother(workspace_size=3, config_size=c, embed_size=128, obs_size=6)
        ms.load_state_dict(torch.load(os.path.join(REF, "data/weights", w), map_location="cpu"))
        ms.eval()
        lo, hi = (-1, 1) if c == 2 else (-2.9, 2.9)
        path = torch.from_numpy(np.cumsum(rng.uniform(-0.1, 0.15, (p, c)), 0).astype(np.float32)).clamp(lo, hi)
        free = torch.from_numpy(rng.uniform(lo, hi, (nf, c)).astype(np.float32))
        coll = torch.from_numpy(rng.uniform(lo, hi, (ncol, c)).astype(np.float32))
        e = torch.cat((torch.arange(1, p).reshape(1, -1), torch.arange(0, p - 1).reshape(1, -1)), 0)
        e = torch.cat((e, e.flip(0)), -1)
        e, _ = _pyg_stubs.add_self_loops(e, num_nodes=p)
        for loop in (1, 3):
            with torch.no_grad():
                newp = ms(path=path.clone(), free=free, collided=coll, obstacles=None, edge_index=e, loop=loop)
            sm["%s_out_loop%d" % (tag, loop)] = newp.numpy()
        sm.update({tag + "_path": path.numpy(), tag + "_free": free.numpy(), tag + "_collided": coll.numpy(),
                   tag + "_edge_index": e.numpy()})
        print("smoother", tag, newp.shape)
    np.savez_compressed(os.path.join(HERE, "smoother.npz"), **sm)

    # ---------------------------------------------------------------- arm problem subsets (data of maze_files/kukas_*.pkl)
    # obstacles + the PyBullet-verified free states the dataset ships (start, goal, solution waypoints): a sanity band
    # for the arm model (PyBullet itself cannot run here: arm parity is unpinned).
    import pickle
    arm = {}
    for tag, fn in (("kuka7", "kukas_7_3000.pkl"), ("kuka14", "kukas_14_3000.pkl"), ("kuka13", "kukas_13_3000.pkl"),
                    ("ur5", "ur5s_6_3000.pkl")):
        with open(os.path.join(REF, "maze_files", fn), "rb") as f:
            pr = pickle.load(f)[:48]
        boxes, ptr, known, known_p, ea, eb, ep = [], [0], [], [], [], [], []
        for i, (obs, st, go, path) in enumerate(pr):
            for h, b in obs:   # ur5 entries are ragged, e.g. [0.01, 0.01, array([0.84])]
                boxes.append(np.array([float(np.ravel(x)[0]) for x in list(h) + list(b)]))
            ptr.append(len(boxes))
            for q in [st, go] + list(path):
                known.append(np.asarray(q, np.float64).reshape(-1)); known_p.append(i)
            path = np.asarray(path, np.float64)
            for u, w in zip(path[:-1], path[1:]):
                ea.append(u); eb.append(w); ep.append(i)
        arm.update({tag + "_boxes": np.array(boxes), tag + "_box_ptr": np.array(ptr, np.int32),
                    tag + "_known_free": np.array(known), tag + "_known_free_problem": np.array(known_p, np.int32),
                    tag + "_path_a": np.array(ea), tag + "_path_b": np.array(eb), tag + "_path_problem": np.array(ep, np.int32),
                    tag + "_start": np.array([p[1] for p in pr]), tag + "_goal": np.array([p[2] for p in pr])})
        print("arm problems", tag, len(boxes), "boxes,", len(known), "known-free states")
    with np.load(os.path.join(REF, "maze_files", "snakes_15_2_3000.npz")) as f:     # snake: maps + PyBullet-free init / goal states
        arm.update({"snake7_maps": f["maps"][:48].astype(np.uint8), "snake7_start": f["init_states"][:48], "snake7_goal": f["goal_states"][:48]})
    np.savez_compressed(os.path.join(HERE, "arm_problems.npz"), **arm)

    # ---------------------------------------------------------------- end-to-end explore() golden: BASELINE config C1
    # reference eval_gnn.explore(batch=100, t_max=100, k=10, smoother='none') on real maze problems
    # (main.ipynb cell 8 / BASELINE.json configs[0]) with the reference MazeEnv and reference model.
    import time as _time
    ns = dict(torch=torch, np=np, Data=_pyg_stubs.Data, knn_graph=_pyg_stubs.knn_graph, coalesce=_pyg_stubs.coalesce,
              time=_time.time, device=torch.device("cpu"), loop=5)
    ns["DotDict"] = type("DotDict", (dict,), dict(__getattr__=dict.get, __setattr__=dict.__setitem__))
    ref_functions("eval_gnn.py", ["create_data", "explore", "obs_data", "to_np", "path_cost"], ns)
    m = ref_model.EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2)
    m.load_state_dict(torch.load(os.path.join(REF, "data/weights/weights_maze.pt"), map_location="cpu"))
    m.eval()

    # eval_gnn.py:202 `policy[np.array(explored_edges).reshape(2, -1)] = 0` relies on the torch<=1.x rule
    # that a short non-tuple SEQUENCE of sequences (here a (2,M) ndarray) is treated as a TUPLE of index
    # arrays (policy[rows, cols]); torch 2.x indexes ROWS instead, which zeroes row 0 and the search never
    # starts (success 0/6 here, vs 1000/1000 in the author's main.ipynb).  Restore the author-era rule
    # without touching the reference code: the model wrapper returns a Tensor subclass whose
    # __setitem__ re-applies the legacy interpretation.
    class LegacyIndexTensor(torch.Tensor):
        def __setitem__(self, idx, val):
            if isinstance(idx, np.ndarray) and idx.ndim == 2 and idx.shape[0] < 32:
                idx = tuple(torch.from_numpy(r) for r in idx)
            return super().__setitem__(idx, val)

    ref_forward = m.forward

    def legacy_forward(*a, **k):
        return ref_forward(*a, **k).as_subclass(LegacyIndexTensor)
    m.forward = legacy_forward
    e2e = {}
    ids = [2000, 2001, 2002, 2003, 2004, 2005]
    for pid in ids:
        np.random.seed(1234 + pid)
        env.init_new_problem(pid)
        r = ns["explore"](env, m, None, smooth=True, batch=100, t_max=100, k=10, smoother="none")
        e2e["p%d_success" % pid] = np.array(r["success"])
        e2e["p%d_c_explore" % pid] = np.array(r["c_explore"])
        e2e["p%d_explored" % pid] = np.array(r["explored"])
        e2e["p%d_path" % pid] = np.array(r["path"])
        e2e["p%d_n_nodes" % pid] = np.array(len(r["data"].v))
        print("explore", pid, r["success"], r["c_explore"], len(r["explored"]), len(r["data"].v))
    e2e["ids"] = np.array(ids)
    np.savez_compressed(os.path.join(HERE, "explore_c1.npz"), **e2e)


def more_goldens():
    """Round-2 fixtures: every (config, embed, obs) combination of reference str2name.py:12-66 that round 1 left
    uncovered, the remaining shipped smoother weights, and the a5 row (smoother.py:194-246 run for real).
    Own RNGs, own files: the round-1 fixtures above stay bit-identical."""
    os.chdir(REF)
    sys.path.insert(0, REF)
    MazeEnv = load_ref_maze_env()
    import model as ref_model
    import model_smoother as ref_model_smoother
    wdir = os.path.join(HERE, "weights")
    os.makedirs(wdir, exist_ok=True)
    for w in ("weights_kuka_13.pt", "weights_maze_3.pt", "smooth_13d_attv3.pt", "smooth_14d_attv3.pt",
              "smooth_snake_attv3.pt", "smooth_ur5_attv3.pt"):
        shutil.copyfile(os.path.join(REF, "data/weights", w), os.path.join(wdir, w))
    env = MazeEnv(dim=2)
    rng = np.random.default_rng(20261017)

    # ---------------------------------------------------------------- explorer (model.py:115-150): snake7, ur5, kuka13, maze3
    ex = {}
    cfgs = {
        "snake7": dict(w="weights_snake.pt", c=7, e=32, s=2, ws=3, n=180, k=9, lo=-1.0, hi=1.0),
        "ur5": dict(w="weights_ur5.pt", c=6, e=32, s=6, ws=3, n=170, k=8, lo=-3.1, hi=3.1),
        "kuka13": dict(w="weights_kuka_13.pt", c=13, e=32, s=6, ws=3, n=150, k=8, lo=-2.0, hi=2.0),
        "maze3": dict(w="weights_maze_3.pt", c=3, e=32, s=2, ws=2, n=160, k=8, lo=-1.0, hi=1.0),
    }
    for tag, cf in cfgs.items():
        m = ref_model.EncoderProcessDecoder(workspace_size=cf["ws"], config_size=cf["c"], embed_size=cf["e"],
                                            obs_size=cf["s"])
        m.load_state_dict(torch.load(os.path.join(REF, "data/weights", cf["w"]), map_location="cpu"))
        m.eval()
        v = torch.from_numpy(rng.uniform(cf["lo"], cf["hi"], (cf["n"], cf["c"])).astype(np.float32))
        ei = sym_knn_edges(v, cf["k"])
        if cf["s"] == 2:                                               # occupied cells of a real map (maze_env.py:73-79)
            env.init_new_problem(2001 if tag == "snake7" else 2003)
            obs = torch.FloatTensor(env.obstacles)
        else:                                                          # boxes (half extents, position), 2..12 of them
            nb = 12 if tag == "ur5" else 7
            obs = torch.from_numpy(np.concatenate([rng.uniform(0.05, 0.3, (nb, 1, 3)),
                                                   rng.uniform(-0.8, 0.8, (nb, 1, 3))], 1).astype(np.float32))
        for loop in (1, 5):
            with torch.no_grad():
                dense = m(goal=v[1], loop=loop, v=v, obstacles=obs, free=None, collided=None, edge_index=ei, labels=None)
            ex["%s_logits_loop%d" % (tag, loop)] = dense[ei[1], ei[0]].numpy()
        m.use_obstacles = False
        with torch.no_grad():
            dense = m(goal=v[1], loop=5, v=v, obstacles=obs, free=None, collided=None, edge_index=ei)
        ex[tag + "_logits_noobs"] = dense[ei[1], ei[0]].numpy()
        ex.update({tag + "_v": v.numpy(), tag + "_edge_index": ei.numpy(), tag + "_obstacles": obs.numpy(),
                   tag + "_goal": v[1].numpy()})
        print("explorer", tag, tuple(ei.shape), float(ex[tag + "_logits_loop5"].mean()), float(ex[tag + "_logits_loop5"].std()))
    np.savez_compressed(os.path.join(HERE, "explorer_more.npz"), **ex)

    # ---------------------------------------------------------------- smoother (model_smoother.py:104-142): 14d, 13d, ur5, snake
    sm = {}
    UR5_SCALE = 6.28318530718       # np.max(env.bound) of UR5Env = the largest joint limit of ur5/ur5.urdf:38 (str2name.py:40)
    for tag, (w, c, p, nf, ncol, scale, lo, hi) in {
            "14d": ("smooth_14d_attv3.pt", 14, 11, 80, 50, 1.0, -2.9, 2.9),
            "13d": ("smooth_13d_attv3.pt", 13, 8, 60, 60, 1.0, -2.0, 2.0),
            "ur5": ("smooth_ur5_attv3.pt", 6, 13, 90, 30, UR5_SCALE, -6.2, 6.2),
            "snake": ("smooth_snake_attv3.pt", 7, 10, 70, 45, 1.0, -1.0, 1.0)}.items():
        ms = ref_model_smoother.ModelSmoother(workspace_size=3, config_size=c, embed_size=128, obs_size=6, scale=scale)
        ms.load_state_dict(torch.load(os.path.join(REF, "data/weights", w), map_location="cpu"))
        ms.eval()
        step = (hi - lo) / 25
        path = torch.from_numpy(np.cumsum(rng.uniform(-step, 1.5 * step, (p, c)), 0).astype(np.float32)).clamp(lo, hi)
        free = torch.from_numpy(rng.uniform(lo, hi, (nf, c)).astype(np.float32))
        coll = torch.from_numpy(rng.uniform(lo, hi, (ncol, c)).astype(np.float32))
        e = torch.cat((torch.arange(1, p).reshape(1, -1), torch.arange(0, p - 1).reshape(1, -1)), 0)
        e = torch.cat((e, e.flip(0)), -1)
        e, _ = _pyg_stubs.add_self_loops(e, num_nodes=p)
        for loop in (1, 3):
            with torch.no_grad():
                newp = ms(path=path.clone(), free=free, collided=coll, obstacles=None, edge_index=e, loop=loop)
            sm["%s_out_loop%d" % (tag, loop)] = newp.numpy()
        sm.update({tag + "_path": path.numpy(), tag + "_free": free.numpy(), tag + "_collided": coll.numpy(),
                   tag + "_edge_index": e.numpy(), tag + "_scale": np.array(scale)})
        print("smoother", tag, tuple(newp.shape), float(np.abs(newp.numpy() - path.numpy()).max()))
    np.savez_compressed(os.path.join(HERE, "smoother_more.npz"), **sm)

    # ---------------------------------------------------------------- a5: model_smooth / proposed_path_smootherv2 (smoother.py:194-246)
    # The reference's own steering loop + smoother model + MazeEnv, on paths found by the reference's own explore().
    import time as _time
    from copy import deepcopy
    dot = type("DotDict", (dict,), dict(__getattr__=dict.get, __setattr__=dict.__setitem__))
    ns_s = dict(torch=torch, np=np, deepcopy=deepcopy, device=torch.device("cpu"), DotDict=dot,
                add_self_loops=_pyg_stubs.add_self_loops)
    ref_functions("smoother.py", ["obs_data", "proposed_path_smootherv2", "model_smooth"], ns_s)
    ns = dict(torch=torch, np=np, Data=_pyg_stubs.Data, knn_graph=_pyg_stubs.knn_graph, coalesce=_pyg_stubs.coalesce,
              time=_time.time, device=torch.device("cpu"), loop=5, DotDict=dot, model_smooth=ns_s["model_smooth"])
    ref_functions("eval_gnn.py", ["create_data", "explore", "obs_data", "to_np", "path_cost"], ns)
    m = ref_model.EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2)
    m.load_state_dict(torch.load(os.path.join(REF, "data/weights/weights_maze.pt"), map_location="cpu"))
    m.eval()
    ms = ref_model_smoother.ModelSmoother(workspace_size=2, config_size=2, embed_size=128, obs_size=6)
    ms.load_state_dict(torch.load(os.path.join(REF, "data/weights/smooth_2d_attv3.pt"), map_location="cpu"))
    ms.eval()

    class LegacyIndexTensor(torch.Tensor):     # eval_gnn.py:202 under the author's torch (see main())
        def __setitem__(self, idx, val):
            if isinstance(idx, np.ndarray) and idx.ndim == 2 and idx.shape[0] < 32:
                idx = tuple(torch.from_numpy(r) for r in idx)
            return super().__setitem__(idx, val)
    ref_forward = m.forward
    m.forward = lambda *a, **k: ref_forward(*a, **k).as_subclass(LegacyIndexTensor)

    # record every model_smooth call the reference makes (inputs, output, collision-check delta) ...
    calls = []
    ref_model_smooth = ns_s["model_smooth"]

    def recording_model_smooth(model, free, collided, old_path, env_, iter=5):
        rec = dict(free=np.array(free), collided=np.array(collided) if len(collided) else np.zeros((0, 2)),
                   path=np.array(old_path), c0=env_.collision_check_count)
        out = ref_model_smooth(model, free, collided, old_path, env_, iter=iter)
        rec["out"] = np.array(out)
        rec["checks"] = env_.collision_check_count - rec["c0"]
        calls.append(rec)
        return out
    ns["model_smooth"] = recording_model_smooth
    # ... and every steering call inside it
    steer = []
    ref_steer = ns_s["proposed_path_smootherv2"]

    def recording_steer(old_path, new_path, env_):
        c0 = env_.collision_check_count
        out = ref_steer(old_path, new_path, env_)
        steer.append(dict(old=np.array(old_path), new=np.array(new_path), out=np.array(out),
                          checks=env_.collision_check_count - c0, old_dtype=str(np.array(old_path[1]).dtype)))
        return out
    ns_s["proposed_path_smootherv2"] = recording_steer

    a5 = {}
    ids = [2000, 2001, 2002, 2003, 2004, 2005]
    for pid in ids:
        np.random.seed(4321 + pid)
        env.init_new_problem(pid)
        n_before, s_before = len(calls), len(steer)
        r = ns["explore"](env, m, ms, smooth=True, batch=100, t_max=100, k=10, smoother="model")
        a5["p%d_success" % pid] = np.array(r["success"])
        a5["p%d_c_explore" % pid] = np.array(r["c_explore"])
        a5["p%d_c_smooth" % pid] = np.array(r["c_smooth"])
        a5["p%d_path" % pid] = np.array(r["path"])
        a5["p%d_smooth_path" % pid] = np.array(r["smooth_path"])
        if len(calls) > n_before:
            c = calls[-1]
            a5.update({"p%d_ms_free" % pid: c["free"], "p%d_ms_collided" % pid: c["collided"], "p%d_ms_path" % pid: c["path"],
                       "p%d_ms_out" % pid: c["out"], "p%d_ms_checks" % pid: np.array(c["checks"])})
            for j, st in enumerate(steer[s_before:]):
                a5.update({"p%d_steer%d_old" % (pid, j): st["old"], "p%d_steer%d_new" % (pid, j): st["new"],
                           "p%d_steer%d_out" % (pid, j): st["out"], "p%d_steer%d_checks" % (pid, j): np.array(st["checks"])})
            a5["p%d_n_steer" % pid] = np.array(len(steer) - s_before)
        print("a5 explore+smooth", pid, r["success"], r["c_explore"], r["c_smooth"], len(r["path"]),
              "cost %.4f -> %.4f" % (ns["path_cost"](r["path"]), ns["path_cost"](r["smooth_path"])) if r["success"] else "")
    a5["ids"] = np.array(ids)
    np.savez_compressed(os.path.join(HERE, "model_smooth.npz"), **a5)

    # ---------------------------------------------------------------- explore() with RESAMPLING rounds (eval_gnn.py:235-247)
    # small batches so that the first graphs are exhausted: the tree, the explored-edge list (with the reshape(2,-1) quirk of
    # eval_gnn.py:202) and the counters carry over to the next graph.  smoother='none'.
    ns["model_smooth"] = ns_s["model_smooth"]
    mr = {}
    cases = [(pid, b, t, kk) for pid in (2000, 2001, 2002, 2003, 2004, 2005) for (b, t, kk) in ((25, 200, 6), (40, 300, 8))]
    for pid, b, t, kk in cases:
        np.random.seed(777 + pid)
        env.init_new_problem(pid)
        r = ns["explore"](env, m, None, smooth=True, batch=b, t_max=t, k=kk, smoother="none")
        tag = "p%d_b%d" % (pid, b)
        mr[tag + "_success"] = np.array(r["success"])
        mr[tag + "_c_explore"] = np.array(r["c_explore"])
        mr[tag + "_explored"] = np.array(r["explored"])
        mr[tag + "_path"] = np.array(r["path"]) if r["success"] else np.zeros((0, 2), np.float32)
        mr[tag + "_n_nodes"] = np.array(len(r["data"].v))
        mr[tag + "_n_explored_edges"] = np.array(len(r["explored_edges"]))
        print("explore multi-round", tag, r["success"], r["c_explore"], len(r["explored"]), len(r["data"].v))
    mr["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "explore_rounds.npz"), **mr)

    # ---------------------------------------------------------------- construct_graph (algorithm/dijkstra.py:15-31): kNN(5) + every edge checked
    from collections import defaultdict
    ns_c = dict(torch=torch, np=np, knn_graph=_pyg_stubs.knn_graph, coalesce=_pyg_stubs.coalesce, defaultdict=defaultdict,
                INFINITY=float("inf"))
    ref_functions("algorithm/dijkstra.py", ["construct_graph"], ns_c)
    cg = {}
    for pid in (2000, 2003):
        np.random.seed(555 + pid)
        env.init_new_problem(pid)
        pts = np.array([env.init_state, env.goal_state] + list(env.sample_n_points(300)))      # float64, as the dataset builder feeds it
        c0 = env.collision_check_count
        edge_cost, neighbors, edge_index, edge_free = ns_c["construct_graph"](env, pts)
        tag = "p%d" % pid
        cg[tag + "_points"] = pts
        cg[tag + "_edge_index"] = np.asarray(edge_index)
        cg[tag + "_edge_free"] = np.array(edge_free)
        cg[tag + "_checks"] = np.array(env.collision_check_count - c0)
        cg[tag + "_cost_flat"] = np.concatenate([np.asarray(edge_cost[i], np.float64) for i in range(len(pts))])
        cg[tag + "_nbr_flat"] = np.concatenate([np.asarray(neighbors[i], np.int64) for i in range(len(pts))])
        cg[tag + "_deg"] = np.array([len(neighbors[i]) for i in range(len(pts))])
        print("construct_graph", tag, np.asarray(edge_index).shape, float(np.mean(edge_free)), int(cg[tag + "_checks"]))
    cg["ids"] = np.array([2000, 2003])
    np.savez_compressed(os.path.join(HERE, "construct_graph.npz"), **cg)

    # ---------------------------------------------------------------- 3-D stick maze (maze_env.py:245-264, 279-291, 327-347)
    # MazeEnv(dim=3): state = (x, y, theta), a stick of length 0.2 centred at (x, y); _state_fp = both end points free + the
    # bisection of the stick; _edge_fp = K = int(d / 0.015) interpolated poses, each stick checked as a 2-D edge.
    env3 = MazeEnv(dim=3)
    ids3 = np.array([0, 1, 2, 5, 2000, 2001, 2002, 2003])
    m3 = {"ids": ids3, "maps": env3.maps[ids3].astype(np.uint8), "init_states": env3.init_states[ids3],
          "goal_states": env3.goal_states[ids3]}
    rng3 = np.random.default_rng(333)
    for dt in (np.float32, np.float64):
        S, SP, SF, SC, SK = [], [], [], [], []
        A, B, EP, EF, EC, EK = [], [], [], [], [], []
        for pi, pid in enumerate(ids3):
            env3.init_new_problem(int(pid))
            st = np.concatenate([rng3.uniform(-1.05, 1.05, (250, 2)), rng3.uniform(-0.42, 0.42, (250, 1))], 1)
            st = np.concatenate([st, [[0, 0, 0.4], [0, 0, -0.4], [1, 1, 0], [-1, -1, 0.2], [0.5, 0.5, 0.40000001]]]).astype(dt)
            for s_ in st:
                c0 = env3.collision_check_count
                f = bool(env3._state_fp(s_))
                S.append(s_); SP.append(pi); SF.append(f); SC.append(env3.collision_check_count - c0); SK.append(env3.k)
            np.random.seed(900 + int(pid))
            free = np.array(env3.sample_n_points(60)).astype(dt)
            pairs = []
            d = ((free[:, None, :2] - free[None, :, :2]) ** 2).sum(-1)
            nn = np.argsort(d, axis=1)[:, :4]
            for i in range(len(free)):
                for j in nn[i]:
                    pairs.append((free[i], free[j]))
            for _ in range(60):
                i, j = rng3.integers(0, len(free), 2)
                pairs.append((free[i], free[j]))
            for _ in range(20):                                   # theta wrap-around (|dtheta| > 0.4) and out-of-range states
                i = rng3.integers(0, len(free))
                q = free[i].copy(); q[2] = -q[2] if abs(q[2]) > 0.25 else (0.39 if q[2] < 0 else -0.39)
                pairs.append((free[i], q.astype(dt)))
                pairs.append((free[i], (free[i] + np.array([0.03, -0.02, 0.9])).astype(dt)))
            for a_, b_ in pairs:
                c0 = env3.collision_check_count
                f = bool(env3._edge_fp(a_.copy(), b_.copy()))
                A.append(a_); B.append(b_); EP.append(pi); EF.append(f)
                EC.append(env3.collision_check_count - c0); EK.append(env3.k)
        tag = "f32" if dt == np.float32 else "f64"
        m3.update({"states_" + tag: np.array(S, dt), "state_problem_" + tag: np.array(SP, np.int32),
                   "state_free_" + tag: np.array(SF, np.uint8), "state_checks_" + tag: np.array(SC, np.int32),
                   "state_k_" + tag: np.array(SK, np.int32),
                   "edge_a_" + tag: np.array(A, dt), "edge_b_" + tag: np.array(B, dt), "edge_problem_" + tag: np.array(EP, np.int32),
                   "edge_free_" + tag: np.array(EF, np.uint8), "edge_checks_" + tag: np.array(EC, np.int32),
                   "edge_k_" + tag: np.array(EK, np.int32)})
        print("maze3", tag, len(S), "states free", np.mean(SF), len(A), "edges free", np.mean(EF), "mean checks", np.mean(EC))
    np.savez_compressed(os.path.join(HERE, "maze3_collision.npz"), **m3)

    # ---------------------------------------------------------------- explore() on the 3-D stick maze (str2name 'maze3', smoother='none')
    m3x = ref_model.EncoderProcessDecoder(workspace_size=2, config_size=3, embed_size=32, obs_size=2)
    m3x.load_state_dict(torch.load(os.path.join(REF, "data/weights/weights_maze_3.pt"), map_location="cpu"))
    m3x.eval()
    ref_forward3 = m3x.forward
    m3x.forward = lambda *a, **k: ref_forward3(*a, **k).as_subclass(LegacyIndexTensor)
    ns["model_smooth"] = ns_s["model_smooth"]
    e3 = {}
    ids = [2000, 2001, 2002, 2003]
    for pid in ids:
        np.random.seed(31 + pid)
        env3.init_new_problem(pid)
        r = ns["explore"](env3, m3x, None, smooth=True, batch=100, t_max=300, k=10, smoother="none")
        e3["p%d_success" % pid] = np.array(r["success"])
        e3["p%d_c_explore" % pid] = np.array(r["c_explore"])
        e3["p%d_explored" % pid] = np.array(r["explored"])
        e3["p%d_path" % pid] = np.array(r["path"]) if r["success"] else np.zeros((0, 3), np.float32)
        e3["p%d_n_nodes" % pid] = np.array(len(r["data"].v))
        print("explore maze3", pid, r["success"], r["c_explore"], len(r["explored"]), len(r["data"].v))
    e3["ids"] = np.array(ids)
    np.savez_compressed(os.path.join(HERE, "explore_maze3.npz"), **e3)


if __name__ == "__main__":
    if "--more-only" not in sys.argv:
        main()
    more_goldens()
