#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2|C3|C4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic planning problems:
    k-NN RGG construction (gmp_knn_graph)  ->  explorer forward (gmp_explorer_forward)  ->
    collision check of every edge of every graph (gmp_maze_edge_fp_graph; maze workloads)
Workload at N=1 is BASELINE.json configs[1]: 2-D maze, batch = 256 problems, 1000-node k=50 RGG,
shipped weights_maze.pt, loop = 5, obstacles = occupied cells of real maze maps.
Multi-GPU: the batch of independent problems is sharded, 256 problems per rank (weak scaling); the only
collective is the all-gather of the per-problem result rows at the end of a step.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference's PyG path cannot be
installed here) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    # name: env, c, e, s, ws, N, k, batch per GPU, weights
    "C2": dict(env="maze2", c=2, e=32, s=2, ws=2, n=1000, k=50, batch=256, weights="weights_maze.pt", lo=-1.0, hi=1.0),
    "C3": dict(env="kuka7", c=7, e=64, s=6, ws=3, n=1000, k=50, batch=256, weights="weights_kuka.pt", lo=-2.9, hi=2.9),
    "C4": dict(env="kuka14", c=14, e=32, s=6, ws=3, n=2000, k=50, batch=128, weights="kuka_14.pt", lo=-2.9, hi=2.9),
    # C5 = BASELINE.json configs[4]: mixed env sweep, 256 problems per GPU split over the five environments
    "C5": dict(env="mixed", parts=[("maze2", 52), ("snake7", 51), ("ur5", 51), ("kuka7", 51), ("kuka14", 51)], n=1000, k=50, batch=256),
}
PARTS = {
    "maze2": dict(env="maze2", c=2, e=32, s=2, ws=2, weights="weights_maze.pt", lo=-1.0, hi=1.0),
    "snake7": dict(env="snake7", c=7, e=32, s=2, ws=3, weights="weights_snake.pt"),
    "ur5": dict(env="ur5", c=6, e=32, s=6, ws=3, weights="weights_ur5.pt"),
    "kuka7": dict(env="kuka7", c=7, e=64, s=6, ws=3, weights="weights_kuka.pt"),
    "kuka14": dict(env="kuka14", c=14, e=32, s=6, ws=3, weights="kuka_14.pt"),
}


def make_problem(wl, g):
    """Synthetic problem g of the workload: node 0 = init, node 1 = goal, all nodes free."""
    rng = np.random.default_rng(1234 + g)
    v = rng.uniform(wl["lo"], wl["hi"], (wl["n"], wl["c"])).astype(np.float32)
    if wl["s"] == 2:
        maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
        occ = np.argwhere(maps[g % len(maps)] == 1)
        obs = (occ / 15.0 - 0.5).astype(np.float32)
    else:
        ap = _arm_problems()
        tag = wl["env"]
        ptr = ap[tag + "_box_ptr"]
        pi = g % (len(ptr) - 1)
        obs = ap[tag + "_boxes"][ptr[pi]:ptr[pi + 1]].astype(np.float32)      # [O,6] = (halfExtents, basePosition)
        lo, hi = ARM_LIMITS[tag]
        v = rng.uniform(lo, hi, (wl["n"], wl["c"])).astype(np.float32)
    return v, obs


_AP = {}
ARM_LIMITS = {"kuka7": (np.array([-2.96705972839, -2.09439510239, -2.96705972839, -2.09439510239, -2.96705972839, -2.09439510239, -3.05432619099]),
                        np.array([2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 3.05432619099]))}
ARM_LIMITS["kuka14"] = (np.tile(ARM_LIMITS["kuka7"][0], 2), np.tile(ARM_LIMITS["kuka7"][1], 2))
ARM_LIMITS["ur5"] = (np.array([-2 * np.pi, -2 * np.pi, -np.pi, -2 * np.pi, -2 * np.pi, -2 * np.pi]),
                     np.array([2 * np.pi, 2 * np.pi, np.pi, 2 * np.pi, 2 * np.pi, 2 * np.pi]))
ARM_LIMITS["snake7"] = (np.array([-9, -9] + [-np.pi] * 5), np.array([9, 9] + [np.pi] * 5))
ARM_MODEL = {"kuka7": 0, "kuka14": 1, "ur5": 3, "snake7": 4}
ARM_EPS = {"kuka7": 0.5, "kuka14": 0.5, "ur5": 0.1, "snake7": 0.1}


def _arm_problems():
    if "d" not in _AP:
        _AP["d"] = np.load(os.path.join(G, "arm_problems.npz"))
    return _AP["d"]


def run_mixed(args, wl, rank, local_rank, world):
    """BASELINE.json configs[4]: mixed environment sweep.  Every rank runs one HotPath per environment over its share of the
    problems; a step = all five sub-batches (graph build + forward + all-edge collision check each)."""
    import torch
    import torch.distributed as dist
    from gnn_motion_planning_b200 import _lib, collision, shard
    from gnn_motion_planning_b200.batch import HotPath
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    _lib.load()
    dev = torch.device("cuda", local_rank)
    ap = _arm_problems()
    maps_np = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    N, k = wl["n"], wl["k"]
    subs = []
    first = rank * wl["batch"]
    for env, B in wl["parts"]:
        pw = PARTS[env]
        model = EncoderProcessDecoder(workspace_size=pw["ws"], config_size=pw["c"], embed_size=pw["e"], obs_size=pw["s"]).to(dev)
        model.load_state_dict(torch.load(os.path.join(G, "weights", pw["weights"]), map_location="cpu"))
        model.set_timing(True)
        vs, obss, probs = [], [], []
        for g in range(B):
            rng = np.random.default_rng(1234 + first + g)
            if env == "maze2":
                vs.append(rng.uniform(-1, 1, (N, 2)).astype(np.float32))
                pi = (first + g) % len(maps_np)
                obss.append((np.argwhere(maps_np[pi] == 1) / 15.0 - 0.5).astype(np.float32))
            else:
                lo, hi = ARM_LIMITS[env]
                vs.append(rng.uniform(lo, hi, (N, pw["c"])).astype(np.float32))
                if env == "snake7":
                    pi = (first + g) % len(ap["snake7_maps"])
                    obss.append((np.argwhere(ap["snake7_maps"][pi] == 1) / 15.0 - 0.5).astype(np.float32))
                else:
                    ptr = ap[env + "_box_ptr"]
                    pi = (first + g) % (len(ptr) - 1)
                    obss.append(ap[env + "_boxes"][ptr[pi]:ptr[pi + 1]].astype(np.float32))
            probs.append(pi)
        if env == "maze2":
            hp = HotPath(model, B, N, k, kind="maze", maps=torch.from_numpy(maps_np).to(dev), first_problem_id=first, device=dev)
        else:
            if env == "snake7":
                boxes_d, ptr_d = collision.pack_boxes(collision.snake_obstacles(ap["snake7_maps"]), dev)
            else:
                boxes_d, ptr_d = torch.from_numpy(ap[env + "_boxes"]).to(dev), torch.from_numpy(ap[env + "_box_ptr"]).to(dev)
            hp = HotPath(model, B, N, k, kind="arm", boxes=boxes_d, box_ptr=ptr_d, arm_model=ARM_MODEL[env], rrt_eps=ARM_EPS[env],
                         first_problem_id=first, device=dev)
        first += B
        v_h = torch.from_numpy(np.concatenate(vs)).pin_memory()
        subs.append(dict(env=env, B=B, hp=hp, model=model, v_h=v_h, goal_h=torch.from_numpy(np.stack([x[1] for x in vs])).pin_memory(),
                         obs_h=torch.from_numpy(np.concatenate(obss)).pin_memory(),
                         obs_ptr=np.cumsum([0] + [len(o) for o in obss]).astype(np.int32),
                         prob_h=torch.from_numpy(np.array(probs, np.int32)).pin_memory()))
    for s in subs:
        s["dev_in"] = [s[k_].to(dev) for k_ in ("v_h", "goal_h", "obs_h", "prob_h")]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    per_env_ms, edges, checks = {}, {}, {}

    def step(timed=False):
        for s in subs:
            a, b = ev(), ev()
            a.record()
            bufs = s["hp"].compute(s["dev_in"][0], s["dev_in"][1], s["dev_in"][2], s["obs_ptr"], s["dev_in"][3])
            b.record()
            if world > 1:
                shard.gather_result_rows(bufs["rows"], out=s.setdefault("gather", torch.empty((world * s["B"], 4), device=dev)), equal_shards=True)
            if timed:
                b.synchronize()
                per_env_ms[s["env"]] = per_env_ms.get(s["env"], 0.0) + a.elapsed_time(b)
                edges[s["env"]] = bufs["et"]
                checks[s["env"]] = int(bufs["checks"][:bufs["et"]].sum())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps, finish=None):
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            fn()
        if finish:
            finish()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # the timed region runs the bare loop; the per-environment breakdown (an event synchronise and a checks read-back per
    # environment, i.e. host gaps that are not part of the path) comes from a separate pass afterwards
    ms_step = timed_region(lambda: step(False), args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(args.steps):
        step(True)
    pending, io = [], {"h2d": 0, "d2h": 0}

    def e2e_step():
        io["h2d"] = io["d2h"] = 0
        for s in subs:
            t = s["hp"].submit(s["v_h"], s["goal_h"], s["obs_h"], s["obs_ptr"], s["prob_h"])
            pending.append(t)
            io["h2d"] += t["h2d_bytes"]
            io["d2h"] += t["d2h_bytes"]
        while len(pending) > len(subs):
            HotPath.wait(pending.pop(0))

    def e2e_finish():
        while pending:
            HotPath.wait(pending.pop(0))

    for _ in range(2):
        e2e_step()
    e2e_finish()
    ms_e2e = timed_region(e2e_step, args.steps, finish=e2e_finish)
    if rank != 0:
        return None
    K = args.steps
    Btot = wl["batch"]
    line = {
        "metric": "explorer_graphs_per_sec", "value": Btot * world / (ms_step / 1e3), "unit": "graphs/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C5: mixed env sweep %s, %d-node k=%d RGG, loop=5, shipped weights" % (wl["parts"], N, k),
                   "step": "per environment: knn_graph + explorer_forward + edge collision check of all edges",
                   "global_batch": Btot * world, "parallelism": "dp%d" % world, "l2": "working set >> 126 MB L2"},
        "per_env_ms_per_step": {k_: v_ / K for k_, v_ in per_env_ms.items()},
        "per_env_edges": edges, "per_env_state_checks_per_edge": {k_: checks[k_] / max(edges[k_], 1) for k_ in edges},
        "roofline": None, "cpu_baseline": None,
        "e2e": {"value": Btot * world / (ms_e2e / 1e3), "unit": "graphs/s", "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                "ms_per_step": ms_e2e},
        "gpu_launches": K * len(subs) * 27, "clocks": clocks,
        "note": "mixed sweep: roofline / cpu_baseline are reported by the single-environment workloads (C2, C3, C4)",
    }
    return line


def ref_oplist_flops(n, e_cnt, o, c, e, s, loop=5):
    from oracle.explorer import flops_per_graph
    return flops_per_graph(n, e_cnt, o, c, e, s, loop)


def edge_feature_flops(e_cnt, o, c, e):
    """Algorithmic FLOPs of what ONE edge_feature_kernel launch produces, reference formulation (model.py:120,123,
    130 + the edge columns of lin_0[0] and policy[0]), excluding the obstacle-token side."""
    enc = 2 * 2 * e_cnt * (2 * c * e + e * e)
    blk = 2 * (3 * e_cnt * e * e + 2 * e_cnt * (o + 1) * e + 2 * e_cnt * e * e)
    return enc + 3 * blk + 2 * e_cnt * 2 * e * e + 2 * e_cnt * e * e


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference) on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_graphs_per_sec(wl, graph_ids, threads):
    import torch
    from oracle import explorer as o_explorer, knn_graph as o_knn, maze as o_maze
    torch.set_num_threads(threads)
    sd = torch.load(os.path.join(G, "weights", wl["weights"]), map_location="cpu")
    maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    t0 = time.perf_counter()
    n_edges = 0
    for g in graph_ids:
        v, obs = make_problem(wl, g)
        ei = o_knn.knn_graph_edges(v, wl["n"], wl["k"])
        o_explorer.explorer_forward(sd, torch.from_numpy(v), torch.from_numpy(ei), torch.from_numpy(v[1]),
                                    torch.from_numpy(obs), loop=5, dense=False)
        if wl["env"] == "maze2":
            o_maze.edge_fp(v[ei[0]], v[ei[1]], maps, np.full(ei.shape[1], g % len(maps), np.int32))
        else:
            from oracle import arm as o_arm
            ap = _arm_problems()
            tag = wl["env"]
            ptr = ap[tag + "_box_ptr"]
            o_arm.edge_fp(ARM_MODEL[tag], v[ei[0]], v[ei[1]], ap[tag + "_boxes"], ptr,
                          np.full(ei.shape[1], g % (len(ptr) - 1), np.int32), rrt_eps=0.5)
        n_edges += ei.shape[1]
    dt = time.perf_counter() - t0
    return len(graph_ids) / dt, dt, n_edges


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    per_step = args.ref_graphs_per_step
    for w in range(args.warmup):
        cpu_graphs_per_sec(wl, [w % 4], threads)
    t = []
    for k in range(args.steps):
        gps, dt, _ = cpu_graphs_per_sec(wl, list(range(k * per_step, (k + 1) * per_step)), threads)
        t.append(dt)
    ms = 1000.0 * float(np.mean(t))
    value = per_step / (ms / 1000.0)
    sample = "%d of the %d graphs of the workload per step (knn graph + explorer forward + edge checks, oracle port)" % (
        per_step, wl["batch"])
    line = {
        "impl": "reference", "metric": "explorer_graphs_per_sec", "value": value, "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, world),
        "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference PyG path not installable (torch_geometric/torch_cluster/torch_sparse absent, no network): "
                "timed the oracle restatement of eval_gnn.create_data + model.py forward + maze_env._edge_fp on the host cores",
    }
    _emit(line)


def workload_config(args, wl, world):
    return {"workload": "%s: %s, batch=%d problems/GPU, %d-node k=%d RGG, loop=5, shipped %s" % (
        args.workload, wl["env"], wl["batch"], wl["n"], wl["k"], wl["weights"]),
        "step": "knn_graph + explorer_forward + edge collision check of all edges" + (
            " + smoother forward x5 on a 24-waypoint path per problem" if wl["env"] == "kuka7" and wl["n"] >= 1000 else ""),
        "global_batch": wl["batch"] * world,
        "parallelism": "dp%d (independent problems sharded, all-gather of result rows only)" % world,
        "l2": "per-step working set (~3.8 GB of edge features) >> 126 MB L2; no explicit flush needed"}


# --------------------------------------------------------------------------------------------------------------
def _emit(line):
    """Write the ONE JSON line to the real stdout (fd saved before library chatter was redirected to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries (NCCL "version" banner, torch warnings) print to fd 1; keep stdout for the JSON line only.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-graphs-per-step", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="only the headline workload (no C3/C4/C5/planner sub-records)")
    ap.add_argument("--sub-steps", type=int, default=5)
    ap.add_argument("--ef-mode", default="auto", choices=["auto", "tc", "tc4", "tcrd", "simt"], help="edge-feature kernel variant (A/B profiling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.workload == "C5":
            if rank == 0:
                _emit({"impl": "reference", "unavailable": "mixed sweep: run --workload C2/C3/C4 for the per-environment CPU port baseline"})
            return
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gnn_motion_planning_b200 import _lib
    _lib.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _pin_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def run(name, steps, warmup, cpu_baseline):
        a = argparse.Namespace(**vars(args))
        a.workload, a.steps, a.warmup, a.no_cpu_baseline = name, steps, warmup, not cpu_baseline
        line = run_mixed(a, WORKLOADS[name], rank, local_rank, world) if name == "C5" else run_single(a, WORKLOADS[name], rank, local_rank, world)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        return line

    line = run(args.workload, args.steps, args.warmup, world == 1 and not args.no_cpu_baseline)
    if args.workload == "C2" and not args.no_sub_records:
        # BASELINE.json configs[2..4] and the planner loop (configs[0] batched) as sub-records of the same JSON line.  C4 / C5 are
        # the configs BASELINE defines as multi-GPU: at N ranks they run 128 / 256 problems per rank (1 024 / 2 048 at N = 8).
        subs = {}
        names = ["C3", "C4", "C5"] if world == 1 else ["C4", "C5"]
        for name in names:
            sub = run(name, args.sub_steps, 3, world == 1 and not args.no_cpu_baseline and name != "C5")
            if rank == 0:
                subs[name] = {k_: v_ for k_, v_ in sub.items() if k_ not in ("gpu_launches_note",)}
        pl = run_planner(args, rank, local_rank, world)
        if rank == 0:
            subs["C1_planner"] = pl
            line["sub_records"] = subs
            line["gpu_launches"] += sum(int(v_.get("gpu_launches", 0)) for v_ in subs.values())
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _pin_to_gpu_numa_node(local_rank):
    """Bind this rank's host threads (and therefore its pinned staging buffers, first-touch) to the CPUs next to its GPU:
    round 1's 8-rank run had every rank on NUMA node 0 and lost a third of its end-to-end throughput to the read-back."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w in range(n_words) for b in range(64) if (int(mask[w]) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def run_planner(args, rank, local_rank, world):
    """BASELINE.json configs[0] batched: explore(batch=100, t_max=100, k=10, smoother='none') (main.ipynb cell 8) for 256 maze
    problems per rank through gnn_motion_planning_b200.search.explore_batch -- sampling, graph build, forward and the lazy tree
    search on the device (gmp_maze_tree_search), host code only between rounds.  Problems: the 256 real maps of maze_maps_256.npz
    with init / goal drawn like MazeEnv.set_random_init_goal (seeded).  Reported beside the repo's own host mirror of the
    reference loop (one problem at a time, one edge per Python iteration, same kernels) on a sample of the same problems."""
    import torch
    import torch.distributed as dist
    from gnn_motion_planning_b200 import search
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    dev = torch.device("cuda", local_rank)
    maps = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    P = len(maps)
    rng = np.random.default_rng(4242)
    init, goal = np.zeros((P, 2)), np.zeros((P, 2))
    for p in range(P):            # free cell centres, jittered: init != goal
        free = np.argwhere(maps[p] == 0)
        i, j = rng.choice(len(free), 2, replace=False)
        init[p] = (free[i] + rng.uniform(0.3, 0.7, 2)) * 2.0 / 15 - 1.0
        goal[p] = (free[j] + rng.uniform(0.3, 0.7, 2)) * 2.0 / 15 - 1.0
    model = EncoderProcessDecoder(workspace_size=2, config_size=2, embed_size=32, obs_size=2).to(dev)
    model.load_state_dict(torch.load(os.path.join(G, "weights", "weights_maze.pt"), map_location="cpu"))
    ids = list(range(P))
    seeds = [1234 + rank * P + p for p in ids]
    out = {}
    for spec_k in (1, 8):
        search.explore_batch(model, maps, init, goal, ids[:32], seeds[:32], batch=100, t_max=100, k=10, spec_k=spec_k, device=dev)   # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        best = None
        for _rep in range(3):               # host code between the launches: take the best of three
            t0 = time.perf_counter()
            res = search.explore_batch(model, maps, init, goal, ids, seeds, batch=100, t_max=100, k=10, spec_k=spec_k, device=dev)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0) if best is not None else time.perf_counter() - t0
        dt = torch.tensor([best], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[spec_k] = (float(dt), res)
    # the only collective of the planner run: the per-problem rows (id, success, path cost, checks, ...) of eval_gnn.py:120-134
    from gnn_motion_planning_b200 import shard
    rows_local = torch.tensor(np.stack([r["row"] for r in out[1][1]]), dtype=torch.float32, device=dev)
    rows_local[:, 0] += rank * P
    summary = shard.summarize_search(shard.gather_result_rows(rows_local, equal_shards=True))
    tm = {}
    search.explore_batch(model, maps, init, goal, ids, seeds, batch=100, t_max=100, k=10, spec_k=1, device=dev, timings=tm)
    tmd = {}
    dt_dev = None
    for _rep in range(3):
        t0 = time.perf_counter()
        res_dev = search.explore_batch(model, maps, init, goal, ids, seeds, batch=100, t_max=100, k=10, spec_k=1, device=dev, sampler="device")
        torch.cuda.synchronize()
        dt_dev = min(dt_dev, time.perf_counter() - t0) if dt_dev is not None else time.perf_counter() - t0
    search.explore_batch(model, maps, init, goal, ids, seeds, batch=100, t_max=100, k=10, spec_k=1, device=dev, sampler="device", timings=tmd)
    # the same loop on an arm environment (kuka7, the 48 real problems of tests/golden/arm_problems.npz): the edge check is K
    # forward-kinematics passes, so checking several candidate edges per iteration (spec_k) pays
    arm = _arm_problems()
    bx, bp = arm["kuka7_boxes"], arm["kuka7_box_ptr"]
    aprobs = [([(bx[j, :3], bx[j, 3:]) for j in range(bp[i], bp[i + 1])], arm["kuka7_start"][i], arm["kuka7_goal"][i]) for i in range(len(bp) - 1)]
    amodel = EncoderProcessDecoder(workspace_size=3, config_size=7, embed_size=64, obs_size=6).to(dev)
    amodel.load_state_dict(torch.load(os.path.join(G, "weights", "weights_kuka.pt"), map_location="cpu"))
    aseeds = [99 + rank * 1000 + i for i in range(len(aprobs))]
    arm_out = {}
    for spec_k in (1, 8):
        search.explore_batch_arm(amodel, 0, aprobs[:8], aseeds[:8], batch=100, t_max=200, k=10, spec_k=spec_k, device=dev)      # warm-up
        best = None
        for _rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ares = search.explore_batch_arm(amodel, 0, aprobs, aseeds, batch=100, t_max=200, k=10, spec_k=spec_k, device=dev)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0) if best is not None else time.perf_counter() - t0
        arm_out[spec_k] = (best, ares)
    if rank != 0:
        return None
    dt, res = out[1]
    # the reference's loop structure on the same kernels: one problem at a time, host arg-max, one edge check per iteration
    from gnn_motion_planning_b200.eval_gnn import explore
    n_host = 8
    env = _SyntheticMazeEnv(maps, init, goal, dev)
    t0 = time.perf_counter()
    same = 0
    for p in range(n_host):
        np.random.seed(seeds[p])
        env.init_new_problem(p)
        h = explore(env, model, None, smooth=True, batch=100, t_max=100, k=10, smoother="none")
        same += int(h["explored"] == res[p]["explored"] and h["c_explore"] == res[p]["c_explore"])
    dt_host = time.perf_counter() - t0
    return {
        "metric": "planner_problems_per_sec", "value": P * world / dt, "unit": "problems/s", "n_gpus": world, "wall_s": dt,
        "config": {"workload": "C1 batched: explore(batch=100, t_max=100, k=10, smoother='none') x %d maze problems/GPU (maze_maps_256.npz maps, "
                               "seeded init/goal), sampling + create_data + forward + lazy tree search all batched" % P},
        "success": sum(r["success"] for r in res), "problems": P, "all_ranks": summary,
        "mean_collision_checks": float(np.mean([r["c_explore"] for r in res])), "mean_explored": float(np.mean([len(r["explored"]) for r in res])),
        "spec_k8": {"value": P * world / out[8][0], "unit": "problems/s", "uncommitted_speculative_checks_per_problem": float(np.mean([r["spec_checks"] for r in out[8][1]])),
                    "identical_results": all(a["explored"] == b["explored"] and a["c_explore"] == b["c_explore"] for a, b in zip(res, out[8][1]))},
        "host_loop": {"value": n_host / dt_host, "unit": "problems/s", "sample": "first %d problems through eval_gnn.explore (reference loop, one edge per "
                      "iteration, same CUDA kernels)" % n_host, "identical_results": same == n_host},
        "phase_wall_ms": {k_: 1e3 * v_ for k_, v_ in tm.items()},
        "device_sampler": {"value": P / dt_dev, "unit": "problems/s (this rank)", "success": sum(r["success"] for r in res_dev),
                           "phase_wall_ms": {k_: 1e3 * v_ for k_, v_ in tmd.items()},
                           "note": "counter-based Philox sampler on the GPU (gmp_maze_sample_points): a new stream, same semantics"},
        "kuka7": {"problems": len(aprobs), "config": "explore(batch=100, t_max=200, k=10, smoother='none') on the 48 kuka7 problems of arm_problems.npz",
                  "value": len(aprobs) / arm_out[1][0], "unit": "problems/s (this rank, spec_k=1)", "success": sum(r["success"] for r in arm_out[1][1]),
                  "mean_collision_checks": float(np.mean([r["c_explore"] for r in arm_out[1][1]])),
                  "spec_k8": {"value": len(aprobs) / arm_out[8][0], "unit": "problems/s",
                              "uncommitted_speculative_checks_per_problem": float(np.mean([r["spec_checks"] for r in arm_out[8][1]])),
                              "identical_results": all(a["explored"] == b["explored"] and a["c_explore"] == b["c_explore"]
                                                       for a, b in zip(arm_out[1][1], arm_out[8][1]))}},
        "published_reference": {"value": 11.66, "unit": "problems/s", "source": "main.ipynb raw line 140 (author's machine, unknown hardware; different random mazes)"},
        "timing": "host wall clock around explore_batch (the loop has host code between rounds), best of 3, max over ranks",
        "gpu_launches": 12,
    }


class _SyntheticMazeEnv:
    """Minimal MazeEnv over in-memory arrays for the host-loop comparison of run_planner (same kernels as the package's MazeEnv)."""

    def __new__(cls, maps, init, goal, dev):
        import tempfile
        from gnn_motion_planning_b200.environment import MazeEnv
        with tempfile.NamedTemporaryFile(suffix=".npz", delete=False) as f:
            np.savez(f, maps=maps.astype(np.float64), init_states=init, goal_states=goal)
            path = f.name
        env = MazeEnv(dim=2, map_file=path, device=dev)
        os.unlink(path)
        return env


def run_single(args, wl, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from gnn_motion_planning_b200 import _lib, collision, graph, shard
    from gnn_motion_planning_b200.model import EncoderProcessDecoder
    dev = torch.device("cuda", local_rank)

    B, N, c = wl["batch"], wl["n"], wl["c"]
    # ---- synthetic inputs of this rank's shard (problems rank*B .. rank*B+B-1), host side
    vs, obss = zip(*[make_problem(wl, rank * B + g) for g in range(B)])
    v_h = torch.from_numpy(np.concatenate(vs)).pin_memory()
    goal_h = torch.from_numpy(np.stack([x[1] for x in vs])).pin_memory()
    obs_h = torch.from_numpy(np.concatenate(obss)).pin_memory()
    node_ptr = (np.arange(B + 1) * N).astype(np.int32)
    obs_ptr = np.cumsum([0] + [len(o) for o in obss]).astype(np.int32)
    n_free = np.full(B, N, np.int32)
    k1 = np.full(B, wl["k"], np.int32)
    maps_np = np.load(os.path.join(G, "maze_maps_256.npz"))["maps"]
    maps_h = torch.from_numpy(np.ascontiguousarray(maps_np)).pin_memory()
    prob_h = torch.from_numpy(((rank * B + np.arange(B)) % len(maps_np)).astype(np.int32)).pin_memory()
    is_maze = wl["env"] == "maze2"
    if not is_maze:
        ap = _arm_problems()
        boxes_d = torch.from_numpy(ap[wl["env"] + "_boxes"]).to(dev)
        box_ptr_d = torch.from_numpy(ap[wl["env"] + "_box_ptr"]).to(dev)
        prob_h = torch.from_numpy(((rank * B + np.arange(B)) % (len(ap[wl["env"] + "_box_ptr"]) - 1)).astype(np.int32)).pin_memory()

    model = EncoderProcessDecoder(workspace_size=wl["ws"], config_size=c, embed_size=wl["e"], obs_size=wl["s"]).to(dev)
    model.load_state_dict(torch.load(os.path.join(G, "weights", wl["weights"]), map_location="cpu"))
    model.eval()
    model.set_timing(True)
    if args.ef_mode != "auto":
        model.set_edge_feature_mode(args.ef_mode)

    # ---- the public batched API (gnn_motion_planning_b200.batch.HotPath) drives both measurements
    from gnn_motion_planning_b200.batch import HotPath
    v_d, goal_d, obs_d = v_h.to(dev), goal_h.to(dev), obs_h.to(dev)
    maps_d, prob_d = maps_h.to(dev), prob_h.to(dev)
    if is_maze:
        hp = HotPath(model, B, N, wl["k"], kind="maze", maps=maps_d, first_problem_id=rank * B, device=dev)
    else:
        hp = HotPath(model, B, N, wl["k"], kind="arm", boxes=boxes_d, box_ptr=box_ptr_d, arm_model=ARM_MODEL[wl["env"]], rrt_eps=0.5,
                     first_problem_id=rank * B, device=dev)
    gather_buf = torch.empty((world, B, 4), dtype=torch.float32, device=dev) if world > 1 else None
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    phase_ms = {}
    state = {}

    # ---- smoother forward (model_smoother.py:104-142), as model_smooth drives it: 5 x loop=1 on a path (smoother.py:233-246).
    # Synthetic path per problem: init (node 0), nodes 2..P-1, goal (node 1); samples = the problem's first 500 nodes as free and
    # the next 500 as collided (smoother.py:57-58 truncation); chain both ways + self loops (smoother.py:238-241).
    sm = None
    smooth_w = {"kuka7": ("smooth_7d_attv3.pt", 3)}.get(wl["env"])   # BASELINE configs[2] names the smoother; configs[1] is explorer only
    if smooth_w and N >= 1000:
        from gnn_motion_planning_b200.model_smoother import ModelSmoother
        sm = ModelSmoother(workspace_size=3, config_size=c, obs_size=6, embed_size=128).to(dev)   # (both sizes unused by the live forward)
        sm.load_state_dict(torch.load(os.path.join(G, "weights", smooth_w[0]), map_location="cpu"))
        sm.eval()
        P = 24
        idx = torch.tensor([0] + list(range(2, P)) + [1], device=dev)
        sm_path0 = v_d.view(B, N, c)[:, idx, :].reshape(B * P, c).contiguous()
        sm_samples = v_d.view(B, N, c)[:, :1000, :].reshape(B * 1000, c).contiguous()
        a_ = np.arange(P - 1)
        ei1 = np.concatenate([np.stack([a_, a_ + 1]), np.stack([a_ + 1, a_]), np.stack([np.arange(P), np.arange(P)])], 1)
        sm_ei = torch.from_numpy(np.tile(ei1, (1, B)).astype(np.int64)).to(dev)
        sm_args = (np.arange(B + 1) * P, np.arange(B + 1) * 1000, np.full(B, 500), np.arange(B + 1) * ei1.shape[1])

    def run_smoother():
        p_ = sm_path0
        for _ in range(5):
            p_ = sm.forward_batch(p_, sm_samples, sm_ei, sm_args[0], sm_args[1], sm_args[2], sm_args[3], loop=1)
        return p_

    def step(timed=False):
        evs = [ev() for _ in range(4)]
        bufs = hp.compute(v_d, goal_d, obs_d, obs_ptr, prob_d, events=evs)
        if world > 1:   # the only collective: per-problem result rows
            shard.gather_result_rows(bufs["rows"], out=gather_buf.view(world * B, 4), equal_shards=True)
        state.update(et=bufs["et"], edge_ptr=bufs["edge_ptr"], rows=bufs["rows"], checks=bufs["checks"], bufs=bufs)
        if sm is not None:
            se = (ev(), ev())
            se[0].record()
            state["smooth_path"] = run_smoother()
            se[1].record()
        if timed:
            torch.cuda.current_stream().synchronize()
            if sm is not None:
                phase_ms["smoother_forward_x5"] = phase_ms.get("smoother_forward_x5", 0.0) + se[0].elapsed_time(se[1])
            for k_, v_ in model.last_timings().items():
                phase_ms[k_] = phase_ms.get(k_, 0.0) + v_
            phase_ms["knn_graph"] = phase_ms.get("knn_graph", 0.0) + evs[0].elapsed_time(evs[1])
            phase_ms["explorer_forward"] = phase_ms.get("explorer_forward", 0.0) + evs[1].elapsed_time(evs[2])
            phase_ms["collision"] = phase_ms.get("collision", 0.0) + evs[2].elapsed_time(evs[3])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps, finish=None):
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            fn()
        if finish:
            finish()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # ---- kernel-only: inputs resident in HBM
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # the timed region runs the bare steps (a step's only host synchronisation is its own edge_ptr read-back); the per-phase
    # breakdown -- a stream synchronise and event reads after every step -- is collected in a second pass of the same length
    ms_step = timed_region(lambda: step(timed=False), args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(args.steps):
        step(timed=True)
    et = state["et"]
    checks_total = int(state["checks"][:et].sum())

    # ---- end to end through the public API: pinned host buffers in, pinned host buffers out, every step's
    # host->device and device->host copies inside the timed region.  Results of step k are awaited (HotPath.wait) while
    # step k+1 is already enqueued: the read-back runs on a second stream and overlaps the next step's kernels.
    io = {}
    pending = []

    def e2e_step():
        t = hp.submit(v_h, goal_h, obs_h, obs_ptr, prob_h, maps_h=maps_h if is_maze else None)
        if sm is not None:
            state["smooth_path"] = run_smoother()
        if world > 1:
            shard.gather_result_rows(t["rows"], out=gather_buf.view(world * B, 4), equal_shards=True)
        pending.append(t)
        if len(pending) > 1:
            res = HotPath.wait(pending.pop(0))          # the host owns step k-1's results from here on
            io["last_logit"] = float(res["logits"][0])
        io["h2d"], io["d2h"] = t["h2d_bytes"], t["d2h_bytes"]

    def e2e_finish():
        while pending:
            res = HotPath.wait(pending.pop(0))
            io["last_logit"] = float(res["logits"][0])
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    e2e_finish()
    ms_e2e = timed_region(e2e_step, args.steps, finish=e2e_finish)

    # ---- the drop-in output: dense [N,N] per graph (model.py:148-150) instead of the sparse [E] logits the batched path returns
    dense_info = None
    if args.workload == "C2":
        n_dense = B * N * N
        dense_buf = torch.empty(n_dense, dtype=torch.float32, device=dev)
        ei_d, ep_h = state["bufs"]["ei"], state["bufs"]["edge_ptr"]
        t_ = {}
        for dense in (False, True):
            for it in range(4):
                a_, b_ = ev(), ev()
                a_.record()
                model.forward_batch(v_d, ei_d, goal_d, obs_d, node_ptr, ep_h, obs_ptr, loop=5, dense=dense, dense_out=dense_buf if dense else None)
                b_.record()
                b_.synchronize()
                t_[dense] = a_.elapsed_time(b_)
        dense_info = {"forward_ms_sparse": t_[False], "forward_ms_dense": t_[True], "dense_bytes_per_step": n_dense * 4,
                      "note": "timed separately: the batched path (value, e2e) returns sparse [E] logits; the reference-shaped dense [N,N] "
                              "matrices add a memset + scatter of %.2f GB per step on the device and would be %.1fx the current read-back"
                              % (n_dense * 4 / 1e9, n_dense * 4 / max(io["d2h"], 1))}
        del dense_buf

    if rank != 0:
        return None

    K = args.steps
    graphs_per_s = B * world / (ms_step / 1000.0)
    o_mean = float(np.mean(np.diff(obs_ptr)))
    e_mean = et / B
    ef_ms = phase_ms["edge_feature"] / K
    ef_flops = sum(edge_feature_flops(int(ne), int(no), c, wl["e"]) for ne, no in zip(np.diff(state["edge_ptr"]), np.diff(obs_ptr)))
    fwd_flops = sum(ref_oplist_flops(N, int(ne), int(no), c, wl["e"], wl["s"]) for ne, no in zip(np.diff(state["edge_ptr"]), np.diff(obs_ptr)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = {}
    try:   # DRAM bytes per launch of the two named kernels, from the committed ncu --set full captures (C2 workload only)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))) if args.workload == "C2" else {}
    except Exception:
        pass

    def traffic_of(kernel):
        t = traffic.get(kernel)
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    hbm = peaks.get("hbm_gbs", 6650.0)
    ef_tflops = ef_flops / (ef_ms * 1e-3) / 1e12
    sm_mhz = (clocks or {}).get("sm_mhz") or 1900.0
    fp32_peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    msg_ms = phase_ms["edge_msg"] / K / 5.0
    # algorithmic HBM bytes of one message round: the loop-invariant edge term P (e*4 B/edge, streamed by TMA bulk copies)
    # + the CSR (src, dst) ids (8 B/edge).  The gathered A[src] / B[dst] rows (2*e*4 B/edge) are L2-resident (2 x 33 MB
    # per 256-graph batch) and are NOT counted.
    msg_bytes = et * (wl["e"] * 4 + 8)
    if wl["e"] == 32:
        # tcgen05 path: 3xTF32 (three kind::tf32 MMAs per product, fp32 accumulate in TMEM).  TF32 dense runs at half the bf16
        # rate and every product is issued three times, so the ceiling of this arithmetic is peak/6 of the bf16 figure.
        halves = 1 if (args.ef_mode == "tc4" or (args.ef_mode == "auto" and 2 * c > 8)) else 2     # explorer.cu: auto mode
        # explorer.cu: one MMA issuer warp per tile (third template argument) for the eight-warp organisation when every graph
        # has 1..128 obstacles (true of every bench workload) unless GMP_TC_RD=0 / an explicit lockstep mode
        rd = halves == 2 and args.ef_mode in ("auto", "tcrd") and os.environ.get("GMP_TC_RD", "1")[0] != "0"
        ef_kernel = "edge_feature_tc_kernel<%d,%d%s>" % (c, halves, ",true" if rd else "")
        ef_note = ("tcgen05.mma kind::tf32, 3xTF32 split operands (1e-4 logit tolerance rules out 1-pass TF32/BF16), A operands and "
                   "accumulators in TMEM; second template argument = epilogue warp groups per 128-edge tile (2: eight warps, columns split), third = one MMA issuer warp per tile; against the 3xTF32 ceiling (bf16 peak / 6 = %.0f TFLOP/s) the fraction is %.3f; the fp32 SIMT "
                   "kernel it replaces peaked at 148 SM x 128 FMA x %.0f MHz = %.1f TFLOP/s"
                   % (peak_tf / 6.0, ef_tflops / (peak_tf / 6.0), sm_mhz, fp32_peak_tf))
    else:
        # embed 64: the phase-split tcgen05 stage (encoder, Block x3, tail = 5 launches; graphs here have <= 32 obstacles)
        ef_kernel = "edge_feature64_tc_kernel<%d,phase 0|1|1|1|2>" % c
        ef_note = ("five launches of the phase-split tcgen05 3xTF32 stage, timed together (ms_per_launch = the whole stage); against "
                   "the 3xTF32 ceiling (bf16 peak / 6 = %.0f TFLOP/s) the fraction is %.3f" % (peak_tf / 6.0, ef_tflops / (peak_tf / 6.0)))
    line = {
        "metric": "explorer_graphs_per_sec", "value": graphs_per_s, "unit": "graphs/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl, world),
        "edges_per_graph": e_mean, "obstacles_per_graph": o_mean,
        "phases_ms_per_step": {k_: v_ / K for k_, v_ in sorted(phase_ms.items())},
        "explorer_forward_graphs_per_sec": B * world / (phase_ms["explorer_forward"] / K / 1e3),
        "collision_checks_per_sec": et * world / (phase_ms["collision"] / K / 1e3),
        "collision_state_checks_per_edge": checks_total / et,
        "collision_free_edge_fraction": float(state["rows"][:, 2].sum()) / et,
        "knn_graphs_per_sec": B * world / (phase_ms["knn_graph"] / K / 1e3),
        "forward_tflops_ref_oplist": fwd_flops / (phase_ms["explorer_forward"] / K / 1e3) / 1e12,
        "roofline": {"kernel": ef_kernel, "bound": "tensor", "achieved": ef_tflops, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": ef_tflops / peak_tf, "traffic": traffic_of(ef_kernel),
                     "peak_source": peak_src,
                     "ms_per_launch": ef_ms, "algorithmic_gflop_per_launch": ef_flops / 1e9,
                     "timing": "CUDA events on the launching stream around the phase, averaged over %d steps run right after the timed "
                               "region (same inputs; reading the events needs a stream synchronise per step, which is kept out of the "
                               "timed region)" % K,
                     "note": ef_note},
        "roofline_hbm_kernel": {"kernel": "edge_msg_tc_kernel<%d>" % wl["e"], "bound": "hbm", "achieved": msg_bytes / (msg_ms * 1e-3) / 1e9,
                                "peak": hbm, "unit": "GB/s", "frac": msg_bytes / (msg_ms * 1e-3) / 1e9 / hbm,
                                "traffic": traffic_of("edge_msg_tc_kernel<%d>" % wl["e"]), "algorithmic_bytes_per_launch": msg_bytes,
                                "ms_per_launch": msg_ms},
        "e2e": {"value": B * world / (ms_e2e / 1000.0), "unit": "graphs/s", "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                "ms_per_step": ms_e2e, "api": "gnn_motion_planning_b200.batch.HotPath.submit/wait (double-buffered; read-back of "
                                              "step k overlaps the kernels of step k+1)"},
        "gpu_launches": K * (5 + 3 + 1 + 1 + 1 + (3 if wl["e"] == 32 else 7) + 6 + 5 + 1 + 1 + 1 + (20 if sm is not None else 0)),
        "gpu_launches_note": "per step: knn 5 (select,row_count,row_scan,graph_scan,emit) + csr 3 + goal_index + obstacle + node_pre + "
                             "edge_feature (e=32: obs_table_tc + unit_meta + edge_feature_tc; e=64: obs_table_tc64 + unit_meta + 5 phases) + node_loop x6 + edge_msg x5 + policy + "
                             "{maze,arm}_edge_graph (arms: node_flags + edge_graph_cached) + result_rows + smoother 5 x (graph,node,msg,path); memsets/copies not counted",
        "clocks": clocks,
    }
    if dense_info:
        line["dense_output"] = dense_info
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = args.cpu_sample if args.workload == "C2" else (4 if args.workload == "C3" else 2)
        gps, dt, _ = cpu_graphs_per_sec(wl, list(range(n_s)), threads)
        line["cpu_baseline"] = {"value": gps, "unit": "graphs/s", "cores": threads, "kind": "port",
                                "sample": "first %d graphs of the workload (oracle: knn graph + explorer forward + edge checks), %.1f s" % (n_s, dt)}
    return line


if __name__ == "__main__":
    main()
