"""Batched hot path: many independent planning problems per call (the capability the reference lacks).

``HotPath`` strings the C-ABI entry points together for a packed batch of problems that share a size class:

    k-NN RGG construction (``gmp_knn_graph``)            <- eval_gnn.create_data      eval_gnn.py:150-165
    explorer forward      (``gmp_explorer_forward``)     <- model(**data, loop=5)     eval_gnn.py:194
    collision check of EVERY edge (``gmp_maze_edge_fp_graph`` / ``gmp_arm_edge_fp_graph``)
                                                         <- env._edge_fp, one edge at a time, eval_gnn.py:215
    per-problem result rows (``gmp_result_rows``)        <- the `solutions` tuples    eval_gnn.py:120-134

``compute`` runs one batch from device-resident inputs.  ``submit`` / ``wait`` are the end-to-end path: inputs come
from pinned host buffers, results (edge_index, logits, free flags, edge_ptr) are delivered into pinned host buffers;
two device/host buffer sets and three streams (graph build | forward + collision | read-back) keep the GPU busy
across batches.  The edge list travels back as LOCAL ids in int16 (int32 when a graph has more than 32 767 nodes):
4 B/edge instead of the 16 B/edge of the int64 ``[2,E]`` PyG layout, which was 76 % of the read-back (the int64 tensor
stays on the device for the kernels; ``result["edge_index"].to(torch.int64)`` restores the reference dtype on the host).
"""
import numpy as np
import torch

from . import _lib, collision, graph


class HotPath:
    def __init__(self, model, n_problems, n_nodes, k, kind="maze", maps=None, boxes=None, box_ptr=None, arm_model=None,
                 rrt_eps=0.5, loop=5, first_problem_id=0, device=None):
        self.model, self.B, self.N, self.k, self.kind, self.loop = model, int(n_problems), int(n_nodes), int(k), kind, loop
        self.dev = torch.device(device if device is not None else "cuda")
        self.c = model.config_size
        self.first_problem_id = first_problem_id
        self.node_ptr = (np.arange(self.B + 1) * self.N).astype(np.int32)
        self.node_ptr_d = torch.from_numpy(self.node_ptr).to(self.dev)
        self.n_free = np.full(self.B, self.N, np.int32)
        self.k1 = np.full(self.B, self.k, np.int32)
        if maps is not None and maps.dtype != torch.uint8:      # reference map files are float64 {0,1}
            maps = (maps != 0).to(torch.uint8).contiguous()
        self.maps, self.boxes, self.box_ptr, self.arm_model, self.rrt_eps = maps, boxes, box_ptr, arm_model, rrt_eps
        self.cap = int(self.B * _lib.load().gmp_knn_graph_max_edges(self.N, self.k))
        self.id_dtype = torch.int16 if self.N <= 32767 else torch.int32
        self.sets = [self._alloc_set() for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.graph_stream = torch.cuda.Stream(device=self.dev)
        self._k = 0

    def _alloc_set(self):
        d = self.dev
        s = dict(ei=torch.empty((2, self.cap), dtype=torch.int64, device=d), logits=torch.empty(self.cap, dtype=torch.float32, device=d),
                 free=torch.empty(self.cap, dtype=torch.uint8, device=d), checks=torch.empty(self.cap, dtype=torch.int32, device=d),
                 rows=torch.empty((self.B, 4), dtype=torch.float32, device=d),
                 ei_narrow=torch.empty((2, self.cap), dtype=self.id_dtype, device=d), pending=None,
                 compute_done=torch.cuda.Event(), copy_done=torch.cuda.Event(), et=0, edge_ptr=None, host=None, inputs=None)
        return s

    # ------------------------------------------------------------------ one batch, device-resident inputs
    def _build_graphs(self, v, bufs):
        if bufs.get("edge_ptr_pin") is None:
            bufs["edge_ptr_pin"] = torch.empty(self.B + 1, dtype=torch.int32).pin_memory()
        ei, edge_ptr = graph.knn_graph_batch(v, self.node_ptr, self.n_free, self.k1, edge_index_out=bufs["ei"],
                                             edge_ptr_host=bufs["edge_ptr_pin"])  # syncs: B+1 ints, posted by a kernel
        bufs["et"], bufs["edge_ptr"] = int(edge_ptr[-1]), edge_ptr
        return ei

    def _score_and_check(self, v, goal, obstacles, obs_ptr, problem_of_graph, bufs, maps=None, events=None):
        ei, edge_ptr, et = bufs["ei"], bufs["edge_ptr"], bufs["et"]
        logits = self.model.forward_batch(v, ei, goal, obstacles, self.node_ptr, edge_ptr, obs_ptr, loop=self.loop, dense=False,
                                          out=bufs["logits"])
        if events:
            events[2].record()
        edge_ptr_d = torch.from_numpy(edge_ptr).to(self.dev, non_blocking=True)
        if self.kind == "maze":
            collision.maze_edge_fp_graph(v, ei, self.node_ptr_d, edge_ptr_d, maps if maps is not None else self.maps, et,
                                         problem_of_graph=problem_of_graph, want_checks=True, free_out=bufs["free"],
                                         checks_out=bufs["checks"])
        else:
            collision.arm_edge_fp_graph(self.arm_model, v, ei, self.node_ptr_d, edge_ptr_d, self.boxes, self.box_ptr, et,
                                        rrt_eps=self.rrt_eps, problem_of_graph=problem_of_graph, want_checks=True,
                                        free_out=bufs["free"], checks_out=bufs["checks"])
        if events:
            events[3].record()
        collision.result_rows(logits, bufs["free"], edge_ptr_d, self.first_problem_id, out=bufs["rows"])

    def compute(self, v, goal, obstacles, obs_ptr, problem_of_graph, bufs=None, events=None):
        """v [B*N,c] f32, goal [B,c], obstacles [O_total,s], obs_ptr host [B+1], problem_of_graph [B] i32 (device).
        Fills bufs (edge_index, logits, free, checks, rows) on the current stream; returns bufs.  `events`: optional 4 CUDA
        events recorded at the phase boundaries (graph build | forward | collision)."""
        bufs = bufs if bufs is not None else self.sets[0]
        if events:
            events[0].record()
        self._build_graphs(v, bufs)
        if events:
            events[1].record()
        self._score_and_check(v, goal, obstacles, obs_ptr, problem_of_graph, bufs, events=events)
        return bufs

    # ------------------------------------------------------------------ end to end: pinned host in, pinned host out
    def _host_set(self):
        pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype).pin_memory()  # noqa: E731
        return dict(ei=pin((2, self.cap), dtype=self.id_dtype), logits=pin(self.cap, dtype=torch.float32),
                    free=pin(self.cap, dtype=torch.uint8), rows=pin((self.B, 4), dtype=torch.float32))

    def submit(self, v_h, goal_h, obs_h, obs_ptr, prob_h, maps_h=None):
        """Enqueue one batch from PINNED host tensors; returns a ticket for ``wait``.  Three streams form a pipeline over
        two buffer sets: host->device copies + graph build of batch k+1 (graph stream) run while the forward / collision
        kernels of batch k occupy the main stream, and the device->host read-back of batch k (copy stream) overlaps both.
        The only host synchronisation is the B+1-int edge_ptr read-back of the graph build, which waits for the graph
        stream only -- the main stream never drains."""
        s = self.sets[self._k % 2]
        if s["pending"] is not None:
            raise RuntimeError("HotPath.submit: at most two batches may be in flight -- wait() on the ticket of batch %d first "
                               "(its host buffers would be overwritten)" % s["pending"])
        s["pending"] = self._k
        self._k += 1
        if s["host"] is None:
            s["host"] = self._host_set()
            s["inputs"] = [None, None, None, None]
            s["graph_done"] = torch.cuda.Event()
            s["maps_d"] = torch.empty_like(self.maps) if (self.kind == "maze" and maps_h is not None) else None
        main = torch.cuda.current_stream(self.dev)
        gs = self.graph_stream
        gs.wait_event(s["copy_done"])                         # the previous read-back of this buffer set has finished
        with torch.cuda.stream(gs):
            staged = []
            for i, h in enumerate((v_h, goal_h, obs_h, prob_h)):   # device staging buffers grow on demand (ragged obstacle counts)
                d = s["inputs"][i]
                if d is None or d.shape[0] < h.shape[0] or d.shape[1:] != h.shape[1:] or d.dtype != h.dtype:
                    d = s["inputs"][i] = torch.empty((max(h.shape[0], 1),) + tuple(h.shape[1:]), dtype=h.dtype, device=self.dev)
                d = d[:h.shape[0]]
                d.copy_(h, non_blocking=True)
                staged.append(d)
            if s["maps_d"] is not None:                       # the occupancy maps are per-problem inputs too
                s["maps_d"].copy_(maps_h, non_blocking=True)
            self._build_graphs(staged[0], s)
            s["graph_done"].record(gs)
        main.wait_event(s["graph_done"])
        self._score_and_check(staged[0], staged[1], staged[2], obs_ptr, staged[3], s, maps=s["maps_d"])
        n = s["et"]
        _lib.check(_lib.load().gmp_edge_index_narrow(_lib.ptr(s["ei"]), s["ei"].stride(0), n, 16 if self.id_dtype == torch.int16 else 32,
                                                     _lib.ptr(s["ei_narrow"]), s["ei_narrow"].stride(0), _lib.stream_ptr(self.dev)))
        s["compute_done"].record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["compute_done"])
            s["host"]["logits"][:n].copy_(s["logits"][:n], non_blocking=True)
            s["host"]["free"][:n].copy_(s["free"][:n], non_blocking=True)
            s["host"]["ei"][0, :n].copy_(s["ei_narrow"][0, :n], non_blocking=True)
            s["host"]["ei"][1, :n].copy_(s["ei_narrow"][1, :n], non_blocking=True)
            s["host"]["rows"].copy_(s["rows"], non_blocking=True)
            s["copy_done"].record(self.copy_stream)
        s["h2d_bytes"] = sum(t.numel() * t.element_size() for t in (v_h, goal_h, obs_h, prob_h) + ((maps_h,) if maps_h is not None else ()))
        s["d2h_bytes"] = n * (4 + 1 + 2 * s["host"]["ei"].element_size()) + self.B * 16 + (self.B + 1) * 4
        # the ticket snapshots what belongs to THIS batch; the buffer set itself is reused two submits later
        return dict(set=s, seq=s["pending"], et=n, edge_ptr=s["edge_ptr"], rows=s["rows"], h2d_bytes=s["h2d_bytes"], d2h_bytes=s["d2h_bytes"])

    @staticmethod
    def wait(ticket):
        """Block until the results of a submitted batch are in its pinned host buffers; returns them."""
        s = ticket["set"]
        if s["pending"] != ticket["seq"]:
            raise RuntimeError("HotPath.wait: this ticket's buffers were already waited on and reused")
        s["copy_done"].synchronize()
        s["pending"] = None
        n = ticket["et"]
        h = s["host"]
        return dict(edge_ptr=ticket["edge_ptr"], edge_index=h["ei"][:, :n], logits=h["logits"][:n], free=h["free"][:n], rows=h["rows"])
