// Batched lazy tree search on the device -- generic over the environment (maze.cu and arm.cu instantiate it).
#pragma once
#include "common.cuh"

namespace gmp {

// =====================================================================================================================
// Batched lazy tree search on the device: the inner loop of explore()                        (eval_gnn.py:198-233)
// Env supplies: struct Args (launch-constant, by value), struct Ctx (per problem), make(Args, graph) -> Ctx,
//   edge(Ctx, pa, pb, cnt)  = env._edge_fp(v[a], v[b]) with its collision_check_count increments,
//   goal(Ctx, pb, goal, cnt) = env.in_goal_region(v[b]).
// =====================================================================================================================
// The reference masks a dense [N,N] policy on the host and then, one edge per Python iteration, takes the arg-max over the rows
// of the explored nodes (three torch.where passes over a len(explored) x N slice), collision-checks that edge, and either grows
// the tree or zeroes the edge.  Here one CTA owns one problem and does the same thing on the SPARSE logits:
//   * policy[a, b] = logit of edge (b -> a) (model.py:148-150), so row a is the in-edge list of a: a CSR by target is built in
//     the prologue (counting sort in shared / global scratch), with the masks of eval_gnn.py:198-202 applied while filling --
//     diagonal, collided rows and columns, and the explored-edge list INCLUDING the reference's reshape(2,-1) quirk (pairs
//     (L[i], L[M+i]) of the flattened list are zeroed, not the recorded pairs);  columns of explored nodes are
//     masked through a bitmap instead of being written;
//   * every explored row caches its best remaining entry; an iteration is a block-wide arg-max over those (value, then position
//     in the explored list, then column: the order torch.where + argmax gives), one edge check, and the recomputation of the
//     rows whose cached entry died (the row of the checked edge, or every row that pointed at the newly explored node);
//   * SPECULATION (spec_k > 1): besides the winner, the next spec_k - 1 best cached entries are collision-checked in the same
//     iteration by otherwise idle threads and their outcome is remembered per edge.  A later iteration that selects such an edge
//     commits the remembered outcome instead of checking again.  Commits happen strictly in the reference's order, so success,
//     path, explored order and collision_check_count are IDENTICAL for every spec_k; checks that were never committed are
//     counted separately (n_spec_checks).
// A value of exactly 0.0 means "no edge", as in the reference (eval_gnn.py:204-210).  The reference's loop condition is
// `policy[explored, :].sum() != 0`; this kernel stops when no non-zero entry is left (the two differ only if non-zero logits
// cancel to exactly 0.0 in the float sum).
constexpr int kSearchThreads = 128;
constexpr int kSearchMaxNodes = 16384;     // explored bitmap in shared memory

struct SearchArgs {
  // this round's packed batch (v: [N_total, dim] f32, goal: [S, dim] f64 = env.goal_state per slot)
  const float* v; int dim; const int32_t* node_ptr; const int32_t* n_free; const int64_t* edge_index; int64_t row_stride; const int32_t* edge_ptr;
  const float* logits; const double* goal; const int32_t* slot_of_graph;
  int spec_k, first_round;
  // scratch (per call)
  int32_t* in_ptr; int32_t* cursor; float* bestv; int32_t* bests; int32_t* csr_src; float* csr_val; int32_t* chk;
  // persistent search state, one row per slot
  int32_t* explored; int32_t* n_explored; int32_t* prev; int32_t* elist; int32_t* n_elist; int32_t* n_checks; int32_t* n_spec;
  int32_t* status; int32_t* path; int32_t* path_len; float* path_cost;
  int cap_nodes, cap_elist;
};

__device__ __forceinline__ bool better(float v1, int p1, int c1, float v2, int p2, int c2) {
  // arg-max order of the reference: larger value first; ties -> the earlier (row position in `explored`, column)
  if (v1 != v2) return v1 > v2;
  if (p1 != p2) return p1 < p2;
  return c1 < c2;
}

template <class Env>
__global__ void __launch_bounds__(kSearchThreads) tree_search_kernel(SearchArgs A, typename Env::Args EA) {
  __shared__ uint32_t s_expl[kSearchMaxNodes / 32];
  __shared__ float s_v[kSearchThreads];
  __shared__ int s_p[kSearchThreads], s_s[kSearchThreads];
  __shared__ int s_win[3];             // winner: position, slot, (unused)
  __shared__ int s_spec[32];           // slots to check speculatively
  __shared__ int s_ctl[4];             // 0: continue flag, 1: recompute mode (0 row only, 1 column b died), 2: a (row), 3: b
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot_id = A.slot_of_graph ? A.slot_of_graph[g] : g;
  if (!A.first_round && A.status[slot_id] != 2) return;        // only problems waiting for a new graph continue
  const int n0 = A.node_ptr[g], N = A.node_ptr[g + 1] - n0, F = A.n_free[g];
  const int e0 = A.edge_ptr[g], E = A.edge_ptr[g + 1] - e0;
  const int dim = A.dim;
  const float* vg = A.v + (size_t)n0 * dim;
  const typename Env::Ctx ctx = Env::make(EA, g);
  int32_t* in_ptr = A.in_ptr + n0 + g;             // N + 1 entries per graph
  int32_t* cursor = A.cursor + n0;
  float* bestv = A.bestv + n0;                     // indexed by position in the explored list
  int32_t* bests = A.bests + n0;
  int32_t* csr_src = A.csr_src + e0;
  float* csr_val = A.csr_val + e0;
  int32_t* chk = A.chk + e0;                       // remembered edge checks: (count << 2) | 1 free / 2 blocked; 0 unknown
  int32_t* explored = A.explored + (int64_t)slot_id * A.cap_nodes;
  int32_t* prev = A.prev + (int64_t)slot_id * A.cap_nodes;
  int32_t* elist = A.elist + (int64_t)slot_id * A.cap_elist;
  if (N > kSearchMaxNodes || N > A.cap_nodes) { if (tid == 0) A.status[slot_id] = 3; return; }

  // ---- prologue: CSR by target with the masks of eval_gnn.py:198-202
  for (int i = tid; i < N; i += kSearchThreads) cursor[i] = 0;
  for (int i = tid; i < (N + 31) / 32; i += kSearchThreads) s_expl[i] = 0;
  __syncthreads();
  for (int e = tid; e < E; e += kSearchThreads) atomicAdd(cursor + (int)A.edge_index[A.row_stride + e0 + e], 1);
  __syncthreads();
  if (warp == 0) {                                 // exclusive scan of the in-degrees, 32 nodes per step
    int run = 0;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const int c = i < N ? cursor[i] : 0;
      int x = c;
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (i < N) in_ptr[i] = run + x - c;
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) in_ptr[N] = run;
  }
  __syncthreads();
  for (int i = tid; i < N; i += kSearchThreads) cursor[i] = 0;
  __syncthreads();
  for (int e = tid; e < E; e += kSearchThreads) {
    const int src = (int)A.edge_index[e0 + e], dst = (int)A.edge_index[A.row_stride + e0 + e];
    const int sl = in_ptr[dst] + atomicAdd(cursor + dst, 1);
    float val = A.logits[e0 + e];
    if (src == dst || src >= F || dst >= F) val = 0.f;          // :198, :200-201 (collided nodes are the rows / columns >= F)
    csr_src[sl] = src;
    csr_val[sl] = val;
    chk[sl] = 0;
  }
  if (tid == 0 && A.first_round) {                 // explored = [0]; explored_edges = [[0, 0]]   (eval_gnn.py:183-186)
    explored[0] = 0; A.n_explored[slot_id] = 1; prev[0] = 0;
    elist[0] = 0; elist[1] = 0; A.n_elist[slot_id] = 2;
    A.n_checks[slot_id] = 0; A.n_spec[slot_id] = 0; A.path_len[slot_id] = 0;
    if (A.path_cost) A.path_cost[slot_id] = 0.f;
  }
  __syncthreads();
  int n_expl = A.n_explored[slot_id];
  int n_el = A.n_elist[slot_id];
  for (int i = tid; i < n_expl; i += kSearchThreads) atomicOr(&s_expl[explored[i] >> 5], 1u << (explored[i] & 31));
  {                                                // :202 as the author's torch evaluated it: rows L[:M], columns L[M:] of the FLAT list
    const int M = n_el / 2;
    for (int i = tid; i < M; i += kSearchThreads) {
      const int r = elist[i], c = elist[M + i];
      if (r < N)
        for (int sl = in_ptr[r]; sl < in_ptr[r + 1]; ++sl)
          if (csr_src[sl] == c) csr_val[sl] = 0.f;
    }
  }
  __syncthreads();
  auto is_expl = [&](int node) { return (s_expl[node >> 5] >> (node & 31)) & 1u; };
  // best remaining entry of the row of explored[pos] (one warp)
  auto row_best = [&](int pos) {
    const int a = explored[pos];
    float bv = 0.f; int bs = -1, bc = 0x7fffffff;
    for (int sl = in_ptr[a] + lane; sl < in_ptr[a + 1]; sl += 32) {
      const float val = csr_val[sl];
      const int c = csr_src[sl];
      if (val != 0.f && !is_expl(c) && (bs < 0 || better(val, 0, c, bv, 0, bc))) { bv = val; bs = sl; bc = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (os >= 0 && (bs < 0 || better(ov, 0, oc, bv, 0, bc))) { bv = ov; bs = os; bc = oc; }
    }
    if (lane == 0) { bestv[pos] = bv; bests[pos] = bs; }
  };
  for (int pos = warp; pos < n_expl; pos += kSearchThreads / 32) row_best(pos);
  __syncthreads();

  int n_chk = A.n_checks[slot_id], n_spec = A.n_spec[slot_id];
  int result = 2;                                  // 1 success, 2 exhausted (needs a new graph), 3 capacity
  while (true) {
    // ---- block-wide arg-max over the cached row bests
    float mv = 0.f; int mp = -1, ms = -1, mc = 0;
    for (int pos = tid; pos < n_expl; pos += kSearchThreads) {
      const int sl = bests[pos];
      if (sl < 0) continue;
      const float val = bestv[pos];
      const int c = csr_src[sl];
      if (mp < 0 || better(val, pos, c, mv, mp, mc)) { mv = val; mp = pos; ms = sl; mc = c; }
    }
    s_v[tid] = mv; s_p[tid] = mp; s_s[tid] = ms;
    __syncthreads();
    if (warp == 0) {
      // winner, then the next spec_k - 1 thread-local bests (the speculative candidates)
      const int rounds = min(A.spec_k, 32);
      for (int r = 0; r < rounds; ++r) {
        float bv = 0.f; int bp = -1, bs = -1, bt = -1;
        for (int t = lane; t < kSearchThreads; t += 32) {
          const int p = s_p[t];
          if (p < 0) continue;
          const float val = s_v[t];
          if (bp < 0 || better(val, p, 0, bv, bp, 0)) { bv = val; bp = p; bs = s_s[t]; bt = t; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int op = __shfl_xor_sync(0xffffffffu, bp, o), os = __shfl_xor_sync(0xffffffffu, bs, o), ot = __shfl_xor_sync(0xffffffffu, bt, o);
          if (op >= 0 && (bp < 0 || better(ov, op, 0, bv, bp, 0))) { bv = ov; bp = op; bs = os; bt = ot; }
        }
        if (lane == 0) {
          if (r == 0) { s_win[0] = bp; s_win[1] = bs; }
          else s_spec[r - 1] = bp >= 0 ? bs : -1;
          if (bt >= 0) s_p[bt] = -1;             // taken
        }
        __syncwarp();
      }
    }
    __syncthreads();
    const int wpos = s_win[0], wslot = s_win[1];
    if (wpos < 0) { result = 2; break; }
    // ---- edge checks: thread 0 the winner (unless remembered), threads 1.. the speculative candidates
    if (tid < min(A.spec_k, 32)) {
      const int sl = tid == 0 ? wslot : s_spec[tid - 1];
      if (sl >= 0 && chk[sl] == 0) {
        int row = 0;                               // the row (target node) of this slot: binary search in in_ptr
        { int lo = 0, hi = N; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (in_ptr[mid] <= sl) lo = mid; else hi = mid; } row = lo; }
        int cnt = 0;
        const bool ok = Env::edge(ctx, vg + (size_t)row * dim, vg + (size_t)csr_src[sl] * dim, cnt);   // env._edge_fp(v[a], v[b])  (eval_gnn.py:215)
        chk[sl] = (cnt << 2) | (ok ? 1 : 2);
      }
    }
    __syncthreads();
    // ---- commit (thread 0), in the reference's order
    if (tid == 0) {
      const int a = explored[wpos], b = csr_src[wslot];
      const int rec = chk[wslot];
      const bool ok = (rec & 3) == 1;
      n_chk += rec >> 2;
      chk[wslot] = rec | 0x40000000;               // committed (not a wasted speculation)
      if (n_el + 4 > A.cap_elist || n_expl + 1 > A.cap_nodes) { s_ctl[0] = 3; }
      else {
        elist[n_el] = a; elist[n_el + 1] = b; elist[n_el + 2] = b; elist[n_el + 3] = a;     // explored_edges.extend (:214)
        n_el += 4;
        if (ok) {
          explored[n_expl] = b; prev[b] = a;                                              // :216-218
          s_expl[b >> 5] |= 1u << (b & 31);                                               // policy[:, end_b] = 0  (:220)
          s_ctl[1] = 1; s_ctl[2] = a; s_ctl[3] = b;
          // in_goal_region(v[b]): distance in float64 against env.goal_state, then one more state check   (:221)
          int gc = 0;
          const bool goal_hit = Env::goal(ctx, vg + (size_t)b * dim, A.goal + (size_t)slot_id * dim, gc);
          n_chk += gc;
          s_ctl[0] = goal_hit ? 1 : 2;
        } else {
          csr_val[wslot] = 0.f;                                                           // policy[end_a, end_b] = 0   (:232)
          for (int sl = in_ptr[b]; sl < in_ptr[b + 1]; ++sl)
            if (csr_src[sl] == a) csr_val[sl] = 0.f;                                      // policy[end_b, end_a] = 0   (:233)
          s_ctl[1] = 0; s_ctl[2] = wpos; s_ctl[3] = b;
          s_ctl[0] = 2;
        }
      }
    }
    __syncthreads();
    const int ctl = s_ctl[0];
    if (ctl == 3) { result = 3; break; }
    if (s_ctl[1] == 1) {                           // a node joined the tree: its row, and every row whose best pointed at it
      const int b = s_ctl[3];
      n_expl += 1;
      if (ctl == 1) { result = 1; break; }
      for (int pos = warp; pos < n_expl; pos += kSearchThreads / 32) {
        const int sl = bests[pos];
        if (pos == n_expl - 1 || (sl >= 0 && csr_src[sl] == b)) row_best(pos);
      }
    } else if (warp == 0) row_best(s_ctl[2]);
    __syncthreads();
  }
  // ---- write back
  __syncthreads();
  if (tid == 0) {
    A.n_explored[slot_id] = n_expl; A.n_elist[slot_id] = n_el; A.n_checks[slot_id] = n_chk;
    A.status[slot_id] = result;
    if (result == 1) {                             // path = [b, prev[b], ..., 0] reversed   (:223-229)
      int len = 0;
      for (int node = explored[n_expl - 1];; node = prev[node]) { ++len; if (node == 0) break; }
      int* out = A.path + (int64_t)slot_id * A.cap_nodes;
      int i = len - 1;
      for (int node = explored[n_expl - 1];; node = prev[node]) { out[i--] = node; if (node == 0) break; }
      A.path_len[slot_id] = len;
      if (A.path_cost) {                           // path_cost(path), eval_gnn.py:53-58: float32 norms accumulated in float32
        float cost = 0.f;
        for (int j = 0; j + 1 < len; ++j) {
          const float* p0 = vg + (size_t)out[j] * dim;
          const float* p1 = vg + (size_t)out[j + 1] * dim;
          float s2 = 0.f;
          for (int q = 0; q < dim; ++q) { const float dq = __fsub_rn(p1[q], p0[q]); s2 = __fadd_rn(s2, __fmul_rn(dq, dq)); }
          cost = __fadd_rn(cost, __fsqrt_rn(s2));
        }
        A.path_cost[slot_id] = cost;
      }
    }
  }
  // speculative checks that were never committed
  int wasted = 0;
  for (int sl = tid; sl < E; sl += kSearchThreads) {
    const int rec = chk[sl];
    if ((rec & 3) != 0 && !(rec & 0x40000000)) wasted += 1;
  }
  for (int o = 16; o > 0; o >>= 1) wasted += __shfl_xor_sync(0xffffffffu, wasted, o);
  if (lane == 0 && wasted) atomicAdd(A.n_spec + slot_id, wasted);
  (void)n_spec;
}


inline int64_t tree_search_carve(Carver& cv, SearchArgs& A, int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total) {
  A.in_ptr = cv.take<int32_t>(n_nodes_total + n_graphs + 1);
  A.cursor = cv.take<int32_t>(n_nodes_total); A.bestv = cv.take<float>(n_nodes_total); A.bests = cv.take<int32_t>(n_nodes_total);
  A.csr_src = cv.take<int32_t>(n_edges_total); A.csr_val = cv.take<float>(n_edges_total); A.chk = cv.take<int32_t>(n_edges_total);
  return cv.bytes();
}

}  // namespace gmp
