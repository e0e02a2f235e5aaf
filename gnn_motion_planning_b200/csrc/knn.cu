// Batched k-NN random-geometric-graph construction (sm_100a).
//
// Replaces create_data's graph build, reference eval_gnn.py:159-164:
//     edge_index      = knn_graph(v, k1, loop=True)                 (torch_cluster)
//     edge_index      = cat(edge_index, edge_index.flip(0))
//     edge_index_free = knn_graph(v[:n_free], k1, loop=True)
//     edge_index      = cat(edge_index, edge_index_free, edge_index_free.flip(0))
//     edge_index, _   = coalesce(edge_index, None, N, N)            (torch_sparse)
// for a packed batch of independent graphs.
//
// Design (no sort at all): the coalesced, symmetrised edge set of a graph IS its adjacency
// bit-matrix read in row-major order.  So
//   1. knn_select_kernel : one warp per (graph, centre i, pass).  The warp computes the fp32 squared
//      distances from i to every candidate into its private shared-memory strip (coalesced reads of v),
//      finds the k-th smallest distance by a 31-step bisection on the float bit pattern (distances are
//      >= 0, so their bit patterns order like the values), then walks the candidates in index order and
//      sets bits (i,j) and (j,i) for d < T plus the first (k - #less) ties -- i.e. neighbours ordered by
//      (distance, index), the canonical rule of oracle/knn_graph.py.
//   2. row_count / scan / emit : popcount rows, exclusive scans (rows within a graph, graphs within the
//      batch), then each warp writes its row's set bits in increasing column order.  Sorted + unique by
//      construction, bit-exact and deterministic.
// Distance arithmetic uses __fmul_rn/__fadd_rn so no FMA contraction changes a rounding:
//   d = ((0 + d0*d0) + d1*d1) + ...   left to right over the dims.
//
// Bound: integer/bit work out of L1/L2 (a graph's v is <= 112 KB); HBM traffic is the output
// (16 B per edge, int64 pairs) plus N*c*4 B in.  See DESIGN.md.
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"

namespace gmp {
namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ int find_graph(const int32_t* __restrict__ ptr, int n_graphs, int x) {
  // largest g with ptr[g] <= x
  int lo = 0, hi = n_graphs;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (ptr[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// work item = (global row r, pass): pass 0 = all nodes, pass 1 = free-only sub-graph.
__global__ void __launch_bounds__(kWarpsPerCta * 32) knn_select_kernel(
    const float* __restrict__ v, int c, const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ n_free,
    const int32_t* __restrict__ k1s, const int64_t* __restrict__ bm_ptr, int n_graphs, int n_rows_total, int strip,
    uint32_t* __restrict__ bitmap) {
  extern __shared__ float smem_dist[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dist = smem_dist + (size_t)warp * strip;
  const int pass = blockIdx.y;
  for (int r = blockIdx.x * kWarpsPerCta + warp; r < n_rows_total; r += gridDim.x * kWarpsPerCta) {
    const int g = find_graph(node_ptr, n_graphs, r);
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int i = r - n0;
    const int nf = n_free[g];
    int cnt;  // candidate count
    if (pass == 0) cnt = n;
    else {
      if (nf >= n || i >= nf) continue;  // identical to pass 0 / row not in the free sub-graph
      cnt = nf;
    }
    const int k = min(k1s[g], cnt);
    const float* vg = v + (size_t)n0 * c;
    const int wpr = (n + 31) >> 5;
    uint32_t* bm = bitmap + bm_ptr[g];

    // 1. distances
    for (int j = lane; j < cnt; j += 32) {
      float d = 0.0f;
      for (int q = 0; q < c; ++q) {
        float diff = __fsub_rn(__ldg(vg + (size_t)i * c + q), __ldg(vg + (size_t)j * c + q));
        d = __fadd_rn(d, __fmul_rn(diff, diff));
      }
      dist[j] = d;
    }
    __syncwarp();

    // 2. k-th smallest distance (value) by bisection over the bit pattern
    uint32_t T = 0;
    int n_less = 0;
    if (k < cnt) {
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int cl = 0;
        for (int j = lane; j < cnt; j += 32) cl += (__float_as_uint(dist[j]) < cand) ? 1 : 0;
        cl = __reduce_add_sync(0xffffffffu, cl);
        if (cl < k) { T = cand; n_less = cl; }
      }
      // n_less must be count(d < T) for the FINAL T: recount (cheap) to keep the logic obvious
      int cl = 0;
      for (int j = lane; j < cnt; j += 32) cl += (__float_as_uint(dist[j]) < T) ? 1 : 0;
      n_less = __reduce_add_sync(0xffffffffu, cl);
    } else {
      T = 0xffffffffu;  // everything is "less"
    }
    int quota = k - n_less;  // ties (d == T) to take, lowest index first

    // 3. mark (i,j) and (j,i)
    for (int base = 0; base < cnt; base += 32) {
      const int j = base + lane;
      bool less = false, tie = false;
      if (j < cnt) {
        uint32_t b = __float_as_uint(dist[j]);
        less = b < T;
        tie = (b == T) && (k < cnt);
      }
      const uint32_t tb = __ballot_sync(0xffffffffu, tie);
      const int rank = __popc(tb & ((1u << lane) - 1u));
      const bool take = less || (tie && rank < quota);
      quota -= min(quota, __popc(tb));
      if (take) {
        atomicOr(bm + (size_t)i * wpr + (j >> 5), 1u << (j & 31));
        atomicOr(bm + (size_t)j * wpr + (i >> 5), 1u << (i & 31));
      }
    }
    __syncwarp();
  }
}

// Register-strip variant for graphs with at most 32*NPL nodes (the BASELINE sizes: 1000 -> NPL 32, 2000 -> NPL 64):
// each lane keeps the bit patterns of its NPL candidate distances in registers, so the 31 bisection passes are pure
// ALU work (no shared-memory re-reads -- the strip version is bound by the single LSU port).  Same selection rule.
// Register budget: the strip (NPL registers) plus a few temporaries.  Left to itself ptxas unrolls the 31 bisection passes and
// takes 223 registers (one 8-warp CTA per SM, issue slots 47 % busy -- ncu, profiles/r1_knn_select_ncu.txt); capping the
// registers and keeping the bit loop rolled gives 3-5 CTAs per SM.
template <int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32, NPL <= 16 ? 5 : (NPL <= 32 ? 4 : 2)) knn_select_reg_kernel(
    const float* __restrict__ v, int c, const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ n_free,
    const int32_t* __restrict__ k1s, const int64_t* __restrict__ bm_ptr, int n_graphs, int n_rows_total,
    uint32_t* __restrict__ bitmap) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pass = blockIdx.y;
  for (int r = blockIdx.x * kWarpsPerCta + warp; r < n_rows_total; r += gridDim.x * kWarpsPerCta) {
    const int g = find_graph(node_ptr, n_graphs, r);
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int i = r - n0;
    const int nf = n_free[g];
    int cnt;
    if (pass == 0) cnt = n;
    else {
      if (nf >= n || i >= nf) continue;
      cnt = nf;
    }
    const int k = min(k1s[g], cnt);
    const float* vg = v + (size_t)n0 * c;
    const int wpr = (n + 31) >> 5;
    uint32_t* bm = bitmap + bm_ptr[g];

    uint32_t d[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) d[t] = 0xffffffffu;   // padding: never below any threshold
    for (int q = 0; q < c; ++q) {
      const float xi = __ldg(vg + (size_t)i * c + q);
#pragma unroll
      for (int t = 0; t < NPL; ++t) {
        const int j = lane + 32 * t;
        if (j < cnt) {
          const float diff = __fsub_rn(xi, __ldg(vg + (size_t)j * c + q));
          const float prev = q == 0 ? 0.0f : __uint_as_float(d[t]);
          d[t] = __float_as_uint(__fadd_rn(prev, __fmul_rn(diff, diff)));
        }
      }
    }
    uint32_t T = 0xffffffffu;
    int n_less = cnt;
    if (k < cnt) {
      T = 0;
#pragma unroll 1
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // four independent count chains
#pragma unroll
        for (int t = 0; t < NPL; t += 4) {
          c0 += (d[t] < cand) ? 1 : 0; c1 += (d[t + 1] < cand) ? 1 : 0; c2 += (d[t + 2] < cand) ? 1 : 0; c3 += (d[t + 3] < cand) ? 1 : 0;
        }
        const int cl = __reduce_add_sync(0xffffffffu, (c0 + c1) + (c2 + c3));
        if (cl < k) T = cand;
      }
      int cl = 0;
#pragma unroll
      for (int t = 0; t < NPL; ++t) cl += (d[t] < T) ? 1 : 0;
      n_less = __reduce_add_sync(0xffffffffu, cl);
    }
    int quota = k - min(n_less, k);
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int j = lane + 32 * t;
      const bool in = j < cnt;
      const bool less = in && d[t] < T;
      const bool tie = in && (d[t] == T) && (k < cnt);
      const uint32_t tb = __ballot_sync(0xffffffffu, tie);
      const bool take = less || (tie && __popc(tb & ((1u << lane) - 1u)) < quota);
      quota -= min(quota, __popc(tb));
      if (take) {
        atomicOr(bm + (size_t)i * wpr + (j >> 5), 1u << (j & 31));
        atomicOr(bm + (size_t)j * wpr + (i >> 5), 1u << (i & 31));
      }
    }
  }
}

// Round-2 select kernel (graphs of at most 32*NPL nodes whose transposed node block fits in shared memory):
//  * one CTA serves a CHUNK of centres of one graph; the graph's candidate block is staged once in shared memory, TRANSPOSED
//    (vT[q][j], odd pitch): the distance loop reads lane-contiguous, conflict-free words instead of c-strided global words
//    (c = 14: 56-byte stride, 14 sectors per warp load in the register-strip kernel above);
//  * the k-th smallest distance is found in TWO levels instead of 31 bisection passes over every candidate: 12 passes fix the
//    top bits (sign-less exponent + 4 mantissa bits); the candidates that share those bits with the threshold -- a handful --
//    are compacted into one register per lane (ballot + find-n-th-set + shuffle) and the remaining 19 bits are bisected on that
//    single register.  Same threshold, same tie rule (distance, then index): the edge set is bit-identical.
template <int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32, NPL <= 32 ? 4 : 2) knn_select_smem_kernel(
    const float* __restrict__ v, int c, const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ n_free,
    const int32_t* __restrict__ k1s, const int64_t* __restrict__ bm_ptr, int centres_per_cta, uint32_t* __restrict__ bitmap) {
  extern __shared__ float vT[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.y, pass = blockIdx.z;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0, nf = n_free[g];
  if (pass == 1 && nf >= n) return;                        // the free-only graph is the whole graph
  const int cnt = pass == 0 ? n : nf;                      // candidates = centres of this pass
  const int i_lo = blockIdx.x * centres_per_cta, i_hi = min(i_lo + centres_per_cta, cnt);
  if (i_lo >= i_hi) return;
  const int pitch = (cnt + 1) | 1;
  const float* vg = v + (size_t)n0 * c;
  for (int idx = threadIdx.x; idx < cnt * c; idx += kWarpsPerCta * 32) {
    const int j = idx / c, q = idx - j * c;
    vT[q * pitch + j] = __ldg(vg + idx);
  }
  __syncthreads();
  const int k = min(k1s[g], cnt);
  const int wpr = (n + 31) >> 5;
  uint32_t* bm = bitmap + bm_ptr[g];
  constexpr int L = 16;                                    // bits resolved on the compacted bucket
  for (int i = i_lo + warp; i < i_hi; i += kWarpsPerCta) {
    uint32_t d[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) d[t] = 0xffffffffu;      // padding: never below any threshold
    for (int q = 0; q < c; ++q) {
      const float xi = vT[q * pitch + i];
      const float* col = vT + q * pitch + lane;
#pragma unroll
      for (int t = 0; t < NPL; ++t) {
        if (lane + 32 * t < cnt) {
          const float diff = __fsub_rn(xi, col[32 * t]);
          const float prev = q == 0 ? 0.0f : __uint_as_float(d[t]);
          d[t] = __float_as_uint(__fadd_rn(prev, __fmul_rn(diff, diff)));
        }
      }
    }
    uint32_t T = 0xffffffffu;
    int n_less = cnt;
    if (k < cnt) {
      T = 0;
      auto count_below = [&](uint32_t cand) {
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int t = 0; t < NPL; t += 4) {
          c0 += (d[t] < cand) ? 1 : 0; c1 += (d[t + 1] < cand) ? 1 : 0; c2 += (d[t + 2] < cand) ? 1 : 0; c3 += (d[t + 3] < cand) ? 1 : 0;
        }
        return __reduce_add_sync(0xffffffffu, (c0 + c1) + (c2 + c3));
      };
      int n_below = 0;                                     // count(d < T) for the current T
      // first level on the top 16 bits only, two candidates per instruction: a non-negative float's upper half IS its
      // truncated bfloat16, the candidate thresholds of this level have a zero lower half (d < cand  <=>  hi(d) < hi(cand)), so
      // one HSET2.BF16 compares two packed distances and one HADD2.BF16 counts them (counts <= NPL, exact in bfloat16).  The
      // padding word 0xffff is a NaN: never below anything, like 0xffffffff in the integer compare.  hi(cand) never reaches the
      // inf / NaN patterns: a candidate with all exponent bits set would have to satisfy count(d < inf) < k with k < cnt.
      uint32_t dp[NPL / 2];
#pragma unroll
      for (int t = 0; t < NPL / 2; ++t) dp[t] = __byte_perm(d[2 * t], d[2 * t + 1], 0x7632);
      auto count_below16 = [&](uint32_t cand) {
        const uint32_t c16 = cand >> 16, cc = c16 | (c16 << 16);
        const __nv_bfloat162 c2 = *reinterpret_cast<const __nv_bfloat162*>(&cc);
        __nv_bfloat162 a0 = __float2bfloat162_rn(0.0f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int t = 0; t < NPL / 2; t += 4) {
          a0 = __hadd2(a0, __hlt2(*reinterpret_cast<const __nv_bfloat162*>(&dp[t]), c2));
          a1 = __hadd2(a1, __hlt2(*reinterpret_cast<const __nv_bfloat162*>(&dp[t + 1]), c2));
          a2 = __hadd2(a2, __hlt2(*reinterpret_cast<const __nv_bfloat162*>(&dp[t + 2]), c2));
          a3 = __hadd2(a3, __hlt2(*reinterpret_cast<const __nv_bfloat162*>(&dp[t + 3]), c2));
        }
        const __nv_bfloat162 a = __hadd2(__hadd2(a0, a1), __hadd2(a2, a3));
        return __reduce_add_sync(0xffffffffu, (int)(__low2float(a) + __high2float(a)));
      };
#pragma unroll 1
      for (int bit = 30; bit >= L; --bit) {
        const uint32_t cand = T | (1u << bit);
        const int cl = count_below16(cand);
        if (cl < k) { T = cand; n_below = cl; }
      }
      // bucket: candidates with the threshold's top bits -> one register per lane
      uint32_t e = 0xffffffffu;
      int nb = 0;
#pragma unroll
      for (int t = 0; t < NPL; ++t) {
        const bool inb = (d[t] >> L) == (T >> L);
        const uint32_t b = __ballot_sync(0xffffffffu, inb);
        if (b) {
          const int want = lane - nb;                      // this lane takes the want-th bucket member of this step
          const int pc = __popc(b);
          const uint32_t src = (want >= 0 && want < pc) ? __fns(b, 0, want + 1) : 0u;
          const uint32_t val = __shfl_sync(0xffffffffu, d[t], src & 31);
          if (want >= 0 && want < pc) e = val;
          nb += pc;
        }
      }
      if (nb <= 32) {
#pragma unroll 1
        for (int bit = L - 1; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          const int cl = n_below + __popc(__ballot_sync(0xffffffffu, e < cand));
          if (cl < k) T = cand;
        }
        n_less = n_below + __popc(__ballot_sync(0xffffffffu, e < T));
      } else {                                             // (many near-equal distances: finish on the full strip)
#pragma unroll 1
        for (int bit = L - 1; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          if (count_below(cand) < k) T = cand;
        }
        n_less = count_below(T);
      }
    }
    int quota = k - min(n_less, k);
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int j = lane + 32 * t;
      const bool in = j < cnt;
      const bool less = in && d[t] < T;
      const bool tie = in && (d[t] == T) && (k < cnt);
      const uint32_t tb = __ballot_sync(0xffffffffu, tie);
      const bool take = less || (tie && __popc(tb & ((1u << lane) - 1u)) < quota);
      quota -= min(quota, __popc(tb));
      if (take) {
        atomicOr(bm + (size_t)i * wpr + (j >> 5), 1u << (j & 31));
        atomicOr(bm + (size_t)j * wpr + (i >> 5), 1u << (i & 31));
      }
    }
  }
}

__global__ void __launch_bounds__(256) row_count_kernel(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ node_ptr,
                                                        const int64_t* __restrict__ bm_ptr, int n_graphs, int n_rows_total,
                                                        int32_t* __restrict__ row_count) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < n_rows_total; r += gridDim.x * 8) {
    const int g = find_graph(node_ptr, n_graphs, r);
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int wpr = (n + 31) >> 5;
    const uint32_t* row = bitmap + bm_ptr[g] + (size_t)(r - n0) * wpr;
    int cnt = 0;
    for (int w = lane; w < wpr; w += 32) cnt += __popc(row[w]);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) row_count[r] = cnt;
  }
}

// One CTA per graph: exclusive scan of its row counts -> row_off (local), graph total -> graph_edges[g].
__global__ void __launch_bounds__(256) row_scan_kernel(const int32_t* __restrict__ row_count, const int32_t* __restrict__ node_ptr,
                                                       int32_t* __restrict__ row_off, int32_t* __restrict__ graph_edges) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int g = blockIdx.x;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    const int x = i < n ? row_count[n0 + i] : 0;
    int incl = x;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wprefix = 0;
    for (int w = 0; w < warp; ++w) wprefix += s_warp[w];
    const int carry = s_carry;
    if (i < n) row_off[n0 + i] = carry + wprefix + incl - x;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + wprefix + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) graph_edges[g] = s_carry;
}

// Single CTA: exclusive scan over graphs -> edge_ptr[B+1].
__global__ void __launch_bounds__(256) graph_scan_kernel(const int32_t* __restrict__ graph_edges, int n_graphs,
                                                         int32_t* __restrict__ edge_ptr) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_graphs; base += 256) {
    const int i = base + threadIdx.x;
    const int x = i < n_graphs ? graph_edges[i] : 0;
    int incl = x;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wprefix = 0;
    for (int w = 0; w < warp; ++w) wprefix += s_warp[w];
    const int carry = s_carry;
    if (i < n_graphs) edge_ptr[i] = carry + wprefix + incl - x;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + wprefix + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) edge_ptr[n_graphs] = s_carry;
}

// One warp per row: write (row, col) for each set bit, columns ascending.
__global__ void __launch_bounds__(256) emit_kernel(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ node_ptr,
                                                   const int64_t* __restrict__ bm_ptr, const int32_t* __restrict__ row_off,
                                                   const int32_t* __restrict__ edge_ptr, int n_graphs, int n_rows_total,
                                                   int64_t* __restrict__ edge_index, int64_t capacity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < n_rows_total; r += gridDim.x * 8) {
    const int g = find_graph(node_ptr, n_graphs, r);
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int wpr = (n + 31) >> 5;
    const int i = r - n0;
    const uint32_t* row = bitmap + bm_ptr[g] + (size_t)i * wpr;
    int64_t pos = (int64_t)edge_ptr[g] + row_off[r];
    for (int wb = 0; wb < wpr; wb += 32) {
      const int w = wb + lane;
      uint32_t bits = w < wpr ? row[w] : 0u;
      const int pc = __popc(bits);
      int incl = pc;
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      int64_t p = pos + incl - pc;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        if (p < capacity) {
          edge_index[p] = i;                         // row 0: source
          edge_index[capacity + p] = (w << 5) + b;   // row 1: target
        }
        ++p;
      }
      pos += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

struct KnnWs {
  int32_t* node_ptr; int32_t* n_free; int32_t* k1; int64_t* bm_ptr;
  uint32_t* bitmap; int32_t* row_count; int32_t* row_off; int32_t* graph_edges;
  int64_t bitmap_words;
};

int64_t carve_knn(Carver& cv, KnnWs& ws, int64_t n_graphs, int64_t n_nodes_total, int64_t bitmap_words) {
  ws.node_ptr = cv.take<int32_t>(n_graphs + 1);
  ws.n_free = cv.take<int32_t>(n_graphs);
  ws.k1 = cv.take<int32_t>(n_graphs);
  ws.bm_ptr = cv.take<int64_t>(n_graphs + 1);
  ws.bitmap = cv.take<uint32_t>(bitmap_words);
  ws.row_count = cv.take<int32_t>(n_nodes_total);
  ws.row_off = cv.take<int32_t>(n_nodes_total);
  ws.graph_edges = cv.take<int32_t>(n_graphs);
  ws.bitmap_words = bitmap_words;
  return cv.bytes();
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" int64_t gmp_knn_graph_max_edges(int64_t n_nodes, int k1) {
  int64_t k = k1 < n_nodes ? k1 : n_nodes;
  int64_t m = 4 * n_nodes * k;
  return m < n_nodes * n_nodes ? m : n_nodes * n_nodes;
}

extern "C" int64_t gmp_knn_graph_workspace_bytes(int64_t n_graphs, int64_t n_nodes_total, int64_t max_nodes_per_graph,
                                                 int /*k1_max*/) {
  // bitmap upper bound: every graph as wide as the widest
  int64_t wpr = (max_nodes_per_graph + 31) / 32;
  Carver cv(nullptr);
  KnnWs ws;
  return carve_knn(cv, ws, n_graphs, n_nodes_total, n_nodes_total * wpr) + 256;
}

extern "C" int gmp_knn_graph(gmp_handle* /*h*/, int64_t n_graphs, const float* v, int c, const int32_t* node_ptr_h,
                             const int32_t* n_free_h, const int32_t* k1_h, int64_t* edge_index_out, int64_t edge_capacity,
                             int32_t* edge_ptr_out, void* workspace, int64_t workspace_bytes, void* stream) {
  GMP_REQUIRE(n_graphs >= 0, "n_graphs < 0");
  GMP_REQUIRE(node_ptr_h && n_free_h && k1_h && edge_ptr_out, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_graphs == 0) {
    GMP_CUDA(cudaMemsetAsync(edge_ptr_out, 0, sizeof(int32_t), st));
    return GMP_OK;
  }
  GMP_REQUIRE(v && edge_index_out && workspace, "null pointer");
  GMP_REQUIRE(c >= 1 && c <= 64, "config_size out of range [1,64]");
  const int64_t n_total = node_ptr_h[n_graphs];
  int64_t max_n = 0, need_cap = 0;
  std::string err;
  int64_t* bm_ptr_h = new int64_t[n_graphs + 1];
  bm_ptr_h[0] = 0;
  for (int64_t g = 0; g < n_graphs; ++g) {
    int64_t n = (int64_t)node_ptr_h[g + 1] - node_ptr_h[g];
    if (n < 0 || n_free_h[g] < 0 || n_free_h[g] > n || k1_h[g] < 1) err = "bad node_ptr / n_free / k1 for a graph";
    if (n > max_n) max_n = n;
    need_cap += gmp_knn_graph_max_edges(n, k1_h[g]);
    bm_ptr_h[g + 1] = bm_ptr_h[g] + n * ((n + 31) / 32);
  }
  const int64_t bitmap_words = bm_ptr_h[n_graphs];
  if (err.empty() && max_n > 6144) err = "graphs with more than 6144 nodes are not supported by the k-NN kernel";
  if (err.empty() && need_cap > edge_capacity) err = "edge_capacity < sum_g min(4*N_g*k1_g, N_g^2)";
  if (err.empty() && n_total >= (int64_t)1 << 30) err = "too many nodes";
  Carver cv(workspace);
  KnnWs ws;
  int64_t need = carve_knn(cv, ws, n_graphs, n_total, bitmap_words);
  if (err.empty() && need > workspace_bytes) err = "workspace too small";
  if (!err.empty()) {
    delete[] bm_ptr_h;
    set_error("gmp_knn_graph: " + err);
    return GMP_E_INVALID;
  }
  cudaError_t e = cudaMemcpyAsync(ws.node_ptr, node_ptr_h, (n_graphs + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws.n_free, n_free_h, n_graphs * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws.k1, k1_h, n_graphs * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws.bm_ptr, bm_ptr_h, (n_graphs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // bm_ptr_h is a temporary (pageable): make the copy complete
  delete[] bm_ptr_h;
  GMP_CUDA(e);
  GMP_CUDA(cudaMemsetAsync(ws.bitmap, 0, bitmap_words * sizeof(uint32_t), st));
  if (n_total > 0) {
    const int strip = (int)align_up(max_n, 32);
    const size_t smem = (size_t)kWarpsPerCta * strip * sizeof(float);
    GMP_CUDA(cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = (int)((n_total + kWarpsPerCta - 1) / kWarpsPerCta);
    if (gx > kNumSMs * 16) gx = kNumSMs * 16;
    // round-2 kernel: transposed node block in shared memory + two-level threshold search (per-graph grid)
    const size_t smem_t = (size_t)c * (size_t)((max_n + 1) | 1) * sizeof(float);
    const bool use_smem = max_n <= 2048 && smem_t <= 113 * 1024 && n_graphs <= 65535 && !getenv("GMP_KNN_LEGACY");
    if (use_smem) {
      // ~16 centres per warp: enough to amortise the staging of the node block, enough CTAs to fill the device
      const int cpc = 128;
      const dim3 grid((unsigned)((max_n + cpc - 1) / cpc), (unsigned)n_graphs, 2);
#define GMP_KNN_SMEM(NPL_)                                                                                                      \
      {                                                                                                                         \
        GMP_CUDA(cudaFuncSetAttribute(knn_select_smem_kernel<NPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
        knn_select_smem_kernel<NPL_><<<grid, kWarpsPerCta * 32, smem_t, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr, cpc, ws.bitmap); \
      }
      if (max_n <= 256) GMP_KNN_SMEM(8)
      else if (max_n <= 512) GMP_KNN_SMEM(16)
      else if (max_n <= 1024) GMP_KNN_SMEM(32)
      else GMP_KNN_SMEM(64)
#undef GMP_KNN_SMEM
    } else
    if (max_n <= 256)
      knn_select_reg_kernel<8><<<dim3(gx, 2), kWarpsPerCta * 32, 0, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr, (int)n_graphs,
                                                                          (int)n_total, ws.bitmap);
    else if (max_n <= 512)
      knn_select_reg_kernel<16><<<dim3(gx, 2), kWarpsPerCta * 32, 0, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr, (int)n_graphs,
                                                                           (int)n_total, ws.bitmap);
    else if (max_n <= 1024)
      knn_select_reg_kernel<32><<<dim3(gx, 2), kWarpsPerCta * 32, 0, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr, (int)n_graphs,
                                                                           (int)n_total, ws.bitmap);
    else if (max_n <= 2048)
      knn_select_reg_kernel<64><<<dim3(gx, 2), kWarpsPerCta * 32, 0, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr, (int)n_graphs,
                                                                           (int)n_total, ws.bitmap);
    else
      knn_select_kernel<<<dim3(gx, 2), kWarpsPerCta * 32, smem, st>>>(v, c, ws.node_ptr, ws.n_free, ws.k1, ws.bm_ptr,
                                                                      (int)n_graphs, (int)n_total, strip, ws.bitmap);
    GMP_LAUNCH_CHECK();
    int gr = (int)((n_total + 7) / 8);
    if (gr > kNumSMs * 16) gr = kNumSMs * 16;
    row_count_kernel<<<gr, 256, 0, st>>>(ws.bitmap, ws.node_ptr, ws.bm_ptr, (int)n_graphs, (int)n_total, ws.row_count);
    GMP_LAUNCH_CHECK();
  }
  row_scan_kernel<<<(int)n_graphs, 256, 0, st>>>(ws.row_count, ws.node_ptr, ws.row_off, ws.graph_edges);
  GMP_LAUNCH_CHECK();
  graph_scan_kernel<<<1, 256, 0, st>>>(ws.graph_edges, (int)n_graphs, edge_ptr_out);
  GMP_LAUNCH_CHECK();
  if (n_total > 0) {
    int gr = (int)((n_total + 7) / 8);
    if (gr > kNumSMs * 16) gr = kNumSMs * 16;
    emit_kernel<<<gr, 256, 0, st>>>(ws.bitmap, ws.node_ptr, ws.bm_ptr, ws.row_off, edge_ptr_out, (int)n_graphs, (int)n_total,
                                    edge_index_out, edge_capacity);
    GMP_LAUNCH_CHECK();
  }
  return GMP_OK;
}
