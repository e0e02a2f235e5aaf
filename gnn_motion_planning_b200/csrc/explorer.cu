// GNN path-explorer forward for a packed batch of graphs (sm_100a).
//
// Replaces EncoderProcessDecoder.forward, reference model.py:115-150, with its sub-modules
//   MPNN.forward / message   model.py:30-41      (max aggregation, model.py:82)
//   Attention.forward        model.py:164-181
//   FeedForward.forward      model.py:193-201
//   Block.forward            model.py:212-218
// The reference materialises [E, 1+O, e] attention tensors and [E, 5e] / [E, 3e] concatenations in
// memory and walks the graph with gather + scatter_max library calls, one graph per call.  Here:
//
//   csr_*            target-sorted CSR of the COO edge list (count / scan / fill), so that edge rows that
//                    aggregate into the same node are adjacent;
//   goal_index       per-graph argmin (knn(v, goal, k=1), model.py:132);
//   obstacle_kernel  the obstacle token stream of both Block stacks.  It never sees map rows
//                    (model.py:215-216), so per block it emits tiles  M = scale * Wq^T (Wk o)  and
//                    V = Wv o  that the map rows consume -- attention over obstacles becomes two small
//                    GEMMs against a shared-memory tile, flash-attention style (online softmax), and
//                    the [M, 1+O, e] tensor never exists;
//   node_pre_kernel  node encoders + 3 node Blocks; emits the loop-invariant parts of `encoder` and
//                    `decoder` (X0, D0) and h_0;
//   edge_feature_kernel  edge encoders + 3 edge Blocks; emits the loop-invariant part of
//                    lin_0[0] (P = W4 ef + W5 ec + b) and of policy[0] (Q = Wc ef + b), in CSR order;
//   node_loop_kernel x (loop+1)   h = lin_1([x, agg]); x = X0 + We4 h; A = (W1+W2) x; B = (W3-W1) x
//                    -- lin_0[0] over cat(x_j - x_i, x_j, x_i, .) split algebraically so the per-edge
//                    first layer is a 3-way add; last call emits decoder + policy[0] node terms;
//   edge_msg_kernel x loop        m = lin_0[2](relu(A[src] + B[dst] + P)); segmented max over the
//                    CSR-adjacent rows of a target, one RED.MAX per (segment, feature);
//   policy_kernel    logit = policy[4](relu(policy[2](relu(G[src] + H[dst] + Q)))), written at the
//                    edge's COO position and (optionally) into the dense [N, N] matrix.
//
// All dense math is fp32 FMA through the row-tile machinery of rowtile.cuh (see there for the layout).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "rowtile.cuh"
#include "explorer_tc.cuh"
#include "explorer_tc64.cuh"

namespace gmp {
namespace {

constexpr float kLnEps = 1e-6f;  // model.py:163,189

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up of a row-tile CTA
// ------------------------------------------------------------------------------------------------
template <int E, bool WITH_S = true>
struct Smem {
  using Cf = RowCfg<E>;
  static constexpr int kBuf = E * Cf::RP;                 // one feature-major activation buffer
  static constexpr int kWB = (E > 64 ? E : 64) * E + 4 * E;  // one weight stage: matrix (K <= max(E, 4c <= 64)) + up to 4 vectors
  static constexpr int kOB = 2 * E * Cf::OT;              // one obstacle tile: Mt [E][OT] | V [OT][E]
  static constexpr int kFloats = (WITH_S ? 2 : 1) * kBuf + kWB + kOB;
  static constexpr size_t kBytes = (size_t)kFloats * sizeof(float) + 2 * Cf::R * sizeof(int);
  float* X; float* S; float* WB; float* OB; int* IDX; int* SRC;
  // the hot per-edge kernels run without the scratch buffer S (chained through registers / X reused in place):
  // half the shared memory per CTA, twice the resident warps
  __device__ explicit Smem(float* base) {
    X = base; S = WITH_S ? X + kBuf : nullptr; WB = X + (WITH_S ? 2 : 1) * kBuf; OB = WB + kWB;
    IDX = reinterpret_cast<int*>(OB + kOB);
    SRC = IDX + Cf::R;
  }
};

template <int TM, int N>
__device__ __forceinline__ void acc_add_vec(float (&acc)[TM][N], const float* __restrict__ vec) {
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    const float4 b = *reinterpret_cast<const float4*>(vec + n);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      acc[r][n] += b.x; acc[r][n + 1] += b.y; acc[r][n + 2] += b.z; acc[r][n + 3] += b.w;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// one Block on the map rows of this tile: X <- map_feed(attention(X, obstacles))   (model.py:212-215)
// `tab` points at this graph's first obstacle tile for this block; tiles are 2*E*OT floats apart.
// ------------------------------------------------------------------------------------------------
template <int E, typename SM>
__device__ __forceinline__ void map_block(const SM& sm, const float* __restrict__ W, const BlockW bw,
                                          const float* __restrict__ tab, int n_obs) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, RP = Cf::RP, OT = Cf::OT;
  float* xcol = sm.X + threadIdx.x;
  float acc[TM][E];
  float m[TM], l[TM];

  // self score  s = x^T (scale Wq^T Wk) x, scale = log2(e)/temperature      (model.py:175,177)
  stage_load(sm.WB, W + bw.Gt, E * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    float s = 0.0f;
#pragma unroll
    for (int n = 0; n < E; ++n) s = fmaf(acc[r][n], xcol[n * RP + r * kRtThreads], s);
    m[r] = s;
    l[r] = 1.0f;
  }
  // value of the row itself, weight exp(s - m) = 1                          (model.py:165)
  stage_load(sm.WB, W + bw.Wvt, E * E + 2 * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);

  // obstacle tiles, online softmax                                          (model.py:174,176-180)
  const int n_tiles = (n_obs + OT - 1) / OT;
  for (int t = 0; t < n_tiles; ++t) {
    stage_load(sm.OB, tab + (size_t)t * (2 * E * OT), 2 * E * OT);
    float s[TM][OT];
    acc_zero(s);
    gemm_smem<E, OT, TM, RP>(s, xcol, sm.OB);
    const int o_left = n_obs - t * OT;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      float tmax = -INFINITY;
#pragma unroll
      for (int o = 0; o < OT; ++o) tmax = (o < o_left) ? fmaxf(tmax, s[r][o]) : tmax;
      const float mnew = fmaxf(m[r], tmax);
      const float corr = exp2f(m[r] - mnew);
      float lsum = l[r] * corr;
#pragma unroll
      for (int n = 0; n < E; ++n) acc[r][n] *= corr;
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        const float p = (o < o_left) ? exp2f(s[r][o] - mnew) : 0.0f;
        lsum += p;
        s[r][o] = p;
      }
      l[r] = lsum;
      m[r] = mnew;
    }
    gemm_reg<OT, E, TM>(acc, s, sm.OB + E * OT);   // P.V with the probabilities still in registers
  }
  // softmax normalisation, residual, attention.layer_norm                   (model.py:181)
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const float inv = 1.0f / l[r];
#pragma unroll
    for (int n = 0; n < E; ++n) acc[r][n] = fmaf(acc[r][n], inv, xcol[n * RP + r * kRtThreads]);
  }
  acc_layernorm(acc, sm.WB + E * E, sm.WB + E * E + E, kLnEps);
  acc_store<TM, E, RP>(acc, xcol);

  // map_feed: the residual stays in registers, the hidden layer overwrites X in place   (model.py:193-201)
  float h[TM][E];
  stage_load(sm.WB, W + bw.W1t, E * E + E);
  acc_zero(h);
  gemm_smem<E, E, TM, RP>(h, xcol, sm.WB);
  acc_add_vec(h, sm.WB + E * E);
  acc_relu(h);
  acc_store<TM, E, RP>(h, xcol);
  stage_load(sm.WB, W + bw.W2t, E * E + 3 * E);
  acc_zero(h);
  gemm_smem<E, E, TM, RP>(h, xcol, sm.WB);
  acc_add_vec(h, sm.WB + E * E);
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int n = 0; n < E; ++n) h[r][n] += acc[r][n];
  acc_layernorm(h, sm.WB + E * E + E, sm.WB + E * E + 2 * E, kLnEps);
  acc_store<TM, E, RP>(h, xcol);
}

// two-layer encoder Seq(Lin, ReLU, Lin) with the first layer's inputs in registers: result -> `out` column.
// `hid` is the scratch column for the hidden layer (may equal `out`).
template <int K, int E, typename SM>
__device__ __forceinline__ void encoder_mlp(const SM& sm, const float* __restrict__ W, int off0, int off2,
                                            const float (&in)[RowCfg<E>::TM][K], float* hid, float* out) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, RP = Cf::RP;
  static_assert(K <= 64, "input layer wider than the weight stage");
  float acc[TM][E];
  stage_load(sm.WB, W + off0, K * E + E);
  acc_zero(acc);
  gemm_reg<K, E, TM>(acc, in, sm.WB);
  acc_add_vec(acc, sm.WB + K * E);
  acc_relu(acc);
  acc_store<TM, E, RP>(acc, hid);
  stage_load(sm.WB, W + off2, E * E + E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, hid, sm.WB);
  acc_add_vec(acc, sm.WB + E * E);
  acc_store<TM, E, RP>(acc, out);
}

// ------------------------------------------------------------------------------------------------
// CSR build
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_count_kernel(const int64_t* __restrict__ edge_index, int64_t row_stride,
                                                        const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ node_ptr,
                                                        int n_graphs, int n_edges, int32_t* __restrict__ indeg, int32_t* __restrict__ bad) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    const int g = find_segment(edge_ptr, n_graphs, e);
    const int n = node_ptr[g + 1] - node_ptr[g];
    const int64_t s64 = edge_index[e], d64 = edge_index[row_stride + e];
    // local ids outside [0, N_g) would write outside the workspace: they are clamped (memory safe, result undefined for that
    // edge) and counted -- gmp_explorer_bad_edges() reports the count of the last forward
    if (s64 < 0 || s64 >= n || d64 < 0 || d64 >= n) atomicAdd(bad, 1);
    const int dst = (int)min(max(d64, (int64_t)0), (int64_t)max(n - 1, 0));
    atomicAdd(indeg + node_ptr[g] + dst, 1);
  }
}

// one CTA per graph: in_ptr[node] = edge_ptr[g] + exclusive prefix of indeg; cursor reset
__global__ void __launch_bounds__(256) csr_scan_kernel(const int32_t* __restrict__ indeg, const int32_t* __restrict__ node_ptr,
                                                       const int32_t* __restrict__ edge_ptr, int32_t* __restrict__ in_ptr,
                                                       int32_t* __restrict__ cursor) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int g = blockIdx.x;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = edge_ptr[g];
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    const int x = i < n ? indeg[n0 + i] : 0;
    int incl = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wprefix = 0;
    for (int w = 0; w < warp; ++w) wprefix += s_warp[w];
    const int carry = s_carry;
    if (i < n) {
      in_ptr[n0 + i] = carry + wprefix + incl - x;
      cursor[n0 + i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + wprefix + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int64_t* __restrict__ edge_index, int64_t row_stride,
                                                       const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ node_ptr,
                                                       int n_graphs, int n_edges, const int32_t* __restrict__ in_ptr,
                                                       int32_t* __restrict__ cursor, int32_t* __restrict__ csr_src,
                                                       int32_t* __restrict__ csr_dst, int32_t* __restrict__ csr_eid) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
    const int g = find_segment(edge_ptr, n_graphs, e);
    const int n0 = node_ptr[g], nm1 = max(node_ptr[g + 1] - n0 - 1, 0);
    const int src = n0 + (int)min(max(edge_index[e], (int64_t)0), (int64_t)nm1);                  // (clamped: see csr_count_kernel)
    const int dst = n0 + (int)min(max(edge_index[row_stride + e], (int64_t)0), (int64_t)nm1);
    const int slot = in_ptr[dst] + atomicAdd(cursor + dst, 1);
    csr_src[slot] = src;
    csr_dst[slot] = dst;
    csr_eid[slot] = e;
  }
}

// ------------------------------------------------------------------------------------------------
// goal index: argmin_i ||goal - v_i||^2, canonical fp32 rule, first minimum        (model.py:132)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) goal_index_kernel(const float* __restrict__ v, const float* __restrict__ goal, int c,
                                                         const int32_t* __restrict__ node_ptr, int32_t* __restrict__ goal_idx) {
  __shared__ float s_d[256];
  __shared__ int s_i[256];
  const int g = blockIdx.x;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += 256) {
    float d = 0.0f;
    for (int q = 0; q < c; ++q) {
      const float diff = __fsub_rn(__ldg(goal + (size_t)g * c + q), __ldg(v + (size_t)(n0 + i) * c + q));
      d = __fadd_rn(d, __fmul_rn(diff, diff));
    }
    if (d < best) { best = d; bi = i; }  // ascending i per thread: keeps the first minimum
  }
  s_d[threadIdx.x] = best;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float d2 = s_d[threadIdx.x + o];
      const int i2 = s_i[threadIdx.x + o];
      if (d2 < s_d[threadIdx.x] || (d2 == s_d[threadIdx.x] && i2 < s_i[threadIdx.x])) {
        s_d[threadIdx.x] = d2;
        s_i[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) goal_idx[g] = (n > 0 && s_i[0] != 0x7fffffff) ? s_i[0] : 0;
}

// ------------------------------------------------------------------------------------------------
// obstacle token stream (rows = padded obstacle slots of the whole batch; blockIdx.y = stream)
// ------------------------------------------------------------------------------------------------
template <int S, int E>
__global__ void __launch_bounds__(kRtThreads) obstacle_kernel(ExplorerW w, const float* __restrict__ W,
                                                              const float* __restrict__ obstacles,
                                                              const int32_t* __restrict__ obs_ptr,
                                                              const int32_t* __restrict__ obs_tile_ptr, int n_graphs,
                                                              int n_slots, float* __restrict__ tables,
                                                              int64_t table_stride /* floats per (stream, blk) */) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP, OT = Cf::OT;
  extern __shared__ __align__(16) float smem_raw[];
  Smem<E> sm(smem_raw);
  const int stream = blockIdx.y;
  float* xcol = sm.X + threadIdx.x;
  float* scol = sm.S + threadIdx.x;

  int slot[TM], o_local[TM];
  bool valid[TM];
  float in[TM][S];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    slot[r] = blockIdx.x * R + r * kRtThreads + threadIdx.x;
    valid[r] = false;
    o_local[r] = 0;
#pragma unroll
    for (int k = 0; k < S; ++k) in[r][k] = 0.0f;
    if (slot[r] < n_slots) {
      const int tile = slot[r] / OT;
      const int g = find_segment(obs_tile_ptr, n_graphs, tile);
      o_local[r] = slot[r] - obs_tile_ptr[g] * OT;
      const int n_obs = obs_ptr[g + 1] - obs_ptr[g];
      valid[r] = o_local[r] < n_obs;
      if (valid[r]) {
#pragma unroll
        for (int k = 0; k < S; ++k) in[r][k] = __ldg(obstacles + (size_t)(obs_ptr[g] + o_local[r]) * S + k);
      }
    }
  }
  encoder_mlp<S, E>(sm, W, w.obs0[stream], w.obs2[stream], in, xcol, xcol);   // model.py:126-127

  float acc[TM][E];
  for (int blk = 0; blk < 3; ++blk) {
    const ObsBlockW bw = w.obs_blk[stream][blk];
    float* tab = tables + (size_t)(stream * 3 + blk) * table_stride;
    // key -> S; M = scale Wq^T key -> table (transposed tile [E][OT])
    stage_load(sm.WB, W + bw.Wkt, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
    acc_store<TM, E, RP>(acc, scol);
    stage_load(sm.WB, W + bw.WqS, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (slot[r] < n_slots) {
        float* chunk = tab + (size_t)(slot[r] / OT) * (2 * E * OT);
        const int oo = slot[r] % OT;
#pragma unroll
        for (int k = 0; k < E; ++k) chunk[k * OT + oo] = valid[r] ? acc[r][k] : 0.0f;
      }
    }
    // value -> table [OT][E]
    stage_load(sm.WB, W + bw.Wvt, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (slot[r] < n_slots) {
        float* chunk = tab + (size_t)(slot[r] / OT) * (2 * E * OT) + E * OT + (slot[r] % OT) * E;
        if (!valid[r]) {
#pragma unroll
          for (int n = 0; n < E; ++n) acc[r][n] = 0.0f;
        }
        acc_store_global<TM, E>(acc, r, chunk);
      }
    }
    if (blk == 2) break;
    // obs_feed                                                              (model.py:216)
    stage_load(sm.WB, W + bw.W1t, E * E + E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
    acc_add_vec(acc, sm.WB + E * E);
    acc_relu(acc);
    acc_store<TM, E, RP>(acc, scol);
    stage_load(sm.WB, W + bw.W2t, E * E + 3 * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
    acc_add_vec(acc, sm.WB + E * E);
    acc_add_col<TM, E, RP>(acc, xcol);
    acc_layernorm(acc, sm.WB + E * E + E, sm.WB + E * E + 2 * E, kLnEps);
    acc_store<TM, E, RP>(acc, xcol);
  }
}

// ------------------------------------------------------------------------------------------------
// node encoders + node Blocks -> X0, D0, h0                          (model.py:119,122,128-135,141,143)
// grid: one CTA per (graph, node tile)
// ------------------------------------------------------------------------------------------------
template <int C, int E>
__global__ void __launch_bounds__(kRtThreads) node_pre_kernel(ExplorerW w, const float* __restrict__ W, const float* __restrict__ v,
                                                              const float* __restrict__ goal, const int32_t* __restrict__ node_ptr,
                                                              const int32_t* __restrict__ tile_ptr, int n_graphs,
                                                              const int32_t* __restrict__ obs_ptr,
                                                              const int32_t* __restrict__ obs_tile_ptr,
                                                              const float* __restrict__ tables, int64_t table_stride,
                                                              int use_obstacles, const int32_t* __restrict__ goal_idx,
                                                              float* __restrict__ X0, float* __restrict__ D0, float* __restrict__ H) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP, OT = Cf::OT;
  extern __shared__ __align__(16) float smem_raw[];
  Smem<E> sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  float* scol = sm.S + threadIdx.x;
  const int g = find_segment(tile_ptr, n_graphs, blockIdx.x);
  const int n0 = node_ptr[g], n1 = node_ptr[g + 1];
  const int row0 = n0 + (blockIdx.x - tile_ptr[g]) * R;

  int row[TM];
  bool valid[TM];
  {
    float in[TM][C];
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      row[r] = row0 + r * kRtThreads + threadIdx.x;
      valid[r] = row[r] < n1;
#pragma unroll
      for (int k = 0; k < C; ++k) in[r][k] = valid[r] ? __ldg(v + (size_t)row[r] * C + k) : 0.0f;
    }
    encoder_mlp<C, E>(sm, W, w.nf0, w.nf2, in, scol, xcol);                  // node_free_code, model.py:122
  }
  if (use_obstacles) {
    const int n_obs = obs_ptr[g + 1] - obs_ptr[g];
    for (int blk = 0; blk < 3; ++blk) {
      const float* tab = tables + (size_t)(0 * 3 + blk) * table_stride + (size_t)obs_tile_ptr[g] * (2 * E * OT);
      map_block<E>(sm, W, w.node_blk[blk], tab, n_obs);                      // model.py:129
    }
  }
  {
    float in[TM][4 * C];                                                     // model.py:119
#pragma unroll
    for (int r = 0; r < TM; ++r) {
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float x = valid[r] ? __ldg(v + (size_t)row[r] * C + k) : 0.0f;
        const float gk = __ldg(goal + (size_t)g * C + k);
        const float d = x - gk;
        in[r][k] = x;
        in[r][C + k] = gk;
        in[r][2 * C + k] = d * d;
        in[r][3 * C + k] = d;
      }
    }
    encoder_mlp<4 * C, E>(sm, W, w.nc0, w.nc2, in, scol, scol);              // node_code -> S
  }
  float acc[TM][E];
  // X0 = We1 nc + We2 nf + b (+ We3 goal_encoder on the goal row)           (model.py:141, loop-invariant part)
  stage_load(sm.WB, W + w.enc_nc, E * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
  stage_load(sm.WB, W + w.enc_nf, E * E + E);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
  acc_add_vec(acc, sm.WB + E * E);
  const int gi = n0 + goal_idx[g];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    if (valid[r]) {
      if (row[r] == gi) {
#pragma unroll
        for (int n = 0; n < E; ++n) acc[r][n] += __ldg(W + w.enc_u3 + n);
      }
      acc_store_global<TM, E>(acc, r, X0 + (size_t)row[r] * E);
    }
  }
  // D0 = Wd1 nc + b                                                         (model.py:143)
  stage_load(sm.WB, W + w.dec_nc, E * E + E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
  acc_add_vec(acc, sm.WB + E * E);
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    if (valid[r]) {
      acc_store_global<TM, E>(acc, r, D0 + (size_t)row[r] * E);
      // h_0: goal_encoder on the goal row, zero elsewhere                   (model.py:133-135)
#pragma unroll
      for (int n = 0; n < E; ++n) acc[r][n] = (row[r] == gi) ? __ldg(W + w.goal_enc + n) : 0.0f;
      acc_store_global<TM, E>(acc, r, H + (size_t)row[r] * E);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// edge encoders + edge Blocks -> P, Q (CSR order)                     (model.py:120,123,130; :39,:145)
// grid: one CTA per (graph, edge tile)
// ------------------------------------------------------------------------------------------------
template <int C, int E>
__global__ void __launch_bounds__(kRtThreads, 3) edge_feature_kernel(
    ExplorerW w, const float* __restrict__ W, const float* __restrict__ v, const int32_t* __restrict__ csr_src,
    const int32_t* __restrict__ csr_dst, const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ tile_ptr, int n_graphs,
    const int32_t* __restrict__ obs_ptr, const int32_t* __restrict__ obs_tile_ptr, const float* __restrict__ tables,
    int64_t table_stride, int use_obstacles, float* __restrict__ P, float* __restrict__ Q) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP, OT = Cf::OT;
  extern __shared__ __align__(16) float smem_raw[];
  Smem<E, false> sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  const int g = find_segment(tile_ptr, n_graphs, blockIdx.x);
  const int slot0 = edge_ptr[g] + (blockIdx.x - tile_ptr[g]) * R;
  const int slot1 = edge_ptr[g + 1];

  int slot[TM];
  bool valid[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    slot[r] = slot0 + r * kRtThreads + threadIdx.x;
    valid[r] = slot[r] < slot1;
  }
  auto gather_in = [&](float (&in)[TM][2 * C]) {                              // cat(v[src], v[dst])
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int s = valid[r] ? csr_src[slot[r]] : 0;
      const int d = valid[r] ? csr_dst[slot[r]] : 0;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        in[r][k] = valid[r] ? __ldg(v + (size_t)s * C + k) : 0.0f;
        in[r][C + k] = valid[r] ? __ldg(v + (size_t)d * C + k) : 0.0f;
      }
    }
  };
  {
    float in[TM][2 * C];
    gather_in(in);
    encoder_mlp<2 * C, E>(sm, W, w.ef0, w.ef2, in, xcol, xcol);              // edge_free_code, model.py:123 (hidden in place)
  }
  if (use_obstacles) {
    const int n_obs = obs_ptr[g + 1] - obs_ptr[g];
    for (int blk = 0; blk < 3; ++blk) {
      const float* tab = tables + (size_t)(1 * 3 + blk) * table_stride + (size_t)obs_tile_ptr[g] * (2 * E * OT);
      map_block<E>(sm, W, w.edge_blk[blk], tab, n_obs);                      // model.py:130
    }
  }
  float acc[TM][E];
  // Q = Wc ef + b : loop-invariant part of policy[0]                        (model.py:145-146)
  stage_load(sm.WB, W + w.p0_ef, E * E + E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
  acc_add_vec(acc, sm.WB + E * E);
#pragma unroll
  for (int r = 0; r < TM; ++r)
    if (valid[r]) acc_store_global<TM, E>(acc, r, Q + (size_t)slot[r] * E);
  // P = W4 ef + W5 ec + b  : loop-invariant part of lin_0[0]                (model.py:39,142)
  // W4 ef stays in registers while edge_code is produced in place in X (ef is dead after this GEMM).
  float pacc[TM][E];
  stage_load(sm.WB, W + w.l0_ef, E * E);
  acc_zero(pacc);
  gemm_smem<E, E, TM, RP>(pacc, xcol, sm.WB);
  {
    float in[TM][2 * C];
    gather_in(in);
    encoder_mlp<2 * C, E>(sm, W, w.ec0, w.ec2, in, xcol, xcol);              // edge_code, model.py:120
  }
  stage_load(sm.WB, W + w.l0_ec, E * E + E);
  gemm_smem<E, E, TM, RP>(pacc, xcol, sm.WB);
  acc_add_vec(pacc, sm.WB + E * E);
#pragma unroll
  for (int r = 0; r < TM; ++r)
    if (valid[r]) acc_store_global<TM, E>(pacc, r, P + (size_t)slot[r] * E);
}

// ------------------------------------------------------------------------------------------------
// per-round node update.  mode 0: first round (h = h_0); 1: middle; 2: after the last round (decoder).
// rows are flat over the batch.
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kRtThreads) node_loop_kernel(ExplorerW w, const float* __restrict__ W, int mode, int n_rows,
                                                               const float* __restrict__ X0, const float* __restrict__ D0,
                                                               float* __restrict__ H, float* __restrict__ Xg,
                                                               float* __restrict__ AGG, float* __restrict__ A, float* __restrict__ B) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP;
  extern __shared__ __align__(16) float smem_raw[];
  Smem<E> sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  float* scol = sm.S + threadIdx.x;
  int row[TM];
  bool valid[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    row[r] = blockIdx.x * R + r * kRtThreads + threadIdx.x;
    valid[r] = row[r] < n_rows;
  }
  float acc[TM][E];
  if (mode == 0) {
#pragma unroll
    for (int r = 0; r < TM; ++r)
      col_load_global<E, RP>(scol + r * kRtThreads, valid[r] ? H + (size_t)row[r] * E : nullptr);
  } else {
    // h = lin_1([x, agg])                                                   (model.py:36)
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      col_load_global<E, RP>(xcol + r * kRtThreads, valid[r] ? Xg + (size_t)row[r] * E : nullptr);
      col_load_global<E, RP>(scol + r * kRtThreads, valid[r] ? AGG + (size_t)row[r] * E : nullptr);
#pragma unroll
      for (int n = 0; n < E; ++n) {  // rows with no incoming edge aggregate to 0 (scatter_max fill)
        float* p = scol + r * kRtThreads + n * RP;
        if (*p == -INFINITY) *p = 0.0f;
      }
    }
    stage_load(sm.WB, W + w.l1_x, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
    stage_load(sm.WB, W + w.l1_a, E * E + E);
    gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
    acc_add_vec(acc, sm.WB + E * E);
    acc_store<TM, E, RP>(acc, scol);  // h -> S
  }
  if (mode == 2) {
    // decode = D0 + Wd2 h ; G = (Wa+Wb) decode ; Hp = -Wb decode            (model.py:143,145)
    stage_load(sm.WB, W + w.dec_h, E * E);
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int n = 0; n < E; n += 4) {
        const float4 x = valid[r] ? __ldg(reinterpret_cast<const float4*>(D0 + (size_t)row[r] * E + n))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[r][n] = x.x; acc[r][n + 1] = x.y; acc[r][n + 2] = x.z; acc[r][n + 3] = x.w;
      }
    gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
    acc_store<TM, E, RP>(acc, xcol);
    stage_load(sm.WB, W + w.p0_G, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
    for (int r = 0; r < TM; ++r)
      if (valid[r]) acc_store_global<TM, E>(acc, r, A + (size_t)row[r] * E);
    stage_load(sm.WB, W + w.p0_H, E * E);
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
    for (int r = 0; r < TM; ++r)
      if (valid[r]) acc_store_global<TM, E>(acc, r, B + (size_t)row[r] * E);
    return;
  }
  // x = X0 + We4 h                                                          (model.py:141)
  stage_load(sm.WB, W + w.enc_h, E * E);
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int n = 0; n < E; n += 4) {
      const float4 x = valid[r] ? __ldg(reinterpret_cast<const float4*>(X0 + (size_t)row[r] * E + n))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[r][n] = x.x; acc[r][n + 1] = x.y; acc[r][n + 2] = x.z; acc[r][n + 3] = x.w;
    }
  gemm_smem<E, E, TM, RP>(acc, scol, sm.WB);
  acc_store<TM, E, RP>(acc, xcol);
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    if (valid[r]) {
      acc_store_global<TM, E>(acc, r, Xg + (size_t)row[r] * E);
      // reset the aggregation target of this round
      const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
      for (int n = 0; n < E; n += 4) *reinterpret_cast<float4*>(AGG + (size_t)row[r] * E + n) = ninf;
    }
  }
  // A = (W1+W2) x  (gathered at the source) ; B = (W3-W1) x  (gathered at the target)   (model.py:39)
  stage_load(sm.WB, W + w.l0_A, E * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
  for (int r = 0; r < TM; ++r)
    if (valid[r]) acc_store_global<TM, E>(acc, r, A + (size_t)row[r] * E);
  stage_load(sm.WB, W + w.l0_B, E * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, xcol, sm.WB);
#pragma unroll
  for (int r = 0; r < TM; ++r)
    if (valid[r]) acc_store_global<TM, E>(acc, r, B + (size_t)row[r] * E);
}

// hidden = relu(A[src] + B[dst] + PQ[slot]) for the R rows of this tile -> X (feature-major), IDX[row] = dst.
// The tile's CSR indices are staged in shared memory first (coalesced), then E/4 lanes share a row (one float4 each):
// a warp-wide LDG.128 covers 32/(E/4) whole rows -- full sectors and 8x fewer L1 wavefronts than one-row-per-lane
// gathers -- with 8 rows (24 LDG.128) in flight per thread.  The transposed shared-memory store is conflict-free
// because RP = R + 1.  Callers must barrier before reading X / IDX (stage_load does).
template <int E, typename SM>
__device__ __forceinline__ void edge_hidden(const SM& sm, int slot0, int n_slots, const int32_t* __restrict__ csr_src,
                                            const int32_t* __restrict__ csr_dst, const float* __restrict__ A,
                                            const float* __restrict__ B, const float* __restrict__ PQ) {
  using Cf = RowCfg<E>;
  constexpr int R = Cf::R, RP = Cf::RP;
  constexpr int LPR = E / 4;                 // lanes per row
  constexpr int RPP = kRtThreads / LPR;      // rows per pass of the CTA
  for (int i = threadIdx.x; i < R; i += kRtThreads) {
    const int slot = slot0 + i;
    const bool ok = slot < n_slots;
    sm.SRC[i] = ok ? __ldg(csr_src + slot) : -1;
    sm.IDX[i] = ok ? __ldg(csr_dst + slot) : -1;
  }
  __syncthreads();
  const int q = threadIdx.x % LPR;
  const int rsub = threadIdx.x / LPR;
  constexpr int UB = 8;   // rows per thread in flight: all 24 row loads are issued before the first is consumed
  for (int base = rsub; base < R; base += UB * RPP) {
    float4 av[UB], bv[UB], pv[UB];
    int sv[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int r = base + u * RPP;
      const int s = sm.SRC[r], d = sm.IDX[r];
      sv[u] = s;
      av[u] = __ldg(reinterpret_cast<const float4*>(A + (size_t)max(s, 0) * E) + q);
      bv[u] = __ldg(reinterpret_cast<const float4*>(B + (size_t)max(d, 0) * E) + q);
      pv[u] = s >= 0 ? __ldg(reinterpret_cast<const float4*>(PQ + (size_t)(slot0 + r) * E) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int r = base + u * RPP;
      const bool ok = sv[u] >= 0;
      float* x = sm.X + (4 * q) * RP + r;
      x[0] = ok ? fmaxf(av[u].x + bv[u].x + pv[u].x, 0.0f) : 0.0f;
      x[RP] = ok ? fmaxf(av[u].y + bv[u].y + pv[u].y, 0.0f) : 0.0f;
      x[2 * RP] = ok ? fmaxf(av[u].z + bv[u].z + pv[u].z, 0.0f) : 0.0f;
      x[3 * RP] = ok ? fmaxf(av[u].w + bv[u].w + pv[u].w, 0.0f) : 0.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// messages + max aggregation for one round; rows = CSR slots, flat over the batch   (model.py:33,38-41)
// ------------------------------------------------------------------------------------------------
// Persistent CTAs (grid = resident CTAs, tiles strided over them).  Per tile:
//   wait for the tile's P block in the staging buffer (landed there by a TMA bulk copy issued one tile earlier);
//   hidden = relu(A[src] + B[dst] + P) -> X          (P from shared memory, A / B rows gathered from L2, coalesced);
//   issue the TMA bulk copy of the NEXT tile's P block into the staging buffer  -- the HBM stream of the
//   loop-invariant edge term is therefore always one tile ahead of the FMAs;
//   m = lin_0[2](hidden)  (weights staged once per CTA);  segmented max -> RED.MAX.
template <int E>
struct MsgSmem {
  using Cf = RowCfg<E>;
  static constexpr int kX = E * Cf::RP;
  static constexpr int kPS = Cf::R * E;          // staging buffer for one P tile, row-major like global memory
  static constexpr int kWB = E * E + 2 * E;
  static constexpr size_t kBytes = (size_t)(kX + kPS + kWB) * sizeof(float) + 2 * Cf::R * sizeof(int) + 16;
};

template <int E>
__global__ void __launch_bounds__(kRtThreads, 3) edge_msg_kernel(ExplorerW w, const float* __restrict__ W, int n_slots,
                                                                 const int32_t* __restrict__ csr_src,
                                                                 const int32_t* __restrict__ csr_dst, const float* __restrict__ A,
                                                                 const float* __restrict__ B, const float* __restrict__ P,
                                                                 float* __restrict__ AGG) {
  using Cf = RowCfg<E>;
  using MS = MsgSmem<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP;
  constexpr int LPR = E / 4, RPP = kRtThreads / LPR;
  extern __shared__ __align__(128) float smem_raw[];
  float* PS = smem_raw;                       // first: 128 B aligned for the bulk copy
  float* X = PS + MS::kPS;
  float* WB = X + MS::kX;
  int* IDX = reinterpret_cast<int*>(WB + MS::kWB);
  int* SRC = IDX + R;
  uint64_t* bar = reinterpret_cast<uint64_t*>(SRC + R);
  const int n_tiles = (n_slots + R - 1) / R;
  int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  auto tile_bytes = [&](int t) { return (uint32_t)(min(R, n_slots - t * R) * E * (int)sizeof(float)); };
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, tile_bytes(tile));
    tma_bulk_g2s(PS, P + (size_t)tile * R * E, tile_bytes(tile), bar);
  }
  // weights of lin_0[2], once per CTA
  for (int i = threadIdx.x; i < (E * E + E) / 4; i += kRtThreads)
    reinterpret_cast<float4*>(WB)[i] = __ldg(reinterpret_cast<const float4*>(W + w.l0_2) + i);
  uint32_t phase = 0;
  const int q = threadIdx.x % LPR, rsub = threadIdx.x / LPR;
  constexpr int IPT = R / kRtThreads;          // CSR indices per thread and tile
  int nsrc[IPT], ndst[IPT];
  auto load_indices = [&](int t) {
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      const int slot = t * R + u * kRtThreads + threadIdx.x;
      const bool ok = t < n_tiles && slot < n_slots;
      nsrc[u] = ok ? __ldg(csr_src + slot) : -1;
      ndst[u] = ok ? __ldg(csr_dst + slot) : -1;
    }
  };
  load_indices(tile);
  for (; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
    for (int u = 0; u < IPT; ++u) {
      SRC[u * kRtThreads + threadIdx.x] = nsrc[u];
      IDX[u * kRtThreads + threadIdx.x] = ndst[u];
    }
    __syncthreads();           // indices visible; mbarrier init visible (first tile); X free (previous scan done)
    load_indices(tile + gridDim.x);   // next tile's indices travel while this tile is processed
    mbar_wait(bar, phase);     // this tile's P block has landed
    phase ^= 1;
    // 8 rows per thread in flight: all 16 row gathers are issued before the first is consumed
    constexpr int UB = 8;
    for (int base = rsub; base < R; base += UB * RPP) {
      float4 av[UB], bv[UB];
      int sv[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int r = base + u * RPP;
        const int s = SRC[r], d = IDX[r];
        sv[u] = s;
        av[u] = __ldg(reinterpret_cast<const float4*>(A + (size_t)max(s, 0) * E) + q);
        bv[u] = __ldg(reinterpret_cast<const float4*>(B + (size_t)max(d, 0) * E) + q);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int r = base + u * RPP;
        const float4 p = *reinterpret_cast<const float4*>(PS + r * E + 4 * q);
        const bool ok = sv[u] >= 0;
        float* x = X + (4 * q) * RP + r;
        x[0] = ok ? fmaxf(av[u].x + bv[u].x + p.x, 0.0f) : 0.0f;
        x[RP] = ok ? fmaxf(av[u].y + bv[u].y + p.y, 0.0f) : 0.0f;
        x[2 * RP] = ok ? fmaxf(av[u].z + bv[u].z + p.z, 0.0f) : 0.0f;
        x[3 * RP] = ok ? fmaxf(av[u].w + bv[u].w + p.w, 0.0f) : 0.0f;
      }
    }
    __syncthreads();           // X complete; staging buffer consumed
    const int next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < n_tiles) {
      mbar_expect_tx(bar, tile_bytes(next));
      tma_bulk_g2s(PS, P + (size_t)next * R * E, tile_bytes(next), bar);
    }
    float acc[TM][E];
    acc_zero(acc);
    gemm_smem<E, E, TM, RP>(acc, X + threadIdx.x, WB);
    acc_add_vec(acc, WB + E * E);
    acc_store<TM, E, RP>(acc, X + threadIdx.x);   // in place: a thread only ever read its own columns of X
    __syncthreads();
    // segmented max: thread = (feature n, row group); rows of a group are walked in CSR order, so a
    // target's rows are consecutive; one RED per (segment, feature), 128 B coalesced across the warp.
    constexpr int GROUPS = kRtThreads / E;
    constexpr int ROWS = R / GROUPS;
    const int n = threadIdx.x % E;
    const int r0 = (threadIdx.x / E) * ROWS;
    int cur = IDX[r0];
    float run = -INFINITY;
    for (int i = 0; i < ROWS; ++i) {
      const int d = IDX[r0 + i];
      if (d != cur) {
        if (cur >= 0) atomic_max_f32(AGG + (size_t)cur * E + n, run);
        cur = d;
        run = -INFINITY;
      }
      run = fmaxf(run, X[n * RP + r0 + i]);
    }
    if (cur >= 0) atomic_max_f32(AGG + (size_t)cur * E + n, run);
    __syncthreads();           // IDX / X reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------------
// policy head; rows = CSR slots, flat over the batch                              (model.py:145-150)
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kRtThreads, 4) policy_kernel(ExplorerW w, const float* __restrict__ W, int n_slots,
                                                               const int32_t* __restrict__ csr_src,
                                                               const int32_t* __restrict__ csr_dst,
                                                               const int32_t* __restrict__ csr_eid, const float* __restrict__ G,
                                                               const float* __restrict__ Hp, const float* __restrict__ Q,
                                                               const int32_t* __restrict__ edge_ptr,
                                                               const int32_t* __restrict__ node_ptr,
                                                               const int64_t* __restrict__ dense_off, int n_graphs,
                                                               float* __restrict__ logits, float* __restrict__ dense) {
  using Cf = RowCfg<E>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP;
  extern __shared__ __align__(16) float smem_raw[];
  Smem<E, false> sm(smem_raw);
  int slot[TM];
  bool valid[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    slot[r] = blockIdx.x * R + r * kRtThreads + threadIdx.x;
    valid[r] = slot[r] < n_slots;
  }
  edge_hidden<E>(sm, blockIdx.x * R, n_slots, csr_src, csr_dst, G, Hp, Q);
  float acc[TM][E];
  stage_load(sm.WB, W + w.p2, E * E + 2 * E);
  acc_zero(acc);
  gemm_smem<E, E, TM, RP>(acc, sm.X + threadIdx.x, sm.WB);
  acc_add_vec(acc, sm.WB + E * E);
  acc_relu(acc);
  const float* w4 = sm.WB + E * E + E;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    float logit = 0.0f;
#pragma unroll
    for (int n = 0; n < E; ++n) logit = fmaf(acc[r][n], w4[n], logit);
    if (valid[r]) {
      logits[csr_eid[slot[r]]] = logit;
      if (dense) {
        const int g = find_segment(edge_ptr, n_graphs, slot[r]);
        const int n0 = node_ptr[g];
        const int64_t ng = node_ptr[g + 1] - n0;
        const int s = csr_src[slot[r]] - n0, d = csr_dst[slot[r]] - n0;
        dense[dense_off[g] + (int64_t)d * ng + s] = logit;   // out[dst, src]   (model.py:149)
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side: weight image
// ------------------------------------------------------------------------------------------------
struct Packer {
  std::vector<float> buf;
  int begin() {
    while (buf.size() % 4) buf.push_back(0.f);
    return (int)buf.size();
  }
  void put(const std::vector<double>& x) {
    for (double d : x) buf.push_back((float)d);
  }
  void put(const std::vector<float>& x) { buf.insert(buf.end(), x.begin(), x.end()); }
};

// torch Linear weight W [out][in] (row-major), optional column window [c0, c0+k) -> K-major Wt[k][n] (doubles)
std::vector<double> transpose_window(const std::vector<float>& W, int out, int in, int c0, int k) {
  std::vector<double> t((size_t)k * out);
  for (int n = 0; n < out; ++n)
    for (int j = 0; j < k; ++j) t[(size_t)j * out + n] = W[(size_t)n * in + c0 + j];
  return t;
}

struct Spec { const char* name; int64_t numel; };

// round-to-nearest (ties away) TF32 of an fp32 value, as cvt.rna.tf32.f32
float tf32_rna_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  std::memcpy(&x, &u, 4);
  return x;
}

// B[n][k] (N x K) -> [hi plane | lo plane] for the tensor-core kernels; a plane is float[KP/4][N][4] (K-major,
// no swizzle), K zero padded to KP (umma.cuh)
void put_planes(Packer& pk, const std::vector<double>& B, int N, int K, int KP) {
  std::vector<float> hi((size_t)N * KP, 0.f), lo((size_t)N * KP, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = (float)B[(size_t)n * K + k];
      const float h = tf32_rna_host(w);
      const size_t i = ((size_t)(k / 4) * N + n) * 4 + (k % 4);
      hi[i] = h;
      lo[i] = tf32_rna_host(w - h);
    }
  pk.put(hi);
  pk.put(lo);
}

// rows [r0, r0 + n) x columns [c0, c0 + k) of a torch Linear weight W [out][in] as doubles
std::vector<double> window(const std::vector<float>& W, int in, int r0, int n, int c0, int k) {
  std::vector<double> t((size_t)n * k);
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < k; ++b) t[(size_t)a * k + b] = W[(size_t)(r0 + a) * in + c0 + b];
  return t;
}

}  // namespace

int explorer_build_image(ExplorerModel& m) {
  const int c = m.c, e = m.e, s = m.s;
  auto T = [&](const std::string& name) -> const std::vector<float>& { return m.tensors.at(name); };
  // ---- validate
  std::vector<std::pair<std::string, int64_t>> need = {
      {"goal_encoder", e},
      {"node_code.0.weight", (int64_t)e * 4 * c}, {"node_code.0.bias", e}, {"node_code.2.weight", (int64_t)e * e}, {"node_code.2.bias", e},
      {"edge_code.0.weight", (int64_t)e * 2 * c}, {"edge_code.0.bias", e}, {"edge_code.2.weight", (int64_t)e * e}, {"edge_code.2.bias", e},
      {"node_free_code.0.weight", (int64_t)e * c}, {"node_free_code.0.bias", e}, {"node_free_code.2.weight", (int64_t)e * e}, {"node_free_code.2.bias", e},
      {"edge_free_code.0.weight", (int64_t)e * 2 * c}, {"edge_free_code.0.bias", e}, {"edge_free_code.2.weight", (int64_t)e * e}, {"edge_free_code.2.bias", e},
      {"obs_node_code.0.weight", (int64_t)e * s}, {"obs_node_code.0.bias", e}, {"obs_node_code.2.weight", (int64_t)e * e}, {"obs_node_code.2.bias", e},
      {"obs_edge_code.0.weight", (int64_t)e * s}, {"obs_edge_code.0.bias", e}, {"obs_edge_code.2.weight", (int64_t)e * e}, {"obs_edge_code.2.bias", e},
      {"encoder.weight", (int64_t)e * 4 * e}, {"encoder.bias", e},
      {"process.lin_0.0.weight", (int64_t)e * 5 * e}, {"process.lin_0.0.bias", e}, {"process.lin_0.2.weight", (int64_t)e * e}, {"process.lin_0.2.bias", e},
      {"process.lin_1.weight", (int64_t)e * 2 * e}, {"process.lin_1.bias", e},
      {"decoder.weight", (int64_t)e * 2 * e}, {"decoder.bias", e},
      {"policy.0.weight", (int64_t)e * 3 * e}, {"policy.0.bias", e}, {"policy.2.weight", (int64_t)e * e}, {"policy.2.bias", e},
      {"policy.4.weight", e},
  };
  for (const char* st : {"node_attentions", "edge_attentions"})
    for (int i = 0; i < 3; ++i) {
      std::string p = std::string(st) + "." + std::to_string(i) + ".";
      for (const char* q : {"attention.key.weight", "attention.query.weight", "attention.value.weight", "map_feed.w_1.weight",
                            "map_feed.w_2.weight", "obs_feed.w_1.weight", "obs_feed.w_2.weight"})
        need.push_back({p + q, (int64_t)e * e});
      for (const char* q : {"attention.layer_norm.weight", "attention.layer_norm.bias", "map_feed.w_1.bias", "map_feed.w_2.bias",
                            "map_feed.layer_norm.weight", "map_feed.layer_norm.bias", "obs_feed.w_1.bias", "obs_feed.w_2.bias",
                            "obs_feed.layer_norm.weight", "obs_feed.layer_norm.bias"})
        need.push_back({p + q, e});
    }
  for (auto& kv : need) {
    auto it = m.tensors.find(kv.first);
    if (it == m.tensors.end()) {
      set_error("explorer weights: missing tensor '" + kv.first + "'");
      return GMP_E_STATE;
    }
    if ((int64_t)it->second.size() != kv.second) {
      set_error("explorer weights: tensor '" + kv.first + "' has " + std::to_string(it->second.size()) + " elements, expected " +
                std::to_string(kv.second));
      return GMP_E_INVALID;
    }
  }
  // ---- pack
  Packer pk;
  ExplorerW& w = m.w;
  // 1 / temperature (model.py:208) times log2(e): scores are produced in base-2 units so that the softmax weights are
  // exp2(s - max) -- one MUFU.EX2 instead of the expf sequence; mathematically the same softmax
  const double scale = 1.4426950408889634074 / std::sqrt((double)e);
  auto lin_pack = [&](const std::string& name, int in) {  // [Wt | b]
    int off = pk.begin();
    pk.put(transpose_window(T(name + ".weight"), e, in, 0, in));
    pk.put(T(name + ".bias"));
    return off;
  };
  auto mat_pack = [&](const std::vector<double>& t) {
    int off = pk.begin();
    pk.put(t);
    return off;
  };
  for (int stream = 0; stream < 2; ++stream) {
    const std::string st = stream == 0 ? "node_attentions." : "edge_attentions.";
    for (int i = 0; i < 3; ++i) {
      const std::string p = st + std::to_string(i) + ".";
      const auto& Wq = T(p + "attention.query.weight");
      const auto& Wk = T(p + "attention.key.weight");
      const auto& Wv = T(p + "attention.value.weight");
      BlockW& bw = stream == 0 ? w.node_blk[i] : w.edge_blk[i];
      // G[i][j] = scale * sum_o Wq[o][i] Wk[o][j];  GEMM layout Wt[k=j][n=i] = G[i][j]
      std::vector<double> Gt((size_t)e * e);
      for (int a = 0; a < e; ++a)
        for (int b = 0; b < e; ++b) {
          double sum = 0;
          for (int o = 0; o < e; ++o) sum += (double)Wq[(size_t)o * e + a] * (double)Wk[(size_t)o * e + b];
          Gt[(size_t)b * e + a] = scale * sum;
        }
      bw.Gt = mat_pack(Gt);
      bw.Wvt = mat_pack(transpose_window(Wv, e, e, 0, e));
      pk.put(T(p + "attention.layer_norm.weight"));
      pk.put(T(p + "attention.layer_norm.bias"));
      bw.W1t = lin_pack(p + "map_feed.w_1", e);
      bw.W2t = lin_pack(p + "map_feed.w_2", e);
      pk.put(T(p + "map_feed.layer_norm.weight"));
      pk.put(T(p + "map_feed.layer_norm.bias"));
      ObsBlockW& ob = w.obs_blk[stream][i];
      ob.Wkt = mat_pack(transpose_window(Wk, e, e, 0, e));
      std::vector<double> WqS((size_t)e * e);  // out[n] = sum_k key[k] * scale*Wq[k][n]
      for (size_t q = 0; q < WqS.size(); ++q) WqS[q] = scale * (double)Wq[q];
      ob.WqS = mat_pack(WqS);
      ob.Wvt = mat_pack(transpose_window(Wv, e, e, 0, e));
      ob.W1t = lin_pack(p + "obs_feed.w_1", e);
      ob.W2t = lin_pack(p + "obs_feed.w_2", e);
      pk.put(T(p + "obs_feed.layer_norm.weight"));
      pk.put(T(p + "obs_feed.layer_norm.bias"));
    }
  }
  w.obs0[0] = lin_pack("obs_node_code.0", s); w.obs2[0] = lin_pack("obs_node_code.2", e);
  w.obs0[1] = lin_pack("obs_edge_code.0", s); w.obs2[1] = lin_pack("obs_edge_code.2", e);
  w.nc0 = lin_pack("node_code.0", 4 * c); w.nc2 = lin_pack("node_code.2", e);
  w.nf0 = lin_pack("node_free_code.0", c); w.nf2 = lin_pack("node_free_code.2", e);
  w.ef0 = lin_pack("edge_free_code.0", 2 * c); w.ef2 = lin_pack("edge_free_code.2", e);
  w.ec0 = lin_pack("edge_code.0", 2 * c); w.ec2 = lin_pack("edge_code.2", e);
  {
    const auto& We = T("encoder.weight");  // [e][4e]: node_code | node_free | h0 | h   (model.py:141)
    w.enc_nc = mat_pack(transpose_window(We, e, 4 * e, 0, e));
    w.enc_nf = mat_pack(transpose_window(We, e, 4 * e, e, e));
    pk.put(T("encoder.bias"));
    std::vector<double> u3(e);
    const auto& ge = T("goal_encoder");
    for (int n = 0; n < e; ++n) {
      double sum = 0;
      for (int k = 0; k < e; ++k) sum += (double)We[(size_t)n * 4 * e + 2 * e + k] * (double)ge[k];
      u3[n] = sum;
    }
    w.enc_u3 = mat_pack(u3);
    w.enc_h = mat_pack(transpose_window(We, e, 4 * e, 3 * e, e));
    const auto& Wd = T("decoder.weight");  // [e][2e]: node_code | h   (model.py:143)
    w.dec_nc = mat_pack(transpose_window(Wd, e, 2 * e, 0, e));
    pk.put(T("decoder.bias"));
    w.dec_h = mat_pack(transpose_window(Wd, e, 2 * e, e, e));
    const auto& W0 = T("process.lin_0.0.weight");  // [e][5e]: x_j-x_i | x_j | x_i | edge_free | edge_code  (model.py:39,142)
    auto W1t = transpose_window(W0, e, 5 * e, 0, e), W2t = transpose_window(W0, e, 5 * e, e, e),
         W3t = transpose_window(W0, e, 5 * e, 2 * e, e);
    w.l0_ef = mat_pack(transpose_window(W0, e, 5 * e, 3 * e, e));
    w.l0_ec = mat_pack(transpose_window(W0, e, 5 * e, 4 * e, e));
    pk.put(T("process.lin_0.0.bias"));
    std::vector<double> At(W1t.size()), Bt(W1t.size());
    for (size_t q = 0; q < At.size(); ++q) { At[q] = W1t[q] + W2t[q]; Bt[q] = W3t[q] - W1t[q]; }
    w.l0_A = mat_pack(At);
    w.l0_B = mat_pack(Bt);
    w.l0_2 = lin_pack("process.lin_0.2", e);
    const auto& Wl = T("process.lin_1.weight");  // [e][2e]: x | agg   (model.py:36)
    w.l1_x = mat_pack(transpose_window(Wl, e, 2 * e, 0, e));
    w.l1_a = mat_pack(transpose_window(Wl, e, 2 * e, e, e));
    pk.put(T("process.lin_1.bias"));
    const auto& Wp = T("policy.0.weight");  // [e][3e]: dec_src | dec_src-dec_dst | edge_free   (model.py:145-146)
    auto Wat = transpose_window(Wp, e, 3 * e, 0, e), Wbt = transpose_window(Wp, e, 3 * e, e, e);
    w.p0_ef = mat_pack(transpose_window(Wp, e, 3 * e, 2 * e, e));
    pk.put(T("policy.0.bias"));
    std::vector<double> Gt(Wat.size()), Ht(Wat.size());
    for (size_t q = 0; q < Gt.size(); ++q) { Gt[q] = Wat[q] + Wbt[q]; Ht[q] = -Wbt[q]; }
    w.p0_G = mat_pack(Gt);
    w.p0_H = mat_pack(Ht);
    w.p2 = lin_pack("policy.2", e);
    pk.put(T("policy.4.weight"));
    w.goal_enc = pk.begin();
    pk.put(ge);
  }
  // ---- tensor-core image of the edge-feature stage (e = 32): hi / lo TF32 planes in the order of TcCfg<C>
  w.tc_img = -1;
  w.tc_l02 = -1;
  w.tc_p2 = -1;
  if (e == 32) {
    const int k0 = (2 * c + 7) / 8 * 8, k4 = (2 * c + 3) / 4 * 4;
    w.tc_img = pk.begin();
    {
      // first encoder layers stacked: rows 0..e-1 edge_free_code.0, rows e..2e-1 edge_code.0 (planes + plain fp32 copy)
      std::vector<double> enc0 = window(T("edge_free_code.0.weight"), 2 * c, 0, e, 0, 2 * c);
      const std::vector<double> ec0 = window(T("edge_code.0.weight"), 2 * c, 0, e, 0, 2 * c);
      enc0.insert(enc0.end(), ec0.begin(), ec0.end());
      put_planes(pk, enc0, 2 * e, 2 * c, k0);
      std::vector<float> plain((size_t)2 * e * k4, 0.f);
      for (int n = 0; n < 2 * e; ++n)
        for (int k = 0; k < 2 * c; ++k) plain[(size_t)n * k4 + k] = (float)enc0[(size_t)n * 2 * c + k];
      pk.put(plain);
    }
    put_planes(pk, window(T("edge_free_code.2.weight"), e, 0, e, 0, e), e, e, e);
    for (int i = 0; i < 3; ++i) {
      const std::string p = "edge_attentions." + std::to_string(i) + ".";
      const auto& Wq = T(p + "attention.query.weight");
      const auto& Wk = T(p + "attention.key.weight");
      const auto& Wv = T(p + "attention.value.weight");
      std::vector<double> GV((size_t)2 * e * e);   // rows 0..e-1: G[a][b]; rows e..2e-1: Wv[n][k]
      for (int a = 0; a < e; ++a)
        for (int b = 0; b < e; ++b) {
          double sum = 0;
          for (int o = 0; o < e; ++o) sum += (double)Wq[(size_t)o * e + a] * (double)Wk[(size_t)o * e + b];
          GV[(size_t)a * e + b] = scale * sum;
          GV[(size_t)(e + a) * e + b] = Wv[(size_t)a * e + b];
        }
      put_planes(pk, GV, 2 * e, e, e);
      put_planes(pk, window(T(p + "map_feed.w_1.weight"), e, 0, e, 0, e), e, e, e);
      put_planes(pk, window(T(p + "map_feed.w_2.weight"), e, 0, e, 0, e), e, e, e);
    }
    std::vector<double> pb(e);
    {
      const auto& Wp = T("policy.0.weight");
      const auto& W0 = T("process.lin_0.0.weight");
      std::vector<double> QP = window(Wp, 3 * e, 0, e, 2 * e, e), Pef = window(W0, 5 * e, 0, e, 3 * e, e);
      QP.insert(QP.end(), Pef.begin(), Pef.end());
      put_planes(pk, QP, 2 * e, e, e);
      // edge_code = W_ec2 relu(.) + b_ec2 is only ever consumed through lin_0's edge_code columns W5 (model.py:39,120):
      // fold  W5 edge_code = (W5 W_ec2) relu(.) + W5 b_ec2  so that edge_code is never materialised
      const std::vector<double> W5 = window(W0, 5 * e, 0, e, 4 * e, e);
      const auto& Wec2 = T("edge_code.2.weight");
      const auto& bec2 = T("edge_code.2.bias");
      const auto& b0 = T("process.lin_0.0.bias");
      std::vector<double> W52((size_t)e * e);
      for (int n = 0; n < e; ++n) {
        double bsum = b0[n];
        for (int j = 0; j < e; ++j) bsum += W5[(size_t)n * e + j] * (double)bec2[j];
        pb[n] = bsum;
        for (int k = 0; k < e; ++k) {
          double sum = 0;
          for (int j = 0; j < e; ++j) sum += W5[(size_t)n * e + j] * (double)Wec2[(size_t)j * e + k];
          W52[(size_t)n * e + k] = sum;
        }
      }
      put_planes(pk, W52, e, e, e);
    }
    pk.put(T("edge_free_code.0.bias")); pk.put(T("edge_code.0.bias")); pk.put(T("edge_free_code.2.bias"));
    for (int i = 0; i < 3; ++i) {
      const std::string p = "edge_attentions." + std::to_string(i) + ".";
      pk.put(T(p + "attention.layer_norm.weight")); pk.put(T(p + "attention.layer_norm.bias"));
      pk.put(T(p + "map_feed.w_1.bias")); pk.put(T(p + "map_feed.w_2.bias"));
      pk.put(T(p + "map_feed.layer_norm.weight")); pk.put(T(p + "map_feed.layer_norm.bias"));
    }
    pk.put(T("policy.0.bias"));
    pk.put(pb);
  }
  for (int q = 0; q < 5; ++q) w.tc64_img[q] = -1;
  if (e == 64) {
    // phase images of the embed-64 tensor-core edge-feature stage (Tc64Cfg<C>)
    const int k0 = (2 * c + 7) / 8 * 8;
    std::vector<double> enc0 = window(T("edge_free_code.0.weight"), 2 * c, 0, e, 0, 2 * c);
    const std::vector<double> ec0 = window(T("edge_code.0.weight"), 2 * c, 0, e, 0, 2 * c);
    enc0.insert(enc0.end(), ec0.begin(), ec0.end());
    w.tc64_img[0] = pk.begin();
    put_planes(pk, enc0, 2 * e, 2 * c, k0);
    put_planes(pk, window(T("edge_free_code.2.weight"), e, 0, e, 0, e), e, e, e);
    pk.put(T("edge_free_code.0.bias")); pk.put(T("edge_code.0.bias")); pk.put(T("edge_free_code.2.bias"));
    for (int i = 0; i < 3; ++i) {
      const std::string p = "edge_attentions." + std::to_string(i) + ".";
      const auto& Wq = T(p + "attention.query.weight");
      const auto& Wk = T(p + "attention.key.weight");
      const auto& Wv = T(p + "attention.value.weight");
      std::vector<double> GV((size_t)2 * e * e);
      for (int a = 0; a < e; ++a)
        for (int b = 0; b < e; ++b) {
          double sum = 0;
          for (int o = 0; o < e; ++o) sum += (double)Wq[(size_t)o * e + a] * (double)Wk[(size_t)o * e + b];
          GV[(size_t)a * e + b] = scale * sum;
          GV[(size_t)(e + a) * e + b] = Wv[(size_t)a * e + b];
        }
      w.tc64_img[1 + i] = pk.begin();
      put_planes(pk, GV, 2 * e, e, e);
      put_planes(pk, window(T(p + "map_feed.w_1.weight"), e, 0, e, 0, e), e, e, e);
      put_planes(pk, window(T(p + "map_feed.w_2.weight"), e, 0, e, 0, e), e, e, e);
      pk.put(T(p + "attention.layer_norm.weight")); pk.put(T(p + "attention.layer_norm.bias"));
      pk.put(T(p + "map_feed.w_1.bias")); pk.put(T(p + "map_feed.w_2.bias"));
      pk.put(T(p + "map_feed.layer_norm.weight")); pk.put(T(p + "map_feed.layer_norm.bias"));
    }
    {
      const auto& Wp = T("policy.0.weight");
      const auto& W0 = T("process.lin_0.0.weight");
      std::vector<double> QP = window(Wp, 3 * e, 0, e, 2 * e, e), Pef = window(W0, 5 * e, 0, e, 3 * e, e);
      QP.insert(QP.end(), Pef.begin(), Pef.end());
      const std::vector<double> W5 = window(W0, 5 * e, 0, e, 4 * e, e);
      const auto& Wec2 = T("edge_code.2.weight");
      const auto& bec2 = T("edge_code.2.bias");
      const auto& b0 = T("process.lin_0.0.bias");
      std::vector<double> W52((size_t)e * e), pb(e);
      for (int n = 0; n < e; ++n) {
        double bsum = b0[n];
        for (int j = 0; j < e; ++j) bsum += W5[(size_t)n * e + j] * (double)bec2[j];
        pb[n] = bsum;
        for (int k = 0; k < e; ++k) {
          double sum = 0;
          for (int j = 0; j < e; ++j) sum += W5[(size_t)n * e + j] * (double)Wec2[(size_t)j * e + k];
          W52[(size_t)n * e + k] = sum;
        }
      }
      w.tc64_img[4] = pk.begin();
      put_planes(pk, QP, 2 * e, e, e);
      put_planes(pk, W52, e, e, e);
      put_planes(pk, ec0, e, 2 * c, k0);
      pk.put(T("edge_code.0.bias")); pk.put(T("policy.0.bias")); pk.put(pb);
    }
  }
  // tensor-core message / policy kernels (e = 32 and 64): [hi plane | lo plane | bias | policy.4 weight or padding]
  w.tc_l02 = pk.begin();
  put_planes(pk, window(T("process.lin_0.2.weight"), e, 0, e, 0, e), e, e, e);
  pk.put(T("process.lin_0.2.bias"));
  pk.put(std::vector<float>(e, 0.f));
  w.tc_p2 = pk.begin();
  put_planes(pk, window(T("policy.2.weight"), e, 0, e, 0, e), e, e, e);
  pk.put(T("policy.2.bias"));
  pk.put(T("policy.4.weight"));
  pk.begin();
  for (int q = 0; q < 4 * e + 64; ++q) pk.buf.push_back(0.f);  // slack: stages may over-read up to a few vectors
  if (m.d_weights) cudaFree(m.d_weights);
  m.d_weights = nullptr;
  GMP_CUDA(cudaMalloc(&m.d_weights, pk.buf.size() * sizeof(float)));
  GMP_CUDA(cudaMemcpy(m.d_weights, pk.buf.data(), pk.buf.size() * sizeof(float), cudaMemcpyHostToDevice));
  m.n_weights = (int64_t)pk.buf.size();
  m.ready = true;
  return GMP_OK;
}

namespace {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct ExWs {
  int32_t *node_ptr, *edge_ptr, *obs_ptr, *obs_tile_ptr, *tile_ptr_e, *tile_ptr_n, *goal_idx;
  int64_t* dense_off;
  int32_t *indeg, *in_ptr, *cursor, *csr_src, *csr_dst, *csr_eid;
  float *tables, *X0, *D0, *H, *Xg, *AGG, *A, *B, *P, *Q;
  int64_t* tc_tab_off;   // tensor-core edge-feature stage: per-graph float offset of its obstacle-table units
  float* tc_tables;      // 3 blocks x tc_rows x 128 floats (hi / lo planes of M and V per 96-obstacle chunk)
  int4* tc_unit_meta;    // per 256-slot unit: first slot, graph's end slot, obstacle count, table offset
  int32_t* tc_unit_ptr;  // [B+1] first 256-slot unit of every graph
};

int64_t carve_explorer(Carver& cv, ExWs& ws, int e, int64_t B, int64_t Nt, int64_t Et, int64_t obs_tiles, int64_t tc_rows) {
  const int ot = (e == 32) ? 32 : 16;
  ws.node_ptr = cv.take<int32_t>(B + 1);
  ws.edge_ptr = cv.take<int32_t>(B + 1);
  ws.obs_ptr = cv.take<int32_t>(B + 1);
  ws.obs_tile_ptr = cv.take<int32_t>(B + 1);
  ws.tile_ptr_e = cv.take<int32_t>(B + 1);
  ws.tile_ptr_n = cv.take<int32_t>(B + 1);
  ws.goal_idx = cv.take<int32_t>(B);
  ws.dense_off = cv.take<int64_t>(B + 1);
  ws.indeg = cv.take<int32_t>(Nt + 1);
  ws.in_ptr = cv.take<int32_t>(Nt + 1);
  ws.cursor = cv.take<int32_t>(Nt + 1);
  ws.csr_src = cv.take<int32_t>(Et);
  ws.csr_dst = cv.take<int32_t>(Et);
  ws.csr_eid = cv.take<int32_t>(Et);
  ws.tables = cv.take<float>(6 * obs_tiles * 2 * e * ot);
  ws.X0 = cv.take<float>(Nt * e);
  ws.D0 = cv.take<float>(Nt * e);
  ws.H = cv.take<float>(Nt * e);
  ws.Xg = cv.take<float>(Nt * e);
  ws.AGG = cv.take<float>(Nt * e);
  ws.A = cv.take<float>(Nt * e);
  ws.B = cv.take<float>(Nt * e);
  ws.P = cv.take<float>(Et * e);
  ws.Q = cv.take<float>(Et * e);
  ws.tc_tab_off = cv.take<int64_t>(B + 1);
  ws.tc_tables = cv.take<float>(3 * tc_rows * 4 * e);
  ws.tc_unit_meta = cv.take<int4>(Et / 256 + B + 1);
  ws.tc_unit_ptr = cv.take<int32_t>(B + 1);
  return cv.bytes();
}

template <int C, int E, int S>
int run_forward(gmp_handle* h, int64_t B, const float* v, const int64_t* edge_index, int64_t row_stride, const float* goal,
                const float* obstacles, const int32_t* node_ptr_h, const int32_t* edge_ptr_h, const int32_t* obs_ptr_h, int loop,
                int use_obstacles, float* logits, float* dense, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  using Cf = RowCfg<E>;
  constexpr int R = Cf::R, OT = Cf::OT;
  const ExplorerModel& m = h->ex;
  const int64_t Nt = node_ptr_h[B], Et = edge_ptr_h[B];
  // host-side tiling metadata
  std::vector<int32_t> meta((size_t)(B + 1) * 6);
  std::vector<int64_t> dense_off(B + 1);
  int32_t* obs_ptr = meta.data() + 2 * (B + 1);
  int32_t* obs_tile_ptr = meta.data() + 3 * (B + 1);
  int32_t* tile_e = meta.data() + 4 * (B + 1);
  int32_t* tile_n = meta.data() + 5 * (B + 1);
  std::memcpy(meta.data(), node_ptr_h, (B + 1) * sizeof(int32_t));
  std::memcpy(meta.data() + (B + 1), edge_ptr_h, (B + 1) * sizeof(int32_t));
  obs_ptr[0] = obs_tile_ptr[0] = tile_e[0] = tile_n[0] = 0;
  dense_off[0] = 0;
  for (int64_t g = 0; g < B; ++g) {
    const int64_t n = (int64_t)node_ptr_h[g + 1] - node_ptr_h[g], ne = (int64_t)edge_ptr_h[g + 1] - edge_ptr_h[g];
    const int64_t no = (use_obstacles && obs_ptr_h) ? (int64_t)obs_ptr_h[g + 1] - obs_ptr_h[g] : 0;
    GMP_REQUIRE(n >= 0 && ne >= 0 && no >= 0, "offset arrays must be non-decreasing");
    obs_ptr[g + 1] = obs_ptr[g] + (int32_t)no;
    obs_tile_ptr[g + 1] = obs_tile_ptr[g] + (int32_t)((no + OT - 1) / OT);
    tile_e[g + 1] = tile_e[g] + (int32_t)((ne + R - 1) / R);
    tile_n[g + 1] = tile_n[g] + (int32_t)((n + R - 1) / R);
    dense_off[g + 1] = dense_off[g] + n * n;
  }
  const int64_t obs_tiles = obs_tile_ptr[B];
  // tensor-core edge-feature stage (e = 32): obstacle-table units of tc_per() rows per chunk
  const bool use_tc = E == 32 && m.w.tc_img >= 0 && m.edge_feature_mode != 0;   // tensor-core edge-feature stage
  const bool use_tc_msg = m.w.tc_l02 >= 0 && m.edge_feature_mode != 0;         // tensor-core message / policy kernels
  // embed 64: the phase-split tensor-core stage (explorer_tc64.cuh) handles graphs with at most 32 obstacles (the arms have <= 12)
  bool use_tc64 = E == 64 && m.w.tc64_img[0] >= 0 && m.edge_feature_mode != 0;
  for (int64_t g = 0; g < B && use_tc64; ++g)
    if (obs_ptr[g + 1] - obs_ptr[g] > 32) use_tc64 = false;
  std::vector<int64_t> tc_off(B + 1, 0);
  std::vector<int32_t> unit_ptr(B + 1, 0);
  if (use_tc || use_tc64)
    for (int64_t g = 0; g < B; ++g) {
      const int no = obs_ptr[g + 1] - obs_ptr[g];
      const int nch = E == 64 ? (no > 0) : tc_nchunks(no);
      tc_off[g + 1] = tc_off[g] + (int64_t)nch * tc_per(no, nch) * 4 * E;
      unit_ptr[g + 1] = unit_ptr[g] + (int32_t)(((int64_t)edge_ptr_h[g + 1] - edge_ptr_h[g] + 255) / 256);
    }
  const int64_t tc_rows = tc_off[B] / (4 * E);
  GMP_REQUIRE(tc_off[B] < (int64_t)1 << 31, "too many obstacle rows in one call for the tensor-core table index");
  Carver cv(workspace);
  ExWs ws;
  const int64_t need = carve_explorer(cv, ws, E, B, Nt, Et, obs_tiles, tc_rows);
  GMP_REQUIRE(need <= workspace_bytes, "workspace too small (see gmp_explorer_workspace_bytes)");
  GMP_CUDA(cudaMemcpyAsync(ws.node_ptr, meta.data(), (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.edge_ptr, meta.data() + (B + 1), (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.obs_ptr, obs_ptr, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.obs_tile_ptr, obs_tile_ptr, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.tile_ptr_e, tile_e, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.tile_ptr_n, tile_n, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.dense_off, dense_off.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (use_tc || use_tc64) {
    GMP_CUDA(cudaMemcpyAsync(ws.tc_tab_off, tc_off.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    GMP_CUDA(cudaMemcpyAsync(ws.tc_unit_ptr, unit_ptr.data(), (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  // (pageable sources: cudaMemcpyAsync has staged them before returning, so the vectors may die)

  const size_t smem = Smem<E, true>::kBytes;     // node / obstacle kernels (two activation buffers)
  const size_t smem1 = Smem<E, false>::kBytes;   // per-edge kernels (one activation buffer)
  // per handle, not per process: cudaFuncSetAttribute is per DEVICE, and a process may hold handles on several GPUs
  bool& attr_done = h->ex_attr_done;
  if (!attr_done) {
    GMP_CUDA(cudaFuncSetAttribute(obstacle_kernel<S, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GMP_CUDA(cudaFuncSetAttribute(node_pre_kernel<C, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GMP_CUDA(cudaFuncSetAttribute(edge_feature_kernel<C, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    GMP_CUDA(cudaFuncSetAttribute(node_loop_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GMP_CUDA(cudaFuncSetAttribute(edge_msg_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MsgSmem<E>::kBytes));
    GMP_CUDA(cudaFuncSetAttribute(policy_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    if constexpr (E == 32)
      {
        GMP_CUDA(cudaFuncSetAttribute(edge_feature_tc_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<C>::kSmemBytes));
        GMP_CUDA(cudaFuncSetAttribute(edge_feature_tc_kernel<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<C>::kSmemBytes));
        GMP_CUDA(cudaFuncSetAttribute(edge_feature_tc_kernel<C, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<C>::kSmemBytes));
      }
    if constexpr (E == 64) {
      GMP_CUDA(cudaFuncSetAttribute(edge_feature64_tc_kernel<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc64Cfg<C>::template smem<0>()));
      GMP_CUDA(cudaFuncSetAttribute(edge_feature64_tc_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc64Cfg<C>::template smem<1>()));
      GMP_CUDA(cudaFuncSetAttribute(edge_feature64_tc_kernel<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc64Cfg<C>::template smem<2>()));
    }
    GMP_CUDA(cudaFuncSetAttribute(edge_msg_tc_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MsgTc<E>::kBytes));
    GMP_CUDA(cudaFuncSetAttribute(edge_msg_tc_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MsgTc<E>::kBytes));
    attr_done = true;
  }
  const float* W = m.d_weights;
  const int64_t table_stride = obs_tiles * 2 * E * OT;

  Timeline& tl = h->tl;
  tl.reset();
  // CSR by target -- on the handle's side stream, forked here and joined before the first consumer of the CSR (the edge-feature
  // stage): its 2 x E atomics are bound by L2 latency and overlap the node-side kernels below, which run at ~12 % occupancy
  if (!h->ex_side) {
    GMP_CUDA(cudaStreamCreateWithFlags(&h->ex_side, cudaStreamNonBlocking));
    GMP_CUDA(cudaEventCreateWithFlags(&h->ex_fork, cudaEventDisableTiming));
    GMP_CUDA(cudaEventCreateWithFlags(&h->ex_join, cudaEventDisableTiming));
  }
  const bool fork = getenv("GMP_NO_FORK") == nullptr;
  cudaStream_t cs = fork ? h->ex_side : st;
  if (fork) {
    GMP_CUDA(cudaEventRecord(h->ex_fork, st));
    GMP_CUDA(cudaStreamWaitEvent(cs, h->ex_fork, 0));
  }
  tl.begin(kPhCsr, cs);
  if (Nt > 0) GMP_CUDA(cudaMemsetAsync(ws.indeg, 0, (Nt + 1) * sizeof(int32_t), cs));
  const int csr_ctas = kNumSMs * (fork ? 8 : 16);           // forked: leave half of each SM's thread slots to the other stream
  if (Et > 0) {
    int gx = (int)std::min<int64_t>((Et + 255) / 256, csr_ctas);
    csr_count_kernel<<<gx, 256, 0, cs>>>(edge_index, row_stride, ws.edge_ptr, ws.node_ptr, (int)B, (int)Et, ws.indeg, ws.indeg + Nt);
    h->ex_bad_edges = ws.indeg + Nt;
    GMP_LAUNCH_CHECK();
  }
  csr_scan_kernel<<<(int)B, 256, 0, cs>>>(ws.indeg, ws.node_ptr, ws.edge_ptr, ws.in_ptr, ws.cursor);
  GMP_LAUNCH_CHECK();
  if (Et > 0) {
    int gx = (int)std::min<int64_t>((Et + 255) / 256, csr_ctas);
    csr_fill_kernel<<<gx, 256, 0, cs>>>(edge_index, row_stride, ws.edge_ptr, ws.node_ptr, (int)B, (int)Et, ws.in_ptr, ws.cursor,
                                        ws.csr_src, ws.csr_dst, ws.csr_eid);
    GMP_LAUNCH_CHECK();
  }
  tl.end(cs);
  if (fork) GMP_CUDA(cudaEventRecord(h->ex_join, cs));
  tl.begin(kPhGoal, st);
  goal_index_kernel<<<(int)B, 256, 0, st>>>(v, goal, C, ws.node_ptr, ws.goal_idx);
  GMP_LAUNCH_CHECK();
  tl.end(st);
  tl.begin(kPhObstacle, st);
  if (use_obstacles && obs_tiles > 0) {
    const int n_slots = (int)(obs_tiles * OT);
    obstacle_kernel<S, E><<<dim3((n_slots + R - 1) / R, 2), kRtThreads, smem, st>>>(m.w, W, obstacles, ws.obs_ptr, ws.obs_tile_ptr,
                                                                                  (int)B, n_slots, ws.tables, table_stride);
    GMP_LAUNCH_CHECK();
  }
  tl.end(st);
  tl.begin(kPhNodePre, st);
  if (tile_n[B] > 0) {
    node_pre_kernel<C, E><<<tile_n[B], kRtThreads, smem, st>>>(m.w, W, v, goal, ws.node_ptr, ws.tile_ptr_n, (int)B, ws.obs_ptr,
                                                              ws.obs_tile_ptr, ws.tables, table_stride, use_obstacles, ws.goal_idx,
                                                              ws.X0, ws.D0, ws.H);
    GMP_LAUNCH_CHECK();
  }
  tl.end(st);
  if (fork) GMP_CUDA(cudaStreamWaitEvent(st, h->ex_join, 0));   // join: everything below reads the CSR
  tl.begin(kPhEdgeFeature, st);
  bool tc_done = false;
  if constexpr (E == 32) {
    if (use_tc && tile_e[B] > 0) {
      // tcgen05 path: re-tile the edge stream's obstacle tables into hi / lo TF32 planes, then one persistent CTA per SM
      const int64_t tc_stride = tc_off[B];
      if (use_obstacles && tc_stride > 0) {
        tc_detail::obs_table_tc_kernel<<<dim3((unsigned)B, 3), 256, 0, st>>>(ws.tables, table_stride, ws.obs_ptr, ws.obs_tile_ptr,
                                                                            ws.tc_tab_off, ws.tc_tables, tc_stride);
        GMP_LAUNCH_CHECK();
      }
      static_assert(RowCfg<32>::R == 256, "a tensor-core unit is one 256-slot row tile");
      tc_detail::unit_meta_kernel<<<(tile_e[B] + 255) / 256, 256, 0, st>>>(ws.tile_ptr_e, (int)B, tile_e[B], ws.edge_ptr, ws.obs_ptr,
                                                                          ws.tc_tab_off, ws.tc_unit_meta);
      GMP_LAUNCH_CHECK();
      // auto: eight warps per tile where the first encoder layers are plain FMAs (2c <= 8: maze); four where they are MMA
      // stages of their own (wider inputs: two more round trips per tile, and the column split does not pay -- measured on
      // kuka14: 11.2 vs 11.8 ms)
      const int mode = m.edge_feature_mode;
      const bool four = mode == 2 || ((mode == -1 || mode == 3) && !TcCfg<C>::kSimtIn);
      // one issuer warp per tile (RD) needs exactly one table load per Block: 1 <= obstacles <= 128 in every graph with edges
      bool rd_ok = use_obstacles != 0;
      for (int64_t g = 0; g < B && rd_ok; ++g) {
        const int no = obs_ptr[g + 1] - obs_ptr[g];
        if (edge_ptr_h[g + 1] > edge_ptr_h[g] && (no < 1 || no > 128)) rd_ok = false;
      }
      // ... and is used with the eight-warps organisation only: with four warps per tile (wide inputs) it measured 19.1 ms
      // against 10.9 ms lockstep on C4 (profiles/r2_rd_issuer.md)
      const char* rd_env = getenv("GMP_TC_RD");                       // A/B switch: "0" lockstep issuer, anything else per-tile issuers
      const bool rd = rd_ok && !four && (rd_env ? rd_env[0] != '0' : (mode == 3 || mode == -1));
      const int grid = std::min<int>(tile_e[B], kNumSMs);
#define GMP_EF_LAUNCH(HALVES, RD, THREADS)                                                                                   \
  edge_feature_tc_kernel<C, HALVES, RD><<<grid, THREADS, TcCfg<C>::kSmemBytes, st>>>(                                       \
      W + m.w.tc_img, v, ws.csr_src, ws.csr_dst, ws.tc_unit_meta, tile_e[B], ws.tc_tables, tc_stride, use_obstacles, ws.P, ws.Q)
      if (four) {                      // four warps per tile, thread == row
        GMP_EF_LAUNCH(1, false, 384);
      } else {                         // eight warps per tile, columns split between warp pairs
        if (rd) GMP_EF_LAUNCH(2, true, 576); else GMP_EF_LAUNCH(2, false, 544);
      }
#undef GMP_EF_LAUNCH
      GMP_LAUNCH_CHECK();
      tc_done = true;
    }
  }
  if constexpr (E == 64) {
    if (use_tc64 && unit_ptr[B] > 0) {
      // tcgen05 path, embed 64: encoder -> three Block launches -> tail, the activation rows travelling through the Q buffer
      const int64_t tc_stride = tc_off[B];
      const int n_units = unit_ptr[B];
      if (use_obstacles && tc_stride > 0) {
        tc_detail::obs_table_tc64_kernel<<<dim3((unsigned)B, 3), 256, 0, st>>>(ws.tables, table_stride, ws.obs_ptr, ws.obs_tile_ptr,
                                                                              ws.tc_tab_off, ws.tc_tables, tc_stride);
        GMP_LAUNCH_CHECK();
      }
      tc_detail::unit_meta_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(ws.tc_unit_ptr, (int)B, n_units, ws.edge_ptr, ws.obs_ptr,
                                                                        ws.tc_tab_off, ws.tc_unit_meta);
      GMP_LAUNCH_CHECK();
      const int grid = std::min<int>(n_units, kNumSMs);
      edge_feature64_tc_kernel<C, 0><<<grid, 384, Tc64Cfg<C>::template smem<0>(), st>>>(W + m.w.tc64_img[0], v, ws.csr_src, ws.csr_dst,
                                                                                      ws.tc_unit_meta, n_units, nullptr, ws.Q, ws.P, ws.Q);
      GMP_LAUNCH_CHECK();
      for (int blk = 0; blk < (use_obstacles ? 3 : 0); ++blk) {
        edge_feature64_tc_kernel<C, 1><<<grid, 384, Tc64Cfg<C>::template smem<1>(), st>>>(
            W + m.w.tc64_img[1 + blk], v, ws.csr_src, ws.csr_dst, ws.tc_unit_meta, n_units, ws.tc_tables + (size_t)blk * tc_stride, ws.Q,
            ws.P, ws.Q);
        GMP_LAUNCH_CHECK();
      }
      edge_feature64_tc_kernel<C, 2><<<grid, 384, Tc64Cfg<C>::template smem<2>(), st>>>(W + m.w.tc64_img[4], v, ws.csr_src, ws.csr_dst,
                                                                                      ws.tc_unit_meta, n_units, nullptr, ws.Q, ws.P, ws.Q);
      GMP_LAUNCH_CHECK();
      tc_done = true;
    }
  }
  if (!tc_done && tile_e[B] > 0) {
    edge_feature_kernel<C, E><<<tile_e[B], kRtThreads, smem1, st>>>(m.w, W, v, ws.csr_src, ws.csr_dst, ws.edge_ptr, ws.tile_ptr_e,
                                                                  (int)B, ws.obs_ptr, ws.obs_tile_ptr, ws.tables, table_stride,
                                                                  use_obstacles, ws.P, ws.Q);
    GMP_LAUNCH_CHECK();
  }
  tl.end(st);
  const int node_tiles = (int)((Nt + R - 1) / R), slot_tiles = (int)((Et + R - 1) / R);
  int& msg_grid = h->ex_msg_grid;   // persistent grid of the message kernel: every resident CTA slot of the device
  if (msg_grid == 0) {
    int per_sm = 0;
    GMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_msg_kernel<E>, kRtThreads, MsgSmem<E>::kBytes));
    msg_grid = kNumSMs * std::max(per_sm, 1);
  }
  for (int it = 0; it <= loop; ++it) {
    const int mode = it == loop ? 2 : (it == 0 ? 0 : 1);
    if (node_tiles > 0) {
      tl.begin(kPhNodeLoop, st);
      node_loop_kernel<E><<<node_tiles, kRtThreads, smem, st>>>(m.w, W, mode, (int)Nt, ws.X0, ws.D0, ws.H, ws.Xg, ws.AGG, ws.A, ws.B);
      GMP_LAUNCH_CHECK();
      tl.end(st);
    }
    if (it < loop && slot_tiles > 0 && use_tc_msg) {
      tl.begin(kPhEdgeMsg, st);
      const int tiles128 = (int)((Et + MsgTc<E>::R - 1) / MsgTc<E>::R);
      edge_msg_tc_kernel<E, false><<<std::min(tiles128, kNumSMs * MsgTc<E>::kCtas), 128, MsgTc<E>::kBytes, st>>>(W + m.w.tc_l02, (int)Et, ws.csr_src, ws.csr_dst,
                                                                                        ws.A, ws.B, ws.P, ws.AGG, PolicyOut{});
      GMP_LAUNCH_CHECK();
      tl.end(st);
    } else if (it < loop && slot_tiles > 0) {
      tl.begin(kPhEdgeMsg, st);
      edge_msg_kernel<E><<<std::min(slot_tiles, msg_grid), kRtThreads, MsgSmem<E>::kBytes, st>>>(m.w, W, (int)Et, ws.csr_src, ws.csr_dst, ws.A, ws.B,
                                                                                                ws.P, ws.AGG);
      GMP_LAUNCH_CHECK();
      tl.end(st);
    }
  }
  tl.begin(kPhPolicy, st);
  if (dense && dense_off[B] > 0) GMP_CUDA(cudaMemsetAsync(dense, 0, dense_off[B] * sizeof(float), st));
  if (slot_tiles > 0 && use_tc_msg) {
    const int tiles128 = (int)((Et + MsgTc<E>::R - 1) / MsgTc<E>::R);
    edge_msg_tc_kernel<E, true><<<std::min(tiles128, kNumSMs * MsgTc<E>::kCtas), 128, MsgTc<E>::kBytes, st>>>(
        W + m.w.tc_p2, (int)Et, ws.csr_src, ws.csr_dst, ws.A, ws.B, ws.Q, nullptr,
        PolicyOut{ws.csr_eid, ws.edge_ptr, ws.node_ptr, ws.dense_off, (int)B, logits, dense});
    GMP_LAUNCH_CHECK();
  } else if (slot_tiles > 0) {
    policy_kernel<E><<<slot_tiles, kRtThreads, smem1, st>>>(m.w, W, (int)Et, ws.csr_src, ws.csr_dst, ws.csr_eid, ws.A, ws.B, ws.Q,
                                                          ws.edge_ptr, ws.node_ptr, ws.dense_off, (int)B, logits, dense);
    GMP_LAUNCH_CHECK();
  }
  tl.end(st);
  return GMP_OK;
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" int gmp_set_timing(gmp_handle* h, int enable) {
  GMP_REQUIRE(h, "null handle");
  h->tl.enabled = enable != 0;
  h->tl.reset();
  return GMP_OK;
}

extern "C" int gmp_get_timings(gmp_handle* h, float* ms_out_h, int n) {
  GMP_REQUIRE(h && ms_out_h && n >= kNumPhases, "need room for 8 phases");
  for (int i = 0; i < n; ++i) ms_out_h[i] = 0.f;
  for (auto& s : h->tl.spans) {
    if (!s.b) continue;
    GMP_CUDA(cudaEventSynchronize(s.b));
    float ms = 0.f;
    GMP_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    ms_out_h[s.phase] += ms;
  }
  return kNumPhases;
}

extern "C" int gmp_explorer_init(gmp_handle* h, int config_size, int embed_size, int obs_size) {
  GMP_REQUIRE(h, "null handle");
  GMP_REQUIRE(embed_size == 32 || embed_size == 64, "embed_size must be 32 or 64");
  GMP_REQUIRE(config_size >= 1 && config_size <= 14, "config_size must be in [1,14]");
  GMP_REQUIRE(obs_size == 2 || obs_size == 6, "obs_size must be 2 or 6");
  h->ex.c = config_size; h->ex.e = embed_size; h->ex.s = obs_size;
  h->ex.ready = false;
  h->ex_attr_done = false;   // another (c, e, s) means other kernel instantiations
  h->ex_msg_grid = 0;
  h->ex.tensors.clear();
  return GMP_OK;
}

extern "C" int gmp_explorer_set_tensor(gmp_handle* h, const char* name, const float* data_h, int64_t numel) {
  GMP_REQUIRE(h && name && (data_h || numel == 0) && numel >= 0, "null pointer");
  GMP_REQUIRE(h->ex.e != 0, "gmp_explorer_init first");
  h->ex.tensors[name] = std::vector<float>(data_h, data_h + numel);
  h->ex.ready = false;
  return GMP_OK;
}

extern "C" int gmp_explorer_bad_edges(gmp_handle* h, void* stream) {
  GMP_REQUIRE(h, "null handle");
  if (!h->ex_bad_edges) return 0;
  int32_t n = 0;
  GMP_CUDA(cudaSetDevice(h->device));
  GMP_CUDA(cudaMemcpyAsync(&n, h->ex_bad_edges, sizeof(n), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  GMP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return n;
}

extern "C" int gmp_explorer_set_edge_feature_mode(gmp_handle* h, int mode) {
  GMP_REQUIRE(h, "null handle");
  GMP_REQUIRE(mode >= -1 && mode <= 3, "mode: -1 auto, 0 fp32 SIMT, 1 tcgen05 3xTF32, 2 tcgen05 with the round-1 four-warp tiles, 3 tcgen05 with one issuer warp per tile");
  h->ex.edge_feature_mode = mode;
  return GMP_OK;
}

extern "C" int gmp_explorer_finalize(gmp_handle* h) {
  GMP_REQUIRE(h, "null handle");
  GMP_REQUIRE(h->ex.e != 0, "gmp_explorer_init first");
  GMP_CUDA(cudaSetDevice(h->device));
  return explorer_build_image(h->ex);
}

extern "C" int64_t gmp_explorer_workspace_bytes(const gmp_handle* h, int64_t n_graphs, int64_t n_nodes_total,
                                                int64_t n_edges_total, int64_t n_obs_total) {
  if (!h || h->ex.e == 0) return -1;
  const int ot = (h->ex.e == 32) ? 32 : 16;
  // every graph may add one partially filled obstacle tile
  const int64_t obs_tiles = n_obs_total / ot + n_graphs;
  // tensor-core table rows: every sub-chunk of <= 64 obstacles is padded to a multiple of 16 rows
  // (embed 64: one chunk of <= 32 rows per graph, 4*64 floats per row -- the same bound in 4*e-float rows)
  const int64_t tc_rows = n_obs_total + 16 * (n_obs_total / 48 + n_graphs);
  Carver cv(nullptr);
  ExWs ws;
  return carve_explorer(cv, ws, h->ex.e, n_graphs, n_nodes_total, n_edges_total, obs_tiles, tc_rows) + 256;
}

#define GMP_DISPATCH(CC, EE, SS)                                                                                              \
  if (c == CC && e == EE && s == SS)                                                                                          \
    return run_forward<CC, EE, SS>(h, n_graphs, v, edge_index, edge_row_stride, goal, obstacles, node_ptr_h, edge_ptr_h,      \
                                   obs_ptr_h, loop, use_obstacles, edge_logits_out, dense_out, workspace, workspace_bytes, st);

extern "C" int gmp_explorer_forward(gmp_handle* h, int64_t n_graphs, const float* v, const int64_t* edge_index,
                                    int64_t edge_row_stride, const float* goal, const float* obstacles, const int32_t* node_ptr_h,
                                    const int32_t* edge_ptr_h, const int32_t* obs_ptr_h, int loop, int use_obstacles,
                                    float* edge_logits_out, float* dense_out, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  GMP_REQUIRE(h, "null handle");
  if (!h->ex.ready) {
    set_error("gmp_explorer_forward: weights not loaded (gmp_explorer_set_tensor* + gmp_explorer_finalize)");
    return GMP_E_STATE;
  }
  GMP_REQUIRE(n_graphs >= 0 && loop >= 0, "negative size");
  if (n_graphs == 0) return GMP_OK;
  GMP_REQUIRE(node_ptr_h && edge_ptr_h, "null offset array");
  GMP_REQUIRE(!use_obstacles || obs_ptr_h, "obs_ptr_h required when use_obstacles");
  GMP_REQUIRE(node_ptr_h[0] == 0 && edge_ptr_h[0] == 0, "offset arrays must start at 0");
  const int64_t Et = edge_ptr_h[n_graphs];
  GMP_REQUIRE(v && goal && workspace && (edge_logits_out || Et == 0) && (edge_index || Et == 0), "null pointer");
  GMP_REQUIRE(edge_row_stride >= Et, "edge_row_stride < n_edges_total");
  GMP_REQUIRE(!use_obstacles || obstacles || obs_ptr_h[n_graphs] == 0, "null obstacles");
  GMP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c = h->ex.c, e = h->ex.e, s = h->ex.s;
  GMP_DISPATCH(2, 32, 2)    // maze2   (str2name.py:14)
  GMP_DISPATCH(3, 32, 2)    // maze3   (str2name.py:22)
  GMP_DISPATCH(7, 64, 6)    // kuka7   (str2name.py:30)
  GMP_DISPATCH(6, 32, 6)    // ur5     (str2name.py:38)
  GMP_DISPATCH(7, 32, 2)    // snake7  (str2name.py:46)
  GMP_DISPATCH(13, 32, 6)   // kuka13  (str2name.py:54)
  GMP_DISPATCH(14, 32, 6)   // kuka14  (str2name.py:62)
  set_error("gmp_explorer_forward: no kernel instantiated for (config_size, embed_size, obs_size) = (" + std::to_string(c) +
            ", " + std::to_string(e) + ", " + std::to_string(s) + "); supported: the str2name.py table");
  return GMP_E_UNSUPPORTED;
}
