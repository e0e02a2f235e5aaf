// GNN path-smoother forward for a packed batch of paths (sm_100a).
//
// Replaces ModelSmoother.forward, reference model_smoother.py:104-142, with its add-aggregating MPNN
// (model_smoother.py:22-39), per loop iteration:
//     e2 = knn(x=nodes[P:], y=path, k=10).flip(0); e2[0] += P          sample -> path edges        (:125-126)
//     E  = coalesce(cat(edge_index, e2))                                 sorted, unique              (:127-128)
//     x  = node_code([nodes | onehot(path, free, collided)])             Lin -> BN(eval) -> ReLU -> Lin (:130-136)
//     h  = x + lin_1( sum_{e: dst=i} lin_0([x_j - x_i, x_j, x_i]) )       (:138, :29-39)
//     path[1:-1] = smooth_node(h[:P])[1:-1]; nodes[:P] = path            (:139-140)
// Only path nodes have incoming edges (model_smoother.py:125-126, smoother.py:238-241), so h is needed on
// path rows only; x / A / B are needed on every node that is the SOURCE of an edge.
//
// Kernels per iteration:
//   smoother_graph_kernel   one CTA per problem: de-duplicates the caller's path-path edges in a shared-memory
//                           bit-matrix, selects the 10 nearest samples of every path node (canonical fp32 distance,
//                           bisection on the bit pattern, ties to the lower index) and writes the message list
//                           grouped by target, sources ascending (= coalesce order, so the fp32 sum order is fixed);
//   smoother_node_kernel    rows = all nodes: x (BatchNorm folded into the first Linear), A = (W1+W2) x,
//                           B = (W3-W1) x + b   (lin_0[0] split as in the explorer);
//   smoother_msg_kernel     rows = messages: m = lin_0[2](relu(A[src] + B[dst]));
//   smoother_path_kernel    rows = path nodes: agg = sum of the node's message segment, h = x + lin_1(agg),
//                           new = smooth_node(h), interior rows written back (times scale on the last iteration).
// Tiny and latency bound; it is batched over problems so that a launch fills the machine.
#include <cmath>

#include "handle.h"
#include "rowtile.cuh"

namespace gmp {
namespace {

constexpr int kE = 128;         // smoother embed size (str2name.py: embed_size=128 for every env)
constexpr int kKnn = 10;        // model_smoother.py:125
constexpr int kMaxPath = 512;   // path nodes per problem supported by the shared-memory bit-matrix
constexpr float kBnEps = 1e-5f;

struct SmSmem {
  using Cf = RowCfg<kE>;
  static constexpr int kBuf = kE * Cf::RP;
  static constexpr int kWB = kE * kE + 4 * kE;
  static constexpr size_t kBytes = (size_t)(2 * kBuf + kWB) * sizeof(float);
  float* X; float* S; float* WB;
  __device__ explicit SmSmem(float* base) { X = base; S = X + kBuf; WB = S + kBuf; }
};

template <int TM, int N>
__device__ __forceinline__ void add_vec(float (&acc)[TM][N], const float* __restrict__ vec) {
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r][n] += vec[n];
}

// ------------------------------------------------------------------------------------------------
// graph kernel: one CTA (256 threads) per problem; the edge set is a [P][n] bit-matrix in shared memory
// nodes layout per problem: [path (P) | free (F) | collided (C)], rows nd_ptr[g] .. nd_ptr[g+1]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smoother_graph_kernel(
    const float* __restrict__ nodes, int c, const int32_t* __restrict__ nd_ptr, const int32_t* __restrict__ path_len,
    const int64_t* __restrict__ edge_index, int64_t row_stride, const int32_t* __restrict__ edge_ptr,
    const int32_t* __restrict__ msg_ptr, const int32_t* __restrict__ prow_ptr, int strip, int32_t* __restrict__ msg_src,
    int32_t* __restrict__ msg_dst, int32_t* __restrict__ seg_ptr /* [P_total + B] (P_g + 1 per problem) */,
    int32_t* __restrict__ act_node /* [N_total]: per problem, the nodes that are a path node or the source of a message */,
    int32_t* __restrict__ act_cnt /* [B] */) {
  extern __shared__ __align__(16) unsigned char smem_u8[];
  const int g = blockIdx.x;
  const int n0 = nd_ptr[g], n = nd_ptr[g + 1] - n0;
  const int P = path_len[g];
  const int S = n - P;                       // samples
  const int wpr = (n + 31) >> 5;
  uint32_t* bm = reinterpret_cast<uint32_t*>(smem_u8);                 // [P][wpr] : bm[dst][src], src over ALL nodes
  int* cnt = reinterpret_cast<int*>(bm + (size_t)P * wpr);             // [P + 1]
  float* dist = reinterpret_cast<float*>(cnt + P + 1);                 // [8][strip]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < P * wpr; i += 256) bm[i] = 0u;
  __syncthreads();
  // caller's edges (model_smoother.py:127).  Only path nodes are read back (h[:P], :139), so edges into
  // sample nodes cannot influence the result and are dropped; sources may be any node.
  for (int e = edge_ptr[g] + threadIdx.x; e < edge_ptr[g + 1]; e += 256) {
    const int s = (int)edge_index[e], d = (int)edge_index[row_stride + e];
    if (s >= 0 && s < n && d >= 0 && d < P) atomicOr(bm + (size_t)d * wpr + (s >> 5), 1u << (s & 31));
  }
  __syncthreads();
  // k nearest samples of every path node -> bits (sample -> path edges, model_smoother.py:125-126)
  const int k = min(kKnn, S);
  for (int i = warp; i < P; i += 8) {
    float* dw = dist + (size_t)warp * strip;
    for (int j = lane; j < S; j += 32) {
      float d = 0.0f;
      for (int q = 0; q < c; ++q) {
        const float diff = __fsub_rn(nodes[(size_t)(n0 + i) * c + q], nodes[(size_t)(n0 + P + j) * c + q]);
        d = __fadd_rn(d, __fmul_rn(diff, diff));
      }
      dw[j] = d;
    }
    __syncwarp();
    uint32_t T = 0xffffffffu;
    int n_less = S;
    if (k < S) {
      T = 0;
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int cl = 0;
        for (int j = lane; j < S; j += 32) cl += (__float_as_uint(dw[j]) < cand) ? 1 : 0;
        cl = __reduce_add_sync(0xffffffffu, cl);
        if (cl < k) T = cand;
      }
      int cl = 0;
      for (int j = lane; j < S; j += 32) cl += (__float_as_uint(dw[j]) < T) ? 1 : 0;
      n_less = __reduce_add_sync(0xffffffffu, cl);
    }
    int quota = k - min(n_less, k);
    for (int base = 0; base < S; base += 32) {
      const int j = base + lane;
      bool less = false, tie = false;
      if (j < S) {
        const uint32_t b = __float_as_uint(dw[j]);
        less = b < T;
        tie = (b == T) && (k < S);
      }
      const uint32_t tb = __ballot_sync(0xffffffffu, tie);
      const bool take = less || (tie && __popc(tb & ((1u << lane) - 1u)) < quota);
      quota -= min(quota, __popc(tb));
      if (take) atomicOr(bm + (size_t)i * wpr + ((P + j) >> 5), 1u << ((P + j) & 31));
    }
    __syncwarp();
  }
  __syncthreads();
  // segment sizes and offsets
  for (int i = threadIdx.x; i < P; i += 256) {
    int cdeg = 0;
    for (int w = 0; w < wpr; ++w) cdeg += __popc(bm[(size_t)i * wpr + w]);
    cnt[i] = cdeg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // P <= 512: a serial scan is fine here
    int run = 0;
    for (int i = 0; i < P; ++i) { const int x = cnt[i]; cnt[i] = run; run += x; }
    cnt[P] = run;
  }
  __syncthreads();
  // ACTIVE nodes: x / A are read for message SOURCES only and h for path rows only (model_smoother.py:125-126,139), so the node
  // MLP (three 128x128 products per row) runs on path nodes + the union of the rows' source sets -- at most P + 10 P + caller
  // sources of ~1 000 nodes per problem.  Ascending node order: the path nodes come first.
  if (threadIdx.x < 32) {
    int base = 0;
    for (int wb = 0; wb < wpr; wb += 32) {
      const int w = wb + lane;
      uint32_t bits = 0u;
      if (w < wpr) {
        for (int i = 0; i < P; ++i) bits |= bm[(size_t)i * wpr + w];
        const int lo = w << 5;                       // path nodes [0, P) are always active
        if (lo + 32 <= P) bits = 0xffffffffu;
        else if (lo < P) bits |= (1u << (P - lo)) - 1u;
        if (lo + 32 > n) bits &= (n - lo >= 32) ? 0xffffffffu : ((1u << (n - lo)) - 1u);
      }
      const int pc = __popc(bits);
      int incl = pc;
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
      int p = n0 + base + incl - pc;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        act_node[p++] = n0 + (w << 5) + b;
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) act_cnt[g] = base;
  }
  const int m0 = msg_ptr[g];
  int32_t* seg = seg_ptr + prow_ptr[g] + g;
  for (int i = threadIdx.x; i <= P; i += 256) seg[i] = m0 + cnt[i];
  for (int i = threadIdx.x; i < P; i += 256) {
    int p = m0 + cnt[i];
    for (int w = 0; w < wpr; ++w) {          // sources ascending = coalesce order
      uint32_t bits = bm[(size_t)i * wpr + w];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        msg_src[p] = n0 + (w << 5) + b;
        msg_dst[p] = n0 + i;
        ++p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// rows = all nodes of the batch: x, A, B
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kRtThreads) smoother_node_kernel(SmootherW w, const float* __restrict__ W,
                                                                   const float* __restrict__ nodes, const int32_t* __restrict__ nd_ptr,
                                                                   const int32_t* __restrict__ path_len, const int32_t* __restrict__ free_len,
                                                                   const int32_t* __restrict__ act_node, const int32_t* __restrict__ act_cnt,
                                                                   float* __restrict__ Xg, float* __restrict__ A, float* __restrict__ B) {
  using Cf = RowCfg<kE>;
  constexpr int TM = Cf::TM, R = Cf::R, RP = Cf::RP;
  static_assert(TM == 1, "smoother tiles use one row per thread");
  extern __shared__ __align__(16) float smem_raw[];
  SmSmem sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  // tile blockIdx.x of the ACTIVE nodes of problem blockIdx.y (smoother_graph_kernel); results land at the node's own row
  const int g = blockIdx.y;
  const int n_act = act_cnt[g];
  if (blockIdx.x * R >= n_act) return;
  const int arow = blockIdx.x * R + threadIdx.x;
  const bool valid = arow < n_act;
  const int row = valid ? act_node[nd_ptr[g] + arow] : 0;
  float in[1][C + 3];
#pragma unroll
  for (int k = 0; k < C + 3; ++k) in[0][k] = 0.0f;
  bool is_path = false;
  if (valid) {
    const int local = row - nd_ptr[g];
    is_path = local < path_len[g];
#pragma unroll
    for (int k = 0; k < C; ++k) in[0][k] = nodes[(size_t)row * C + k];
    const int kind = local < path_len[g] ? 0 : (local < path_len[g] + free_len[g] ? 1 : 2);   // model_smoother.py:130-133
    in[0][C] = kind == 0 ? 1.0f : 0.0f;
    in[0][C + 1] = kind == 1 ? 1.0f : 0.0f;
    in[0][C + 2] = kind == 2 ? 1.0f : 0.0f;
  }
  float acc[1][kE];
  // node_code.0 + BatchNorm(eval) folded, ReLU                      (model_smoother.py:65,135-136)
  stage_load(sm.WB, W + w.nc0, (C + 3) * kE + kE);
  acc_zero(acc);
  gemm_reg<C + 3, kE, 1>(acc, in, sm.WB);
  add_vec(acc, sm.WB + (C + 3) * kE);
  acc_relu(acc);
  acc_store<1, kE, RP>(acc, xcol);
  stage_load(sm.WB, W + w.nc3, kE * kE + kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, xcol, sm.WB);
  add_vec(acc, sm.WB + kE * kE);
  acc_store<1, kE, RP>(acc, xcol);
  if (valid) acc_store_global<1, kE>(acc, 0, Xg + (size_t)row * kE);
  stage_load(sm.WB, W + w.l0_A, kE * kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, xcol, sm.WB);
  if (valid) acc_store_global<1, kE>(acc, 0, A + (size_t)row * kE);
  // B = (W3-W1) x + b is read for message TARGETS only, and only path nodes are targets (model_smoother.py:125-126: the k-NN
  // edges run sample -> path): tiles without a path row (7 of 8 at 1000 samples per problem) skip the product
  if (!__syncthreads_or(is_path ? 1 : 0)) return;
  stage_load(sm.WB, W + w.l0_B, kE * kE + kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, xcol, sm.WB);
  add_vec(acc, sm.WB + kE * kE);
  if (valid) acc_store_global<1, kE>(acc, 0, B + (size_t)row * kE);
}

// rows = messages: m = lin_0[2](relu(A[src] + B[dst]))              (model_smoother.py:36-39)
__global__ void __launch_bounds__(kRtThreads) smoother_msg_kernel(SmootherW w, const float* __restrict__ W, int n_msgs,
                                                                  const int32_t* __restrict__ msg_src, const int32_t* __restrict__ msg_dst,
                                                                  const float* __restrict__ A, const float* __restrict__ B,
                                                                  float* __restrict__ M) {
  using Cf = RowCfg<kE>;
  constexpr int R = Cf::R, RP = Cf::RP;
  extern __shared__ __align__(16) float smem_raw[];
  SmSmem sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  const int row = blockIdx.x * R + threadIdx.x;
  const int s = row < n_msgs ? msg_src[row] : -1;
  const bool valid = s >= 0;
  if (valid) {
    const int d = msg_dst[row];
    const float4* a4 = reinterpret_cast<const float4*>(A + (size_t)s * kE);
    const float4* b4 = reinterpret_cast<const float4*>(B + (size_t)d * kE);
#pragma unroll 8
    for (int n = 0; n < kE / 4; ++n) {
      const float4 a = __ldg(a4 + n), b = __ldg(b4 + n);
      xcol[(4 * n) * RP] = fmaxf(a.x + b.x, 0.0f);
      xcol[(4 * n + 1) * RP] = fmaxf(a.y + b.y, 0.0f);
      xcol[(4 * n + 2) * RP] = fmaxf(a.z + b.z, 0.0f);
      xcol[(4 * n + 3) * RP] = fmaxf(a.w + b.w, 0.0f);
    }
  } else {
    for (int n = 0; n < kE; ++n) xcol[n * RP] = 0.0f;
  }
  float acc[1][kE];
  stage_load(sm.WB, W + w.l0_2, kE * kE + kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, xcol, sm.WB);
  add_vec(acc, sm.WB + kE * kE);
  if (valid) acc_store_global<1, kE>(acc, 0, M + (size_t)row * kE);
}

// rows = path nodes: aggregate, lin_1, residual, smooth_node, write back   (model_smoother.py:33-35,138-140)
template <int C>
__global__ void __launch_bounds__(kRtThreads) smoother_path_kernel(SmootherW w, const float* __restrict__ W,
                                                                   const int32_t* __restrict__ prow_ptr, const int32_t* __restrict__ nd_ptr,
                                                                   int n_graphs, int n_prows, const int32_t* __restrict__ seg_ptr,
                                                                   const float* __restrict__ M, const float* __restrict__ Xg,
                                                                   float* __restrict__ nodes, float out_scale, float* __restrict__ path_out) {
  using Cf = RowCfg<kE>;
  constexpr int R = Cf::R, RP = Cf::RP;
  constexpr int CP = (C + 3) / 4 * 4;
  extern __shared__ __align__(16) float smem_raw[];
  SmSmem sm(smem_raw);
  float* xcol = sm.X + threadIdx.x;
  float* scol = sm.S + threadIdx.x;
  const int prow = blockIdx.x * R + threadIdx.x;
  const bool valid = prow < n_prows;
  int g = 0, local = 0, P = 0, node = 0;
  float acc[1][kE];
  acc_zero(acc);
  if (valid) {
    g = find_segment(prow_ptr, n_graphs, prow);
    local = prow - prow_ptr[g];
    P = prow_ptr[g + 1] - prow_ptr[g];
    node = nd_ptr[g] + local;
    const int m0 = seg_ptr[prow + g], m1 = seg_ptr[prow + g + 1];
    for (int m = m0; m < m1; ++m) {           // fixed order: sources ascending (index_add_ over coalesced edges)
      const float4* r4 = reinterpret_cast<const float4*>(M + (size_t)m * kE);
#pragma unroll 8
      for (int n = 0; n < kE / 4; ++n) {
        const float4 x = __ldg(r4 + n);
        acc[0][4 * n] += x.x; acc[0][4 * n + 1] += x.y; acc[0][4 * n + 2] += x.z; acc[0][4 * n + 3] += x.w;
      }
    }
  }
  acc_store<1, kE, RP>(acc, xcol);
  // lin_1 = Seq(Lin, ReLU, Lin)                                      (model_smoother.py:27,35)
  stage_load(sm.WB, W + w.l1_0, kE * kE + kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, xcol, sm.WB);
  add_vec(acc, sm.WB + kE * kE);
  acc_relu(acc);
  acc_store<1, kE, RP>(acc, scol);
  stage_load(sm.WB, W + w.l1_2, kE * kE + kE);
  acc_zero(acc);
  gemm_smem<kE, kE, 1, RP>(acc, scol, sm.WB);
  add_vec(acc, sm.WB + kE * kE);
  if (valid) {
    const float4* x4 = reinterpret_cast<const float4*>(Xg + (size_t)node * kE);
#pragma unroll 8
    for (int n = 0; n < kE / 4; ++n) {
      const float4 x = __ldg(x4 + n);
      acc[0][4 * n] += x.x; acc[0][4 * n + 1] += x.y; acc[0][4 * n + 2] += x.z; acc[0][4 * n + 3] += x.w;
    }
  }
  acc_store<1, kE, RP>(acc, xcol);  // h
  // smooth_node: Lin(128 -> c), columns padded to CP            (model_smoother.py:102,139)
  stage_load(sm.WB, W + w.sm, kE * CP + CP);
  float o[1][CP];
  acc_zero(o);
  gemm_smem<kE, CP, 1, RP>(o, xcol, sm.WB);
  add_vec(o, sm.WB + kE * CP);
  if (valid) {
    const bool interior = local >= 1 && local < P - 1;     // path[1:-1] only (model_smoother.py:139)
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float val = interior ? o[0][k] : nodes[(size_t)node * C + k];
      if (interior) nodes[(size_t)node * C + k] = val;    // nodes[:P] = path (model_smoother.py:140)
      if (path_out) path_out[(size_t)prow * C + k] = val * out_scale;
    }
  }
}

// nodes = cat(path, free, collided) / scale                           (model_smoother.py:118-121)
__global__ void __launch_bounds__(256) smoother_pack_kernel(const float* __restrict__ path, const float* __restrict__ samples, int c,
                                                            const int32_t* __restrict__ nd_ptr, const int32_t* __restrict__ prow_ptr,
                                                            const int32_t* __restrict__ srow_ptr, int n_graphs, int n_rows, float scale,
                                                            float* __restrict__ nodes) {
  for (int row = blockIdx.x * 256 + threadIdx.x; row < n_rows; row += gridDim.x * 256) {
    const int g = find_segment(nd_ptr, n_graphs, row);
    const int local = row - nd_ptr[g];
    const int P = prow_ptr[g + 1] - prow_ptr[g];
    const float* src = local < P ? path + (size_t)(prow_ptr[g] + local) * c : samples + (size_t)(srow_ptr[g] + local - P) * c;
    for (int k = 0; k < c; ++k) nodes[(size_t)row * c + k] = src[k] / scale;
  }
}

struct SmWs {
  int32_t *nd_ptr, *prow_ptr, *srow_ptr, *path_len, *free_len, *edge_ptr, *msg_ptr, *msg_src, *msg_dst, *seg_ptr, *act_node, *act_cnt;
  float *nodes, *Xg, *A, *B, *M;
};

int64_t carve_smoother(Carver& cv, SmWs& ws, int c, int64_t B, int64_t Nt, int64_t Pt, int64_t Mcap) {
  ws.nd_ptr = cv.take<int32_t>(B + 1);
  ws.prow_ptr = cv.take<int32_t>(B + 1);
  ws.srow_ptr = cv.take<int32_t>(B + 1);
  ws.path_len = cv.take<int32_t>(B);
  ws.free_len = cv.take<int32_t>(B);
  ws.edge_ptr = cv.take<int32_t>(B + 1);
  ws.msg_ptr = cv.take<int32_t>(B + 1);
  ws.msg_src = cv.take<int32_t>(Mcap);
  ws.msg_dst = cv.take<int32_t>(Mcap);
  ws.seg_ptr = cv.take<int32_t>(Pt + B + 1);
  ws.act_node = cv.take<int32_t>(Nt);
  ws.act_cnt = cv.take<int32_t>(B);
  ws.nodes = cv.take<float>(Nt * c);
  ws.Xg = cv.take<float>(Nt * kE);
  ws.A = cv.take<float>(Nt * kE);
  ws.B = cv.take<float>(Nt * kE);
  ws.M = cv.take<float>(Mcap * kE);
  return cv.bytes();
}

template <int C>
int run_smoother(gmp_handle* h, int64_t B, const float* path, const float* samples, const int64_t* edge_index, int64_t row_stride,
                 const int32_t* path_ptr_h, const int32_t* sample_ptr_h, const int32_t* n_free_h, const int32_t* edge_ptr_h,
                 float scale, int loop, float* path_out, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  using Cf = RowCfg<kE>;
  const SmootherModel& m = h->sm;
  const int64_t Pt = path_ptr_h[B], St = sample_ptr_h[B], Nt = Pt + St, Et = edge_ptr_h[B];
  std::vector<int32_t> meta((size_t)(B + 1) * 5 + 2 * B);
  int32_t* nd = meta.data();
  int32_t* mp = nd + (B + 1);
  int32_t* pl = mp + (B + 1);
  int32_t* fl = pl + B;
  nd[0] = mp[0] = 0;
  int max_p = 0, max_s = 0, max_act = 1;
  for (int64_t g = 0; g < B; ++g) {
    const int P = path_ptr_h[g + 1] - path_ptr_h[g], S = sample_ptr_h[g + 1] - sample_ptr_h[g];
    const int ne = edge_ptr_h[g + 1] - edge_ptr_h[g];
    GMP_REQUIRE(P >= 0 && S >= 0 && ne >= 0 && n_free_h[g] >= 0 && n_free_h[g] <= S, "bad offsets / n_free");
    GMP_REQUIRE(P <= kMaxPath, "paths longer than 512 nodes are not supported");
    nd[g + 1] = nd[g] + P + S;
    mp[g + 1] = mp[g] + ne + kKnn * P;
    pl[g] = P;
    fl[g] = n_free_h[g];
    max_p = std::max(max_p, P);
    max_s = std::max(max_s, S);
    max_act = std::max(max_act, std::min(P + S, P + kKnn * P + ne));   // path nodes + at most one new source per message
  }
  const int64_t Mcap = mp[B];
  GMP_REQUIRE(B <= 65535, "more than 65535 problems per call (grid.y limit)");
  Carver cv(workspace);
  SmWs ws;
  GMP_REQUIRE(carve_smoother(cv, ws, C, B, Nt, Pt, Mcap) <= workspace_bytes, "workspace too small (see gmp_smoother_workspace_bytes)");
  GMP_CUDA(cudaMemcpyAsync(ws.nd_ptr, nd, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.msg_ptr, mp, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.path_len, pl, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.free_len, fl, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.prow_ptr, path_ptr_h, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.srow_ptr, sample_ptr_h, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GMP_CUDA(cudaMemcpyAsync(ws.edge_ptr, edge_ptr_h, (B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));

  const size_t smem = SmSmem::kBytes;
  GMP_CUDA(cudaFuncSetAttribute(smoother_node_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GMP_CUDA(cudaFuncSetAttribute(smoother_msg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GMP_CUDA(cudaFuncSetAttribute(smoother_path_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int strip = (int)align_up(std::max(max_s, 1), 32);
  const int wpr = (max_p + max_s + 31) / 32;
  const size_t gsmem = (size_t)max_p * wpr * 4 + (size_t)(max_p + 1) * 4 + (size_t)8 * strip * 4 + 64;
  GMP_REQUIRE(gsmem <= 200 * 1024, "problem too large for the smoother graph kernel (samples per problem)");
  GMP_CUDA(cudaFuncSetAttribute(smoother_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));

  const float* W = m.d_weights;
  if (Nt > 0) {
    int gx = (int)std::min<int64_t>((Nt + 255) / 256, kNumSMs * 8);
    smoother_pack_kernel<<<gx, 256, 0, st>>>(path, samples, C, ws.nd_ptr, ws.prow_ptr, ws.srow_ptr, (int)B, (int)Nt, scale, ws.nodes);
    GMP_LAUNCH_CHECK();
  }
  const int R = Cf::R;
  for (int it = 0; it < loop; ++it) {
    if (Mcap > 0) GMP_CUDA(cudaMemsetAsync(ws.msg_src, 0xff, Mcap * sizeof(int32_t), st));
    smoother_graph_kernel<<<(int)B, 256, gsmem, st>>>(ws.nodes, C, ws.nd_ptr, ws.path_len, edge_index, row_stride, ws.edge_ptr,
                                                      ws.msg_ptr, ws.prow_ptr, strip, ws.msg_src, ws.msg_dst, ws.seg_ptr, ws.act_node,
                                                      ws.act_cnt);
    GMP_LAUNCH_CHECK();
    if (Nt > 0) {
      smoother_node_kernel<C><<<dim3((unsigned)((max_act + R - 1) / R), (unsigned)B), kRtThreads, smem, st>>>(
          m.w, W, ws.nodes, ws.nd_ptr, ws.path_len, ws.free_len, ws.act_node, ws.act_cnt, ws.Xg, ws.A, ws.B);
      GMP_LAUNCH_CHECK();
    }
    if (Mcap > 0) {
      smoother_msg_kernel<<<(int)((Mcap + R - 1) / R), kRtThreads, smem, st>>>(m.w, W, (int)Mcap, ws.msg_src, ws.msg_dst, ws.A, ws.B, ws.M);
      GMP_LAUNCH_CHECK();
    }
    if (Pt > 0) {
      const bool last = it == loop - 1;
      smoother_path_kernel<C><<<(int)((Pt + R - 1) / R), kRtThreads, smem, st>>>(m.w, W, ws.prow_ptr, ws.nd_ptr, (int)B, (int)Pt, ws.seg_ptr,
                                                                                 ws.M, ws.Xg, ws.nodes, scale, last ? path_out : nullptr);
      GMP_LAUNCH_CHECK();
    }
  }
  if (loop == 0 && Pt > 0) GMP_CUDA(cudaMemcpyAsync(path_out, path, Pt * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  (void)Et;
  return GMP_OK;
}

}  // namespace

int smoother_build_image(SmootherModel& m) {
  const int c = m.c, e = m.e;
  auto T = [&](const std::string& name) -> const std::vector<float>& { return m.tensors.at(name); };
  std::vector<std::pair<std::string, int64_t>> need = {
      {"node_code.0.weight", (int64_t)e * (c + 3)}, {"node_code.0.bias", e},
      {"node_code.1.weight", e}, {"node_code.1.bias", e}, {"node_code.1.running_mean", e}, {"node_code.1.running_var", e},
      {"node_code.3.weight", (int64_t)e * e}, {"node_code.3.bias", e},
      {"process.lin_0.0.weight", (int64_t)e * 3 * e}, {"process.lin_0.0.bias", e},
      {"process.lin_0.2.weight", (int64_t)e * e}, {"process.lin_0.2.bias", e},
      {"process.lin_1.0.weight", (int64_t)e * e}, {"process.lin_1.0.bias", e},
      {"process.lin_1.2.weight", (int64_t)e * e}, {"process.lin_1.2.bias", e},
      {"smooth_node.weight", (int64_t)c * e}, {"smooth_node.bias", c},
  };
  for (auto& kv : need) {
    auto it = m.tensors.find(kv.first);
    if (it == m.tensors.end()) {
      set_error("smoother weights: missing tensor '" + kv.first + "'");
      return GMP_E_STATE;
    }
    if ((int64_t)it->second.size() != kv.second) {
      set_error("smoother weights: tensor '" + kv.first + "' has the wrong number of elements");
      return GMP_E_INVALID;
    }
  }
  std::vector<float> buf;
  auto begin = [&]() { while (buf.size() % 4) buf.push_back(0.f); return (int)buf.size(); };
  auto put_t = [&](const std::vector<double>& Wd, int out, int in) {  // [out][in] -> K-major [in][out]
    for (int k = 0; k < in; ++k)
      for (int n = 0; n < out; ++n) buf.push_back((float)Wd[(size_t)n * in + k]);
  };
  auto dbl = [&](const std::vector<float>& x) { return std::vector<double>(x.begin(), x.end()); };
  SmootherW& w = m.w;
  {  // node_code.0 with BatchNorm1d (eval) folded: y = (Wx + b - mu) / sqrt(var + eps) * g + beta   (model_smoother.py:63-65)
    const auto &W0 = T("node_code.0.weight"), &b0 = T("node_code.0.bias"), &g = T("node_code.1.weight"), &be = T("node_code.1.bias"),
               &mu = T("node_code.1.running_mean"), &var = T("node_code.1.running_var");
    const int in = c + 3;
    std::vector<double> Wf((size_t)e * in), bf(e);
    for (int n = 0; n < e; ++n) {
      const double sc = (double)g[n] / std::sqrt((double)var[n] + (double)kBnEps);
      for (int k = 0; k < in; ++k) Wf[(size_t)n * in + k] = (double)W0[(size_t)n * in + k] * sc;
      bf[n] = ((double)b0[n] - (double)mu[n]) * sc + (double)be[n];
    }
    w.nc0 = begin();
    put_t(Wf, e, in);
    for (double d : bf) buf.push_back((float)d);
  }
  auto lin = [&](const std::string& name) {
    int off = begin();
    put_t(dbl(T(name + ".weight")), e, e);
    for (float f : T(name + ".bias")) buf.push_back(f);
    return off;
  };
  w.nc3 = lin("node_code.3");
  {
    const auto& W0 = T("process.lin_0.0.weight");  // [e][3e]: x_j - x_i | x_j | x_i   (model_smoother.py:37)
    std::vector<double> Am((size_t)e * e), Bm((size_t)e * e);
    for (int n = 0; n < e; ++n)
      for (int k = 0; k < e; ++k) {
        const double w1 = W0[(size_t)n * 3 * e + k], w2 = W0[(size_t)n * 3 * e + e + k], w3 = W0[(size_t)n * 3 * e + 2 * e + k];
        Am[(size_t)n * e + k] = w1 + w2;
        Bm[(size_t)n * e + k] = w3 - w1;
      }
    w.l0_A = begin();
    put_t(Am, e, e);
    w.l0_B = begin();
    put_t(Bm, e, e);
    for (float f : T("process.lin_0.0.bias")) buf.push_back(f);
  }
  w.l0_2 = lin("process.lin_0.2");
  w.l1_0 = lin("process.lin_1.0");
  w.l1_2 = lin("process.lin_1.2");
  {
    const int cp = (c + 3) / 4 * 4;
    const auto &Ws = T("smooth_node.weight"), &bs = T("smooth_node.bias");  // [c][e]
    w.sm = begin();
    for (int k = 0; k < e; ++k)
      for (int n = 0; n < cp; ++n) buf.push_back(n < c ? Ws[(size_t)n * e + k] : 0.f);
    for (int n = 0; n < cp; ++n) buf.push_back(n < c ? bs[n] : 0.f);
  }
  begin();
  for (int q = 0; q < 1024; ++q) buf.push_back(0.f);
  if (m.d_weights) cudaFree(m.d_weights);
  m.d_weights = nullptr;
  GMP_CUDA(cudaMalloc(&m.d_weights, buf.size() * sizeof(float)));
  GMP_CUDA(cudaMemcpy(m.d_weights, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice));
  m.n_weights = (int64_t)buf.size();
  m.ready = true;
  return GMP_OK;
}

}  // namespace gmp

using namespace gmp;

extern "C" int gmp_smoother_init(gmp_handle* h, int config_size, int embed_size) {
  GMP_REQUIRE(h, "null handle");
  GMP_REQUIRE(embed_size == kE, "smoother embed_size must be 128 (str2name.py)");
  GMP_REQUIRE(config_size >= 1 && config_size <= 14, "config_size must be in [1,14]");
  h->sm.c = config_size;
  h->sm.e = embed_size;
  h->sm.ready = false;
  h->sm.tensors.clear();
  return GMP_OK;
}

extern "C" int gmp_smoother_set_tensor(gmp_handle* h, const char* name, const float* data_h, int64_t numel) {
  GMP_REQUIRE(h && name && (data_h || numel == 0) && numel >= 0, "null pointer");
  GMP_REQUIRE(h->sm.e != 0, "gmp_smoother_init first");
  h->sm.tensors[name] = std::vector<float>(data_h, data_h + numel);
  h->sm.ready = false;
  return GMP_OK;
}

extern "C" int gmp_smoother_finalize(gmp_handle* h) {
  GMP_REQUIRE(h, "null handle");
  GMP_REQUIRE(h->sm.e != 0, "gmp_smoother_init first");
  GMP_CUDA(cudaSetDevice(h->device));
  return smoother_build_image(h->sm);
}

extern "C" int64_t gmp_smoother_workspace_bytes(const gmp_handle* h, int64_t n_problems, int64_t n_path_total, int64_t n_sample_total,
                                                int64_t n_edges_total) {
  if (!h || h->sm.e == 0) return -1;
  Carver cv(nullptr);
  SmWs ws;
  return carve_smoother(cv, ws, h->sm.c, n_problems, n_path_total + n_sample_total, n_path_total,
                        n_edges_total + kKnn * n_path_total) + 256;
}

#define GMP_SM_DISPATCH(CC)                                                                                                   \
  if (c == CC)                                                                                                                \
    return run_smoother<CC>(h, n_problems, path, samples, edge_index, edge_row_stride, path_ptr_h, sample_ptr_h, n_free_h,    \
                            edge_ptr_h, scale, loop, path_out, workspace, workspace_bytes, st);

extern "C" int gmp_smoother_forward(gmp_handle* h, int64_t n_problems, const float* path, const float* samples,
                                    const int64_t* edge_index, int64_t edge_row_stride, const int32_t* path_ptr_h,
                                    const int32_t* sample_ptr_h, const int32_t* n_free_h, const int32_t* edge_ptr_h, float scale,
                                    int loop, float* path_out, void* workspace, int64_t workspace_bytes, void* stream) {
  GMP_REQUIRE(h, "null handle");
  if (!h->sm.ready) {
    set_error("gmp_smoother_forward: weights not loaded (gmp_smoother_set_tensor* + gmp_smoother_finalize)");
    return GMP_E_STATE;
  }
  GMP_REQUIRE(n_problems >= 0 && loop >= 0, "negative size");
  if (n_problems == 0) return GMP_OK;
  GMP_REQUIRE(path_ptr_h && sample_ptr_h && n_free_h && edge_ptr_h, "null offset array");
  GMP_REQUIRE(path && path_out && workspace, "null pointer");
  GMP_REQUIRE(samples || sample_ptr_h[n_problems] == 0, "null samples");
  GMP_REQUIRE(edge_index || edge_ptr_h[n_problems] == 0, "null edge_index");
  GMP_REQUIRE(scale != 0.0f, "scale must be non-zero");
  GMP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c = h->sm.c;
  GMP_SM_DISPATCH(2) GMP_SM_DISPATCH(3) GMP_SM_DISPATCH(6) GMP_SM_DISPATCH(7) GMP_SM_DISPATCH(13) GMP_SM_DISPATCH(14)
  set_error("gmp_smoother_forward: no kernel instantiated for config_size " + std::to_string(c));
  return GMP_E_UNSUPPORTED;
}
