// Row-tile machinery shared by the explorer and smoother kernels (sm_100a, fp32 SIMT).
//
// Every dense op on the hot path is "rows x small square weight": M in {N, E, O} rows (nodes, edges,
// obstacle tokens) times a K x N weight with K, N in {c..4c, e, 2e}, e in {32, 64, 128}.  The logit
// tolerance (1e-4 absolute on O(10) logits after ~40 chained layers) rules out single-pass
// TF32/BF16 tensor-core math, so these run as fp32 FMAs in the layout that keeps the FMA pipe fed:
//
//   * a CTA owns a tile of R = THREADS*TM rows; thread t owns rows {t, t+THREADS, ...} (TM of them);
//   * activations live in shared memory FEATURE-MAJOR: buf[k * RP + row], RP = R + 1.  A thread only
//     ever touches its own rows' columns, so activation traffic needs no barrier and is conflict-free
//     (consecutive lanes -> consecutive banks);
//   * weights are staged per stage into shared memory K-major (Wt[k][n], pre-transposed at load time),
//     so the inner product is an outer-product update: one scalar LDS per row + N/4 broadcast
//     LDS.128 feed TM*N FMAs per k -- (TM + N/4) LSU wavefronts per TM*N FMA instructions keeps the
//     single LSU port below the 4 FMA-issuing sub-partitions for TM*N = 64;
//   * accumulators (TM x N = 64 floats) stay in registers; bias / ReLU / residual / LayerNorm /
//     online-softmax epilogues are therefore thread-local -- no cross-thread reductions anywhere
//     except the aggregation itself.
#pragma once
#include "common.cuh"

namespace gmp {

constexpr int kRtThreads = 128;

template <int E>
struct RowCfg {
  static constexpr int TM = (E >= 64) ? 1 : 64 / E;     // rows per thread (64 accumulators per GEMM)
  static constexpr int R = kRtThreads * TM;             // rows per CTA tile
  static constexpr int RP = R + 1;                      // padded row pitch (floats) of a feature-major buffer
  static constexpr int OT = (E == 32) ? 32 : 16;        // obstacle tile (attention keys per step)
};

// cooperative global -> shared copy of n floats (n % 4 == 0, both 16 B aligned), with the barriers that
// protect the previous contents and publish the new ones.
__device__ __forceinline__ void stage_load(float* __restrict__ dst, const float* __restrict__ src, int n) {
  __syncthreads();
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
  __syncthreads();
}

// four FMAs acc[n..n+3] += a * w as two packed fma.rn.f32x2 (FFMA2, sm_100): same IEEE result per element, half
// the issue slots of four FFMAs -- measured +17 % on this inner loop (tools/microbench/gemm_inner.cu)
__device__ __forceinline__ void fma4(float* __restrict__ acc, float a, const float4 w) {
  const float2 a2 = make_float2(a, a);
  float2 c0 = make_float2(acc[0], acc[1]), c1 = make_float2(acc[2], acc[3]);
  c0 = __ffma2_rn(a2, make_float2(w.x, w.y), c0);
  c1 = __ffma2_rn(a2, make_float2(w.z, w.w), c1);
  acc[0] = c0.x; acc[1] = c0.y; acc[2] = c1.x; acc[3] = c1.y;
}

// acc[r][n] = bias[n]
template <int TM, int N>
__device__ __forceinline__ void acc_init_bias(float (&acc)[TM][N], const float* __restrict__ bias) {
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      acc[r][n] = b.x; acc[r][n + 1] = b.y; acc[r][n + 2] = b.z; acc[r][n + 3] = b.w;
    }
  }
}

template <int TM, int N>
__device__ __forceinline__ void acc_zero(float (&acc)[TM][N]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int n = 0; n < N; ++n) acc[r][n] = 0.0f;
}

// acc[r][:] += sum_k A[k][row_r] * Wt[k][:]   -- A: this thread's column base in a feature-major buffer
template <int K, int N, int TM, int RP>
__device__ __forceinline__ void gemm_smem(float (&acc)[TM][N], const float* __restrict__ a_col,
                                          const float* __restrict__ wt) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) a[r] = a_col[k * RP + r * kRtThreads];
#pragma unroll
    for (int n = 0; n < N; n += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + k * N + n);
#pragma unroll
      for (int r = 0; r < TM; ++r) fma4(&acc[r][n], a[r], w);
    }
  }
}

// same with the K inputs in registers (input layers: K in {c, 2c, 4c, s, c+3})
template <int K, int N, int TM>
__device__ __forceinline__ void gemm_reg(float (&acc)[TM][N], const float (&in)[TM][K], const float* __restrict__ wt) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int n = 0; n < N; n += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + k * N + n);
#pragma unroll
      for (int r = 0; r < TM; ++r) fma4(&acc[r][n], in[r][k], w);
    }
  }
}

template <int TM, int N>
__device__ __forceinline__ void acc_relu(float (&acc)[TM][N]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int n = 0; n < N; ++n) acc[r][n] = fmaxf(acc[r][n], 0.0f);
}

// store accumulators into this thread's columns of a feature-major buffer
template <int TM, int N, int RP>
__device__ __forceinline__ void acc_store(const float (&acc)[TM][N], float* __restrict__ col) {
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int r = 0; r < TM; ++r) col[n * RP + r * kRtThreads] = acc[r][n];
}

template <int TM, int N, int RP>
__device__ __forceinline__ void acc_add_col(float (&acc)[TM][N], const float* __restrict__ col) {
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r][n] += col[n * RP + r * kRtThreads];
}

// LayerNorm over the N features of each row, biased variance, (x-mu)/sqrt(var+eps)*g+b  (torch.nn.LayerNorm)
template <int TM, int N>
__device__ __forceinline__ void acc_layernorm(float (&acc)[TM][N], const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps) {
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    float mu = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) mu += acc[r][n];
    mu *= (1.0f / N);
    float var = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float d = acc[r][n] - mu;
      var = fmaf(d, d, var);
    }
    var *= (1.0f / N);
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int n = 0; n < N; ++n) acc[r][n] = (acc[r][n] - mu) * rstd * gamma[n] + beta[n];
  }
}

// row vector (N contiguous floats in global memory) -> this thread's column;  p == nullptr -> zeros
// (rows are N * 4 B apart, N a multiple of 8, in 256-byte-aligned buffers: one 256-bit access per 8 floats -- whole 32-byte sectors
//  per lane and half the LSU instructions of float4s in these row-per-thread patterns)
template <int N, int RP>
__device__ __forceinline__ void col_load_global(float* __restrict__ col1, const float* __restrict__ p) {
  static_assert(N % 8 == 0, "N % 8");
#pragma unroll
  for (int n = 0; n < N; n += 8) {
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p)
      asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7]) : "l"(p + n));
#pragma unroll
    for (int i = 0; i < 8; ++i) col1[(n + i) * RP] = x[i];
  }
}

// accumulators of sub-row r -> N contiguous floats in global memory
template <int TM, int N>
__device__ __forceinline__ void acc_store_global(const float (&acc)[TM][N], int r, float* __restrict__ p) {
  static_assert(N % 8 == 0, "N % 8");
#pragma unroll
  for (int n = 0; n < N; n += 8)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + n), "f"(acc[r][n]), "f"(acc[r][n + 1]), "f"(acc[r][n + 2]),
                 "f"(acc[r][n + 3]), "f"(acc[r][n + 4]), "f"(acc[r][n + 5]), "f"(acc[r][n + 6]), "f"(acc[r][n + 7]) : "memory");
}

__device__ __forceinline__ int find_segment(const int32_t* __restrict__ ptr, int n_seg, int x) {
  int lo = 0, hi = n_seg;  // largest g with ptr[g] <= x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace gmp
