// Microbenchmark of the row-tile GEMM inner loop (shared-memory activations, broadcast weights):
// FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100), TM rows per thread, CTAs per SM.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_inner gemm_inner.cu && ./gemm_inner
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <bool PACKED, int K, int N, int TM, int RP>
__device__ __forceinline__ void gemm(float (&acc)[TM][N], const float* __restrict__ a_col, const float* __restrict__ wt) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) a[r] = a_col[k * RP + r * 128];
#pragma unroll
    for (int n = 0; n < N; n += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + k * N + n);
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        if (PACKED) {
          float2 c0 = make_float2(acc[r][n], acc[r][n + 1]), c1 = make_float2(acc[r][n + 2], acc[r][n + 3]);
          const float2 a2 = make_float2(a[r], a[r]);
          c0 = __ffma2_rn(a2, make_float2(w.x, w.y), c0);
          c1 = __ffma2_rn(a2, make_float2(w.z, w.w), c1);
          acc[r][n] = c0.x; acc[r][n + 1] = c0.y; acc[r][n + 2] = c1.x; acc[r][n + 3] = c1.y;
        } else {
          acc[r][n] = fmaf(a[r], w.x, acc[r][n]); acc[r][n + 1] = fmaf(a[r], w.y, acc[r][n + 1]);
          acc[r][n + 2] = fmaf(a[r], w.z, acc[r][n + 2]); acc[r][n + 3] = fmaf(a[r], w.w, acc[r][n + 3]);
        }
      }
    }
  }
}

template <bool PACKED, int TM, int MINB>
__global__ void __launch_bounds__(128, MINB) bench(const float* g, float* out, int iters) {
  constexpr int RP = 128 * TM + 1;
  extern __shared__ float sm[];
  float* X = sm;
  float* W = sm + 32 * RP;
  for (int i = threadIdx.x; i < 32 * RP + 1024; i += 128) sm[i] = g[i % 4096];
  __syncthreads();
  float acc[TM][32];
  for (int r = 0; r < TM; ++r) for (int n = 0; n < 32; ++n) acc[r][n] = 0.f;
  for (int it = 0; it < iters; ++it) gemm<PACKED, 32, 32, TM, RP>(acc, X + threadIdx.x, W);
  float s = 0;
  for (int r = 0; r < TM; ++r) for (int n = 0; n < 32; ++n) s += acc[r][n];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <bool PACKED, int TM, int MINB>
void run(const char* name, const float* g, float* out) {
  constexpr int RP = 128 * TM + 1;
  const size_t smem = (32 * RP + 1024) * sizeof(float);
  cudaFuncSetAttribute(bench<PACKED, TM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench<PACKED, TM, MINB>, 128, smem);
  const int grid = 148 * occ * 4, iters = 400;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  bench<PACKED, TM, MINB><<<grid, 128, smem>>>(g, out, 10);
  cudaEventRecord(a);
  bench<PACKED, TM, MINB><<<grid, 128, smem>>>(g, out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double flop = 2.0 * grid * 128.0 * TM * 32 * 32 * iters;
  printf("%-28s occ=%d CTAs/SM  %.3f ms  %.1f TFLOP/s  (%s)\n", name, occ, ms, flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *g, *out;
  cudaMalloc(&g, 4096 * 4); cudaMalloc(&out, 148 * 64 * 128 * 4);
  std::vector<float> h(4096, 0.001f);
  cudaMemcpy(g, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
  run<false, 2, 3>("FFMA  TM=2 minb3", g, out);
  run<true, 2, 3>("FFMA2 TM=2 minb3", g, out);
  run<false, 2, 4>("FFMA  TM=2 minb4", g, out);
  run<true, 2, 4>("FFMA2 TM=2 minb4", g, out);
  run<false, 4, 2>("FFMA  TM=4 minb2", g, out);
  run<true, 4, 2>("FFMA2 TM=4 minb2", g, out);
  run<false, 1, 4>("FFMA  TM=1 minb4", g, out);
  run<true, 1, 4>("FFMA2 TM=1 minb4", g, out);
  run<true, 1, 8>("FFMA2 TM=1 minb8", g, out);
  return 0;
}
