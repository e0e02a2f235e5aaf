// Microbenchmark / bring-up test of the tcgen05 primitives in csrc/umma.cuh (B200 only):
//   1. MMA issue patterns: n back-to-back tcgen05.mma (M = 128, K = 8) from one or two issuing warps, operands in uniform
//      registers (warp-uniform branch + elect.sync): 128*N/256 cycles per MMA.  (Issued from a divergent branch with
//      per-thread values ptxas emits an ELECT / R2UR waterfall per MMA: ~62 cycles each regardless of N.)
//   2. D = A.B^T with A in shared memory (SS) and in TMEM (TS), 1xTF32 and 3xTF32, error against fp64;
//      tells whether operand conversion truncates and how close 3xTF32 gets to fp32;
//   3. cycle counts: MMA group issue -> commit latency, tcgen05.ld / tcgen05.st throughput with 4 and 8 warps;
//   4. tcgen05.ld throughput by shape: 32x32b.x32 ~410 B/cycle/SM, 16x256b.x8 ~190, 16x128b.x16 ~290 (one load in flight per warp);
//   5. four loads in flight per warp: 550 (4 warps) - 850 (8 warps) B/cycle/SM.
// Lesson kept in the code: index the destination arrays of tcgen05.ld with compile-time constants only -- a dynamic index moves the
// array to local memory and the loop then times the spills (an early version of part 3 reported 57 B/cycle that way).
// Results of the run this library was designed around: profiles/r1_umma_tf32_microbench.log.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gnn_motion_planning_b200/csrc \
//        -o tools/microbench/umma_tf32 tools/microbench/umma_tf32.cu && tools/microbench/umma_tf32
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "umma.cuh"

namespace gmp {
void set_error(const std::string&) {}
int cuda_fail(cudaError_t, const char*, const char*, int) { return -1; }
}  // namespace gmp
using namespace gmp;

constexpr int M = 128;

// mode bit0: 3xTF32 (else 1x, raw fp32 bits fed to the tensor core); bit1: A from TMEM
template <int N, int K>
__global__ void __launch_bounds__(128) gemm_test(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                 int mode, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* Ah = reinterpret_cast<float*>(smem);            // [K/4][128][4]
  float* Al = Ah + M * K;
  float* Bh = Al + M * K;                                // [K/4][N][4]
  float* Bl = Bh + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  const bool x3 = mode & 1, ts = mode & 2;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (t == 0) mbar_init(&bar, 1);
  // operands -> shared memory
  for (int i = t; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const float x = A[i];
    const float h = x3 ? umma::tf32_hi(x) : x;
    Ah[((k / 4) * M + r) * 4 + (k % 4)] = h;
    Al[((k / 4) * M + r) * 4 + (k % 4)] = x - h;
  }
  for (int i = t; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const float x = B[i];
    const float h = x3 ? umma::tf32_hi(x) : x;
    Bh[((k / 4) * N + n) * 4 + (k % 4)] = h;
    Bl[((k / 4) * N + n) * 4 + (k % 4)] = x - h;
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  const uint32_t colD = 0, colAh = 64, colAl = 64 + K;   // K <= 64, N <= 64
  if (ts) {
    float row[K];
    for (int k = 0; k < K; ++k) row[k] = A[t * K + k];
    if (x3) {
      umma::st_split<K>(lane_base + colAh, lane_base + colAl, row);
    } else {
      for (int c = 0; c < K; c += 8) umma::st8(lane_base + colAh + c, row + c);
    }
    umma::wait_st();
  }
  umma::fence_before_sync();
  __syncthreads();
  long long t0 = clock64();
  if (t == 0) {
    umma::fence_after_sync();
    const uint32_t id = umma::idesc_tf32(M, N);
    if (ts) {
      if (x3) {
        umma::gemm3_ts(tm + colD, tm + colAh, tm + colAl, smem_u32(Bh), smem_u32(Bl), N, 0, N, K, false);
      } else {
        for (int ks = 0; ks < K / 8; ++ks) umma::mma_ts(tm + colD, tm + colAh + ks * 8, umma::kmajor_desc(smem_u32(Bh), N, ks), id, ks > 0);
      }
    } else {
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t ah = umma::kmajor_desc(smem_u32(Ah), M, ks), al = umma::kmajor_desc(smem_u32(Al), M, ks);
        const uint64_t bh = umma::kmajor_desc(smem_u32(Bh), N, ks), bl = umma::kmajor_desc(smem_u32(Bl), N, ks);
        if (x3) {
          umma::mma_ss(tm + colD, al, bh, id, ks > 0);
          umma::mma_ss(tm + colD, ah, bl, id, 1);
          umma::mma_ss(tm + colD, ah, bh, id, 1);
        } else {
          umma::mma_ss(tm + colD, ah, bh, id, ks > 0);
        }
      }
    }
    umma::commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  long long t1 = clock64();
  umma::fence_after_sync();
  float out[N];
  for (int c = 0; c < N; c += 8) umma::ld8(lane_base + colD + c, out + c);
  umma::wait_ld();
  for (int n = 0; n < N; ++n) D[t * N + n] = out[n];
  if (t == 0 && cycles) *cycles = t1 - t0;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 256);
}

// TMEM load / store throughput: every warp moves `iters` x 32 columns of its lane quarter
__global__ void __launch_bounds__(256) tmem_bw(int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float r[32];
  for (int i = 0; i < 32; ++i) r[i] = (float)(t + i);
  umma::st32(base, r);
  umma::wait_st();
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    umma::ld32(base + (it & 7) * 32, r);
    umma::wait_ld();
    acc += r[0] + r[31];   // static indices: a dynamic index would push r[] to local memory and time the spills instead
  }
  __syncthreads();
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) {
    r[0] += 1.0f;
    umma::st32(base + (it & 7) * 32, r);
  }
  umma::wait_st();
  __syncthreads();
  long long t2 = clock64();
  if (t == 0) { out[0] = t1 - t0; out[1] = t2 - t1; }
  sink[t] = acc + r[0];
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_slot, 512);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int N, int K>
void run_case(const char* name) {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD; long long* dc;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
  cudaFuncSetAttribute(gemm_test<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(dD, 0, D.size() * 4);
    long long cyc = 0;
    for (int rep = 0; rep < 2; ++rep) gemm_test<N, K><<<1, 128, smem>>>(dA, dB, dD, mode, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s mode %d: CUDA error %s\n", name, mode, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    double e64 = 0, e32 = 0, etr = 0, ern = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0, rtr = 0, rrn = 0; float f32 = 0.f;
        for (int k = 0; k < K; ++k) {
          ref += (double)A[m * K + k] * (double)B[n * K + k];
          f32 = fmaf(A[m * K + k], B[n * K + k], f32);
          rtr += (double)tf32_trunc(A[m * K + k]) * (double)tf32_trunc(B[n * K + k]);
          rrn += (double)tf32_rn(A[m * K + k]) * (double)tf32_rn(B[n * K + k]);
        }
        const double d = D[m * N + n];
        e64 = fmax(e64, fabs(d - ref)); e32 = fmax(e32, fabs((double)f32 - ref));
        etr = fmax(etr, fabs(d - rtr)); ern = fmax(ern, fabs(d - rrn));
      }
    printf("%s N=%d K=%d %s %s: max|D-fp64|=%.3e (fp32 fma chain: %.3e)  |D-trunc model|=%.3e |D-rn model|=%.3e  issue->commit %lld cyc\n",
           name, N, K, (mode & 2) ? "A=TMEM" : "A=SMEM", (mode & 1) ? "3xTF32" : "1xTF32", e64, e32, etr, ern, cyc);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
}

int main2();
int main4();
int main5();
int main() {
  main2();
  main4();
  main5();
  return 0;
  run_case<32, 32>("gemm");
  run_case<64, 32>("gemm");
  run_case<32, 64>("gemm");
  run_case<64, 64>("gemm");
  long long* dout; float* sink; long long h[2];
  cudaMalloc(&dout, 16); cudaMalloc(&sink, 256 * 4);
  for (int threads : {128, 256}) {
    const int iters = 4096;
    tmem_bw<<<1, threads>>>(iters, dout, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem_bw: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * threads * 32 * 4;
    printf("tmem %d threads: ld32+wait %.1f cyc/iter (%.1f B/cyc/SM), st32 %.1f cyc/iter (%.1f B/cyc/SM)\n", threads, (double)h[0] / iters,
           bytes / h[0], (double)h[1] / iters, bytes / h[1]);
  }
  return 0;
}

// ---- part 3: MMA issue patterns.  `n_mma` TS-mode MMAs (M=128, K=8 each) round-robin over `n_acc` accumulators,
// issued by `n_issuers` threads (one per warpgroup-of-4 warps) each with its own accumulators + barrier.
template <int N>
__global__ void __launch_bounds__(256) mma_pattern(int n_mma, int n_acc, int n_issuers, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* Bh = reinterpret_cast<float*>(smem);   // [2][N][4] one k-step of B, zeros
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 2 * N * 4; i += blockDim.x) Bh[i] = 0.f;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (t == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_slot;
  const int grp = t >> 7;
  long long t0 = clock64();
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);          // provably warp-uniform
  if ((warp_u & 3) == 0 && (warp_u >> 2) < n_issuers) {
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tm, 0);
    const int grp_u = warp_u >> 2;
    if (umma::elect_one()) {
      const uint32_t id = umma::idesc_tf32(128, N);
      const uint64_t bd = umma::kmajor_desc(smem_u32(Bh), N, 0);
      const uint32_t d0 = tm_u + grp_u * 256, a0 = tm_u + grp_u * 256 + 192;
      const uint32_t amask = (uint32_t)(n_acc - 1);   // n_acc is a power of two
#pragma unroll 8
      for (int i = 0; i < n_mma; ++i) umma::mma_ts(d0 + (i & amask) * N, a0 + (i & 7) * 8, bd, id, 1u);
      umma::commit(&bar[grp_u]);
    }
    __syncwarp();
  }
  if (grp < n_issuers) mbar_wait(&bar[grp], 0);
  long long t1 = clock64();
  __syncthreads();
  if (t == 0) out[0] = t1 - t0;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
}

template <int N>
void run_pattern(long long* dout) {
  const size_t smem = 2 * N * 16;
  for (int n_mma : {6, 12, 24, 48, 96, 192})
    for (int issuers : {1, 2}) {
      const int n_acc = (N <= 64 && n_mma == 96) ? 2 : 1;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) mma_pattern<N><<<1, 256, smem>>>(n_mma, n_acc, issuers, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mma_pattern: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
      cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
      printf("mma N=%3d: %3d MMAs/issuer, %d issuer(s), %d accumulator(s): %lld cyc total, %.1f cyc/MMA/issuer\n", N, n_mma, issuers, n_acc, h,
             (double)h / n_mma);
    }
}

int main2() {
  long long* dout;
  cudaMalloc(&dout, 16);
  run_pattern<32>(dout);
  run_pattern<64>(dout);
  run_pattern<128>(dout);
  return 0;
}

// ---- part 4: does the tcgen05.ld shape change the TMEM -> register bandwidth?  (4 KB per warp instruction in every case)
template <int SHAPE>
__global__ void __launch_bounds__(128) tmem_ld_shape(int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t base = tmem_slot + ((uint32_t)(warp * 32) << 16);
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (it & 7) * 32;
    if (SHAPE == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
    } else if (SHAPE == 1) {
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
    } else {
      asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
    }
    umma::wait_ld();
    acc += r[0] + r[31];   // static indices (see part 5)
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) out[0] = t1 - t0;
  sink[t] = (float)acc;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_slot, 512);
}

int main4() {
  long long* dout; float* sink; long long h = 0;
  cudaMalloc(&dout, 16); cudaMalloc(&sink, 256 * 4);
  const int iters = 4096;
  const char* names[3] = {"32x32b.x32", "16x256b.x8", "16x128b.x16"};
  for (int sh = 0; sh < 3; ++sh) {
    if (sh == 0) tmem_ld_shape<0><<<1, 128>>>(iters, dout, sink);
    if (sh == 1) tmem_ld_shape<1><<<1, 128>>>(iters, dout, sink);
    if (sh == 2) tmem_ld_shape<2><<<1, 128>>>(iters, dout, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem_ld_shape %s: CUDA error %s\n", names[sh], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld.%s, 4 warps: %.1f cyc per 16 KB (%.1f B/cyc/SM)\n", names[sh], (double)h / iters, (double)iters * 128 * 32 * 4 / h);
  }
  return 0;
}

// ---- part 5: four tcgen05.ld (32x32b.x32) in flight per warp before one wait
__global__ void __launch_bounds__(256) tmem_ld_depth(int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float r0[32], r1[32], r2[32], r3[32];
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    umma::ld32(base, r0); umma::ld32(base + 32, r1); umma::ld32(base + 64, r2); umma::ld32(base + 96, r3);
    umma::wait_ld();
    acc += r0[0] + r1[7] + r2[19] + r3[31];   // static indices: a dynamic index would push the arrays to local memory
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) out[0] = t1 - t0;
  sink[t] = acc;
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_slot, 512);
}

int main5() {
  long long* dout; float* sink; long long h = 0;
  cudaMalloc(&dout, 16); cudaMalloc(&sink, 256 * 4);
  const int iters = 2048;
  for (int threads : {128, 256}) {
    tmem_ld_depth<<<1, threads>>>(iters, dout, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem_ld_depth: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld x4 in flight, %d threads: %.1f cyc per iteration (%.1f B/cyc/SM)\n", threads, (double)h / iters,
           (double)iters * threads * 128 * 4 / h);
  }
  return 0;
}
