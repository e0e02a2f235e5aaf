// tcgen05 / TMEM primitives (sm_100a inline PTX) used by the tensor-core edge-feature kernel.
//
// Conventions used throughout this library:
//   * MMA shape M = 128 (cta_group::1): accumulator row m lives in TMEM lane m, column n in TMEM column base + n;
//     with the 32x32b load/store shape thread t of warp w touches lane 32*(w%4) + t, so in a 128-thread
//     group "thread == row" -- the same ownership as the SIMT row-tile kernels (rowtile.cuh);
//   * kind::tf32, fp32 accumulate.  One instruction consumes K = 8 (32 bytes of K per row);
//   * shared-memory operands are K-major, no swizzle ("interleaved" canonical layout): a matrix of
//     ROWS x K floats is stored as  float[K/4][ROWS][4]  -- 8-row x 16-byte core matrices, consecutive
//     8-row groups 128 B apart (stride byte offset), consecutive 16-byte K chunks ROWS*16 B apart
//     (leading byte offset);
//   * the A operand may instead come from TMEM (row = lane, K index = column), which is how the
//     activations of a chained MLP are fed back without touching shared memory.
//
// 3xTF32: x = hi + lo with hi = x with the 13 low mantissa bits cleared (exactly a TF32 number) and
// lo = x - hi (exact in fp32); A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi accumulated in fp32.  The dropped
// lo.lo term and the truncation of lo to TF32 are both ~2^-22 relative.
#pragma once
#include "common.cuh"

namespace gmp {
namespace umma {

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// descriptor of float[K/4][ROWS][4] starting at K chunk pair `kstep` (8 K values per step)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr, int rows, int kstep) {
  return smem_desc(saddr + (uint32_t)kstep * 2u * (uint32_t)rows * 16u, (uint32_t)rows * 16u, 128u);
}
// instruction descriptor: D f32, A/B tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T      (one thread issues)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers, 32x32b shape: this thread's lane, N consecutive columns ----------------------------------
__device__ __forceinline__ void ld8(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]),
               "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]),
      "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]),
      "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]),
      "r"(u[31])
      : "memory");
}

// TF32 split of an fp32 value: hi = round-to-nearest TF32 of x, lo = round-to-nearest TF32 of the (exact) remainder.
// Rounding (instead of letting the tensor core truncate the operands) keeps the split error unbiased at ~2^-23 |x|.
__device__ __forceinline__ float tf32_rna(float x) {
  // (cvt.rna.tf32.f32 costs four SASS instructions because it special-cases NaN; the bit trick is two and maps
  //  +-inf to itself, which is all the epilogues can produce)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_hi(float x) { return tf32_rna(x); }

// store N values (N % 8 == 0) of this thread's row as hi / lo planes at TMEM columns hi_addr / lo_addr.
// hi is rounded to nearest; lo = x - hi is exact in fp32 and is stored as is: the tensor core drops its 13 low
// mantissa bits on read (measured: operand conversion truncates, tools/microbench/umma_tf32.cu), an error of at most
// 2^-10 |lo| <= 2^-22 |x| -- 3 instructions per element instead of 5 with an explicitly rounded lo (ROUND_LO).
template <int N, bool ROUND_LO = false>
__device__ __forceinline__ void st_split(uint32_t hi_addr, uint32_t lo_addr, const float* x) {
#pragma unroll
  for (int c = 0; c < N; c += 8) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // (truncating hi instead -- one LOP less per element -- measured 7.56 -> 7.34 ms but biases the split and raised the worst
      //  tc-vs-simt logit difference from 1.9e-5 to 3.6e-5 against a 1e-4 tolerance: rejected, profiles/r2_rd_issuer.md)
      h[i] = tf32_rna(x[c + i]);
      l[i] = ROUND_LO ? tf32_rna(x[c + i] - h[i]) : x[c + i] - h[i];
    }
    st8(hi_addr + c, h);
    st8(lo_addr + c, l);
  }
}

// 2^x for x <= 0 (softmax weights): one MUFU.EX2; results below 2^-126 flush to zero
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one 256-bit global store (sm_100: STG.256): the lane writes a whole 32-byte sector, so a row-per-lane store pattern costs half
// the store instructions and no partial-sector writes
__device__ __forceinline__ void st_global_v8(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6),
               "f"(a7) : "memory");
}

// ... and the matching 256-bit load (LDG.256)
__device__ __forceinline__ void ld_global_v8(const float* p, float* a) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait for an mbarrier phase.  Bounded: a protocol bug traps instead of hanging the GPU -- but only after 2^30 polls (tens of
// seconds; round 1 trapped after 2^22, which a long legitimate wait under time-slicing / MPS / a debugger could reach, ADVICE
// r1).  Measured alternatives (profiles/r2_mbar_wait_sweep.log): a suspend-time hint on try_wait (+0.3 ms on the 8.8 ms
// edge-feature kernel: coarser wake-up), __nanosleep back-off (no gain: the poll loop is 30 % of the executed instructions
// but the kernel is bound by the latency of its MMA round trips, not by issue slots), a %globaltimer-based bound (+1-3 %).
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t phase) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(phase)
                 : "memory");
    if (done) break;
    if (++spins > (1u << 30)) __trap();
  }
}

// ---- 3xTF32 products with A in TMEM (hi / lo column planes) and B in shared memory (hi / lo K-major planes):
// D[128 x N] (+)= A[128 x K] . B[N x K]^T.  One elected thread issues; everything it touches should be warp-uniform
// so that the descriptors live in uniform registers and the UTCHMMA instructions issue back to back.
//
// A no-swizzle K-major descriptor is { hi word = 0x4008 (stride byte offset 128 B, version 1),
//                                      lo word = (shared address >> 4) | (B_ROWS << 16)  (leading byte offset = B_ROWS*16 B) }
// and advancing by one K step of 8 adds 2*B_ROWS to the lo word.
__device__ __forceinline__ uint32_t desc_lo32(uint32_t saddr, uint32_t b_rows) { return ((saddr & 0x3FFFFu) >> 4) | (b_rows << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo32) { return ((uint64_t)0x4008u << 32) | (uint64_t)lo32; }

// compile-time shape (weights): fully unrolled
template <int N, int K, int B_ROWS>
__device__ __forceinline__ void gemm3_fixed(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo32, uint32_t bl_lo32,
                                            bool accumulate) {
  constexpr uint32_t id = idesc_tf32(128, N);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t bh = desc64(bh_lo32 + (uint32_t)(ks * 2 * B_ROWS)), bl = desc64(bl_lo32 + (uint32_t)(ks * 2 * B_ROWS));
    mma_ts(d_tmem, a_lo + ks * 8, bh, id, (ks > 0 || accumulate) ? 1u : 0u);
    mma_ts(d_tmem, a_hi + ks * 8, bl, id, 1u);
    mma_ts(d_tmem, a_hi + ks * 8, bh, id, 1u);
  }
}
// runtime N (scores against `n` obstacle rows laid out with b_rows = n), K compile time
template <int K>
__device__ __forceinline__ void gemm3_n(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo32, uint32_t bl_lo32, int n) {
  const uint32_t id = idesc_tf32(128, n);
  const uint32_t step = 2u * (uint32_t)n;
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t bh = desc64(bh_lo32 + ks * step), bl = desc64(bl_lo32 + ks * step);
    mma_ts(d_tmem, a_lo + ks * 8, bh, id, ks > 0 ? 1u : 0u);
    mma_ts(d_tmem, a_hi + ks * 8, bl, id, 1u);
    mma_ts(d_tmem, a_hi + ks * 8, bh, id, 1u);
  }
}
// runtime K (probabilities . values), N and B_ROWS compile time
template <int N, int B_ROWS>
__device__ __forceinline__ void gemm3_k(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo32, uint32_t bl_lo32, int k) {
  constexpr uint32_t id = idesc_tf32(128, N);
  uint32_t acc = 0u;
#pragma unroll 2
  for (int ks = 0; ks < k / 8; ++ks) {
    const uint64_t bh = desc64(bh_lo32), bl = desc64(bl_lo32);
    mma_ts(d_tmem, a_lo, bh, id, acc);
    mma_ts(d_tmem, a_hi, bl, id, 1u);
    mma_ts(d_tmem, a_hi, bh, id, 1u);
    acc = 1u;
    a_hi += 8; a_lo += 8; bh_lo32 += 2 * B_ROWS; bl_lo32 += 2 * B_ROWS;
  }
}

// generic runtime-shape version (bring-up / microbenchmark)
__device__ __forceinline__ void gemm3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_saddr, uint32_t b_lo_saddr,
                                         int b_rows, int b_row0, int N, int K, bool accumulate) {
  const uint32_t id = idesc_tf32(128, N);
  uint32_t acc = accumulate ? 1u : 0u;
  const uint32_t roff = (uint32_t)b_row0 * 16u;
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t bh = kmajor_desc(b_hi_saddr + roff, b_rows, ks);
    const uint64_t bl = kmajor_desc(b_lo_saddr + roff, b_rows, ks);
    mma_ts(d_tmem, a_lo + ks * 8, bh, id, acc);
    mma_ts(d_tmem, a_hi + ks * 8, bl, id, 1u);
    mma_ts(d_tmem, a_hi + ks * 8, bh, id, 1u);
    acc = 1u;
  }
}

}  // namespace umma
}  // namespace gmp
