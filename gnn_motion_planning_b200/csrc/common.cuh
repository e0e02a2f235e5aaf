// Shared helpers for libgnnmp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gnnmp.h"

namespace gmp {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GMP_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return gmp::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define GMP_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      gmp::set_error(std::string("invalid argument: ") + (msg)); \
      return GMP_E_INVALID;                            \
    }                                                  \
  } while (0)

#define GMP_LAUNCH_CHECK() GMP_CUDA(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Carver {
  char* base;
  int64_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(int64_t count) {
    off = align_up(off, 256);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * (int64_t)sizeof(T);
    return r;
  }
  int64_t bytes() const { return align_up(off, 256); }
};

// float atomic max through the sign trick; *addr must have been initialised (e.g. to -inf).
__device__ __forceinline__ void atomic_max_f32(float* addr, float val) {
  val = __fadd_rn(val, 0.0f);  // canonicalise -0 -> +0 (as an int, -0 is INT_MIN and would never win)
  if (val >= 0.0f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(val));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers -----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// contiguous global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` (complete_tx).
// dst, src 16 B aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_DONE_%=;\n\t"
      "bra MBAR_WAIT_%=;\n\t"
      "MBAR_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

}  // namespace gmp
