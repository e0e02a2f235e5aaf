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

}  // namespace gmp
