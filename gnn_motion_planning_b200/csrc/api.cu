// Library-level entry points of libgnnmp.so: error channel, device probe, handle lifetime.
#include "handle.h"

namespace gmp {
namespace {
thread_local std::string g_last_error;
}
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string("CUDA error '") + cudaGetErrorString(e) + "' at " + file + ":" + std::to_string(line) + " in " + what;
  return GMP_E_CUDA;
}
}  // namespace gmp

extern "C" const char* gmp_last_error(void) { return gmp::g_last_error.c_str(); }

extern "C" const char* gmp_version(void) { return "gnnmp-b200 0.1 (sm_100a)"; }

extern "C" int gmp_device_ok(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return 0;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return p.major == 10 ? 1 : 0;
}

extern "C" gmp_handle* gmp_create(int device) {
  if (!gmp_device_ok(device)) {
    gmp::set_error("gmp_create: no sm_100 CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
    return nullptr;
  }
  gmp_handle* h = new gmp_handle();
  h->device = device;
  return h;
}

extern "C" void gmp_destroy(gmp_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->ex.d_weights) cudaFree(h->ex.d_weights);
  if (h->sm.d_weights) cudaFree(h->sm.d_weights);
  if (h->ex_side) cudaStreamDestroy(h->ex_side);
  if (h->ex_fork) cudaEventDestroy(h->ex_fork);
  if (h->ex_join) cudaEventDestroy(h->ex_join);
  delete h;
}
