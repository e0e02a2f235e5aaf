// Per-problem result rows for the final success/cost reduction.
//
// The reference reduces one Python tuple per problem at the end of eval_gnn (eval_gnn.py:120-134).  For a batch
// sharded over GPUs the per-problem row is produced on the device by one CTA per graph and is the ONLY data
// that crosses NVLink (all-gather of [B,4] floats per rank).
//   row = (problem id, E_g, number of collision-free edges, best edge logit)
#include <algorithm>

#include "common.cuh"

namespace gmp {
namespace {

__global__ void __launch_bounds__(256) result_rows_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ edge_free,
                                                          const int32_t* __restrict__ edge_ptr, int first_problem,
                                                          float* __restrict__ rows) {
  __shared__ float s_max[8];
  __shared__ int s_cnt[8];
  const int g = blockIdx.x;
  const int e0 = edge_ptr[g], e1 = edge_ptr[g + 1];
  float mx = -INFINITY;
  int cnt = 0;
  for (int e = e0 + threadIdx.x; e < e1; e += 256) {
    mx = fmaxf(mx, logits[e]);
    if (edge_free) cnt += edge_free[e];
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = mx; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, s_max[w]); cnt += s_cnt[w]; }
    rows[g * 4 + 0] = (float)(first_problem + g);
    rows[g * 4 + 1] = (float)(e1 - e0);
    rows[g * 4 + 2] = (float)cnt;
    rows[g * 4 + 3] = mx;
  }
}

// local node ids as int16 / int32 for the read-back: the planner-facing int64 [2,E] (torch_geometric layout) costs 16 B/edge
// over PCIe for ids below 2 000 -- 76 % of the batched path's device->host bytes
template <typename T>
__global__ void __launch_bounds__(256) narrow_ids_kernel(const int64_t* __restrict__ ei, int64_t row_stride, int64_t n,
                                                         T* __restrict__ out, int64_t out_stride) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    out[i] = (T)ei[i];
    out[out_stride + i] = (T)ei[row_stride + i];
  }
}

// a few words device -> MAPPED pinned host memory with plain stores (posted PCIe writes): unlike cudaMemcpyAsync it does not queue
// on the device->host copy engine behind a large read-back that is still in flight
__global__ void __launch_bounds__(256) post_words_kernel(const uint32_t* __restrict__ src, int64_t n, volatile uint32_t* dst_host) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) dst_host[i] = src[i];
  __threadfence_system();
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" int gmp_post_to_host(const void* src_device, int64_t nbytes, void* dst_mapped_host, void* stream) {
  GMP_REQUIRE(nbytes >= 0 && nbytes % 4 == 0 && nbytes <= (1 << 20), "nbytes must be a multiple of 4, at most 1 MiB (small control data only)");
  if (nbytes == 0) return GMP_OK;
  GMP_REQUIRE(src_device && dst_mapped_host, "null pointer");
  const int64_t n = nbytes / 4;
  post_words_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint32_t*>(src_device), n, static_cast<volatile uint32_t*>(dst_mapped_host));
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_edge_index_narrow(const int64_t* edge_index, int64_t edge_row_stride, int64_t n_edges, int bits,
                                     void* out, int64_t out_row_stride, void* stream) {
  GMP_REQUIRE(n_edges >= 0 && (bits == 16 || bits == 32), "bits must be 16 or 32");
  if (n_edges == 0) return GMP_OK;
  GMP_REQUIRE(edge_index && out && edge_row_stride >= n_edges && out_row_stride >= n_edges, "null pointer / bad stride");
  const int grid = (int)std::min<int64_t>((n_edges + 255) / 256, kNumSMs * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bits == 16)
    narrow_ids_kernel<int16_t><<<grid, 256, 0, st>>>(edge_index, edge_row_stride, n_edges, static_cast<int16_t*>(out), out_row_stride);
  else
    narrow_ids_kernel<int32_t><<<grid, 256, 0, st>>>(edge_index, edge_row_stride, n_edges, static_cast<int32_t*>(out), out_row_stride);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_result_rows(const float* edge_logits, const uint8_t* edge_free, const int32_t* edge_ptr, int64_t n_graphs,
                               int32_t first_problem_id, float* rows_out, void* stream) {
  GMP_REQUIRE(n_graphs >= 0, "n_graphs < 0");
  if (n_graphs == 0) return GMP_OK;
  GMP_REQUIRE(edge_logits && edge_ptr && rows_out, "null pointer");
  result_rows_kernel<<<(unsigned)n_graphs, 256, 0, static_cast<cudaStream_t>(stream)>>>(edge_logits, edge_free, edge_ptr,
                                                                                      first_problem_id, rows_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}
