// Per-problem result rows for the final success/cost reduction.
//
// The reference reduces one Python tuple per problem at the end of eval_gnn (eval_gnn.py:120-134).  For a batch
// sharded over GPUs the per-problem row is produced on the device by one CTA per graph and is the ONLY data
// that crosses NVLink (all-gather of [B,4] floats per rank).
//   row = (problem id, E_g, number of collision-free edges, best edge logit)
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

}  // namespace
}  // namespace gmp

using namespace gmp;

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
