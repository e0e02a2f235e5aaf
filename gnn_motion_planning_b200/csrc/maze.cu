// Batched 2-D maze collision checks (sm_100a).
//
// Replaces, one thread per state / per edge, the reference's scalar NumPy code
//   MazeEnv._transform               environment/maze_env.py:236-239
//   MazeEnv._valid_state             :266-268
//   MazeEnv._point_in_free_space     :270-277
//   MazeEnv._iterative_check_segment :301-314   (recursive bisection -> explicit DFS stack)
//   MazeEnv._edge_fp                 :316-325   (dim == 2 branch)
// bit-exactly, including the number of collision_check_count increments per edge.
//
// Arithmetic follows the input dtype T (f32 or f64) exactly as NumPy does:
//   cell = trunc(((x + 1) * 15) / 2) clamped to 14;   mid = (l + r) / 2;   L1 = |dx| + |dy| > T(0.05)
// None of these contains a mul->add pair, so FMA contraction cannot change a rounding.
//
// Bound: this is HBM/latency work (33 B/edge mandatory traffic, ~3.6 byte-lookups per edge in a
// 225-byte map that lives in L1).  Maps are read through the read-only path; endpoint loads are
// one 8-byte (f32) / 16-byte (f64) vector load per state.
#include "common.cuh"

namespace gmp {
namespace {

constexpr int kW = 15;
constexpr int kStack = 16;  // bisection depth is <= 8 (L1 length <= 4 halves below 0.05 in 7 steps)

template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
  using type = float2;
};
template <>
struct Vec2<double> {
  using type = double2;
};

template <typename T>
__device__ __forceinline__ int cell_of(T x) {
  T t = ((x + T(1)) * T(kW)) / T(2);
  int c = (int)t;  // truncation toward zero, like ndarray.astype(int)
  return c > kW - 1 ? kW - 1 : c;
}

template <typename T>
__device__ __forceinline__ bool in_range(T x, T y) {
  // compared in double against LIMITS = [1., 1.] (maze_env.py:267-268); exact for both dtypes
  return (double)x >= -1.0 && (double)y >= -1.0 && (double)x <= 1.0 && (double)y <= 1.0;
}

template <typename T>
__device__ __forceinline__ bool cell_free(const uint8_t* __restrict__ map, T x, T y) {
  return __ldg(map + cell_of(x) * kW + cell_of(y)) == 0;
}

// maze_env.py:316-325.  Returns free?; cnt = collision_check_count increments.
template <typename T>
__device__ __forceinline__ bool edge_free(const uint8_t* __restrict__ map, T ax, T ay, T bx, T by, int& cnt) {
  cnt = 0;
  if (!in_range(ax, ay) || !in_range(bx, by)) return false;  // :320
  cnt = 1;
  if (!cell_free(map, ax, ay)) return false;                 // :322 (short circuit: b not looked up)
  cnt = 2;
  if (!cell_free(map, bx, by)) return false;

  T lx = ax, ly = ay, rx = bx, ry = by;
  T sx[kStack], sy[kStack], tx[kStack], ty[kStack];
  int sp = 0;
  const T eps = T(5e-2);  // RRT_EPS rounded to T: NumPy-2 weak-scalar comparison (maze_env.py:306)
  while (true) {
    int dc = abs(cell_of(lx) - cell_of(rx)) + abs(cell_of(ly) - cell_of(ry));
    T l1 = fabs(lx - rx) + fabs(ly - ry);
    if (dc > 1 && l1 > eps && sp < kStack) {
      T mx = (lx + rx) / T(2), my = (ly + ry) / T(2);  // :307
      ++cnt;                                            // the midpoint is always in range
      if (!cell_free(map, mx, my)) return false;        // :309-311
      sx[sp] = mx; sy[sp] = my; tx[sp] = rx; ty[sp] = ry;  // right half (mid, r) pending
      ++sp;
      rx = mx; ry = my;                                 // descend into the left half first (:312)
      continue;
    }
    if (sp == 0) return true;
    --sp;
    lx = sx[sp]; ly = sy[sp]; rx = tx[sp]; ry = ty[sp];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maze_state_kernel(const T* __restrict__ states, const uint8_t* __restrict__ maps,
                                                         const int32_t* __restrict__ problem, int64_t n,
                                                         uint8_t* __restrict__ free_out, uint8_t* __restrict__ counted_out) {
  using V = typename Vec2<T>::type;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    V s = reinterpret_cast<const V*>(states)[i];
    const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * (kW * kW);
    bool ok = in_range(s.x, s.y);
    if (counted_out) counted_out[i] = ok ? 1 : 0;
    free_out[i] = (ok && cell_free(map, s.x, s.y)) ? 1 : 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) maze_edge_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                        const uint8_t* __restrict__ maps, const int32_t* __restrict__ problem,
                                                        int64_t n, uint8_t* __restrict__ free_out,
                                                        int32_t* __restrict__ n_checks_out) {
  using V = typename Vec2<T>::type;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    V s = reinterpret_cast<const V*>(a)[i];
    V t = reinterpret_cast<const V*>(b)[i];
    const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * (kW * kW);
    int cnt;
    bool ok = edge_free<T>(map, s.x, s.y, t.x, t.y, cnt);
    free_out[i] = ok ? 1 : 0;
    if (n_checks_out) n_checks_out[i] = cnt;
  }
}

// One CTA row per graph chunk: blockIdx.y = graph, edges of that graph strided over blockIdx.x.
__global__ void __launch_bounds__(256) maze_edge_graph_kernel(const float* __restrict__ v, const int64_t* __restrict__ edge_index,
                                                              int64_t row_stride, const int32_t* __restrict__ node_ptr,
                                                              const int32_t* __restrict__ edge_ptr,
                                                              const int32_t* __restrict__ problem_of_graph,
                                                              const uint8_t* __restrict__ maps, uint8_t* __restrict__ free_out,
                                                              int32_t* __restrict__ n_checks_out) {
  const int g = blockIdx.y;
  const int e0 = edge_ptr[g], e1 = edge_ptr[g + 1];
  const float2* vg = reinterpret_cast<const float2*>(v) + node_ptr[g];
  const uint8_t* map = maps + (int64_t)(problem_of_graph ? problem_of_graph[g] : g) * (kW * kW);
  for (int e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += gridDim.x * blockDim.x) {
    int src = (int)edge_index[e];
    int dst = (int)edge_index[row_stride + e];
    float2 s = __ldg(vg + src);
    float2 t = __ldg(vg + dst);
    int cnt;
    bool ok = edge_free<float>(map, s.x, s.y, t.x, t.y, cnt);
    free_out[e] = ok ? 1 : 0;
    if (n_checks_out) n_checks_out[e] = cnt;
  }
}

inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)kNumSMs * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" int gmp_maze_state_fp(const void* states, int dtype, const uint8_t* maps, const int32_t* problem_of_state,
                                 int64_t n, uint8_t* free_out, uint8_t* counted_out, void* stream) {
  GMP_REQUIRE(n >= 0, "n < 0");
  GMP_REQUIRE(dtype == GMP_DTYPE_F32 || dtype == GMP_DTYPE_F64, "dtype must be GMP_DTYPE_F32 or GMP_DTYPE_F64");
  if (n == 0) return GMP_OK;
  GMP_REQUIRE(states && maps && free_out, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GMP_DTYPE_F32)
    maze_state_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const float*>(states), maps, problem_of_state, n,
                                                               free_out, counted_out);
  else
    maze_state_kernel<double><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const double*>(states), maps, problem_of_state,
                                                                n, free_out, counted_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze_edge_fp(const void* a, const void* b, int dtype, const uint8_t* maps, const int32_t* problem_of_edge,
                                int64_t n, uint8_t* free_out, int32_t* n_checks_out, void* stream) {
  GMP_REQUIRE(n >= 0, "n < 0");
  GMP_REQUIRE(dtype == GMP_DTYPE_F32 || dtype == GMP_DTYPE_F64, "dtype must be GMP_DTYPE_F32 or GMP_DTYPE_F64");
  if (n == 0) return GMP_OK;
  GMP_REQUIRE(a && b && maps && free_out, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GMP_DTYPE_F32)
    maze_edge_kernel<float><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b), maps,
                                                              problem_of_edge, n, free_out, n_checks_out);
  else
    maze_edge_kernel<double><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b),
                                                               maps, problem_of_edge, n, free_out, n_checks_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze_edge_fp_graph(const float* v, const int64_t* edge_index, int64_t edge_row_stride,
                                      const int32_t* node_ptr, const int32_t* edge_ptr, const int32_t* problem_of_graph,
                                      int64_t n_graphs, int64_t n_edges_total, const uint8_t* maps, uint8_t* free_out,
                                      int32_t* n_checks_out, void* stream) {
  GMP_REQUIRE(n_graphs >= 0 && n_edges_total >= 0, "negative size");
  if (n_graphs == 0 || n_edges_total == 0) return GMP_OK;
  GMP_REQUIRE(v && edge_index && node_ptr && edge_ptr && maps && free_out, "null pointer");
  GMP_REQUIRE(n_graphs <= 65535, "n_graphs > 65535 (grid.y limit)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ~E/B edges per graph; enough x-blocks that a graph's edges are covered in ~2 strides
  int64_t per_graph = (n_edges_total + n_graphs - 1) / n_graphs;
  int gx = (int)((per_graph + 511) / 512);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, (unsigned)n_graphs);
  maze_edge_graph_kernel<<<grid, 256, 0, st>>>(v, edge_index, edge_row_stride, node_ptr, edge_ptr, problem_of_graph, maps,
                                               free_out, n_checks_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}
