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
#include "tree_search.cuh"

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

// =====================================================================================================================
// Batched lazy tree search (tree_search.cuh) on the 2-D maze                                   (eval_gnn.py:198-233)
// =====================================================================================================================
struct MazeSearchEnv {
  struct Args { const uint8_t* maps; const int32_t* problem_of_graph; };
  struct Ctx { const uint8_t* map; };
  __device__ static Ctx make(const Args& a, int g) { return Ctx{a.maps + (int64_t)(a.problem_of_graph ? a.problem_of_graph[g] : g) * (kW * kW)}; }
  __device__ static bool edge(const Ctx& c, const float* pa, const float* pb, int& cnt) {
    return edge_free<float>(c.map, pa[0], pa[1], pb[0], pb[1], cnt);
  }
  // in_goal_region (maze_env.py:174-179): distance in float64 against env.goal_state, then one counted state check
  __device__ static bool goal(const Ctx& c, const float* pb, const double* goal, int& cnt) {
    const double d0 = fabs(goal[0] - (double)pb[0]), d1 = fabs(goal[1] - (double)pb[1]);
    if (!(sqrt(d0 * d0 + d1 * d1) < 5e-2)) return false;
    if (!in_range(pb[0], pb[1])) return false;
    cnt += 1;
    return cell_free(c.map, pb[0], pb[1]);
  }
};

// =====================================================================================================================
// Steering rounds of the smoother's caller on the device: proposed_path_smootherv2               (smoother.py:194-216)
// =====================================================================================================================
// One thread per path (the sweep over waypoints is sequential: waypoint i is checked against the ALREADY UPDATED i-1 and the
// not yet updated i+1, so the reference's `next_path = deepcopy(path)` sweep equals an in-place sweep).  float32 arithmetic
// exactly as NumPy evaluates it on the float32 paths the reference holds: norm = sqrt(dx*dx + dy*dy) without FMA, ratio =
// float32(RRT_EPS) / dist, interpolate = from + diff * ratio (maze_env.py:151-172), `dist < RRT_EPS` and `diff < 1e-5` compared
// in float32 (NumPy 2 weak scalars).  collision_check_count increments are returned per path.
__device__ __forceinline__ float norm2_f32(float dx, float dy) { return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))); }

__global__ void __launch_bounds__(32) maze_steer_kernel(const float* __restrict__ old_path, const float* __restrict__ new_path,
                                                        const int32_t* __restrict__ path_ptr, const uint8_t* __restrict__ maps,
                                                        const int32_t* __restrict__ problem_of_path, int n_paths, float rrt_eps,
                                                        float* __restrict__ path_out, int32_t* __restrict__ n_checks_out,
                                                        int32_t* __restrict__ n_rounds_out, float* __restrict__ cost_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_paths) return;
  const int p0 = path_ptr[g], P = path_ptr[g + 1] - p0;
  const uint8_t* map = maps + (int64_t)(problem_of_path ? problem_of_path[g] : g) * (kW * kW);
  const float2* oldp = reinterpret_cast<const float2*>(old_path) + p0;
  const float2* newp = reinterpret_cast<const float2*>(new_path) + p0;
  float2* path = reinterpret_cast<float2*>(path_out) + p0;
  float kmax = 0.f;                                    // K = ceil(max ||old - new|| / RRT_EPS)   (:195)
  for (int i = 0; i < P; ++i) {
    const float2 a = oldp[i], b = newp[i];
    path[i] = a;                                       // path = deepcopy(old_path)
    kmax = fmaxf(kmax, __fdiv_rn(norm2_f32(a.x - b.x, a.y - b.y), rrt_eps));
  }
  const int K = (int)ceilf(kmax);
  int checks = 0, rounds = 0;
  for (int r = 0; r < K; ++r) {
    float diff = 0.f;
    ++rounds;
    for (int i = 1; i + 1 < P; ++i) {
      const float2 old_n = path[i], new_n = newp[i];
      const float dist = norm2_f32(old_n.x - new_n.x, old_n.y - new_n.y);
      float2 cand;
      if (dist < rrt_eps) cand = new_n;
      else {
        const float ratio = __fdiv_rn(rrt_eps, dist);
        cand.x = __fadd_rn(old_n.x, __fmul_rn(new_n.x - old_n.x, ratio));
        cand.y = __fadd_rn(old_n.y, __fmul_rn(new_n.y - old_n.y, ratio));
      }
      const float2 pl = path[i - 1], pr = path[i + 1];
      int c1 = 0, c2 = 0;
      bool ok = edge_free<float>(map, pl.x, pl.y, cand.x, cand.y, c1);          // env._edge_fp(next_path[i-1], next_path[i])
      if (ok) ok = edge_free<float>(map, pr.x, pr.y, cand.x, cand.y, c2);       // and env._edge_fp(next_path[i+1], next_path[i])
      checks += c1 + c2;
      if (ok) {
        path[i] = cand;
        diff = __fadd_rn(diff, norm2_f32(cand.x - new_n.x, cand.y - new_n.y));
      }
    }
    if (diff < 1e-5f) break;
  }
  n_checks_out[g] = checks;
  if (n_rounds_out) n_rounds_out[g] = rounds;
  if (cost_out) {                                      // path_cost (eval_gnn.py:53-58), float32 accumulation
    float c = 0.f;
    for (int i = 0; i + 1 < P; ++i) c = __fadd_rn(c, norm2_f32(path[i + 1].x - path[i].x, path[i + 1].y - path[i].y));
    cost_out[g] = c;
  }
}

// =====================================================================================================================
// Batched rejection sampler with a counter-based RNG                       (maze_env.py:85-100, 127-135)
// =====================================================================================================================
// The reference draws np.random.uniform(-1, 1, 2) one state at a time from the GLOBAL NumPy stream until n states are free;
// rejected draws are kept as `collided`.  Its stream cannot be reproduced in parallel (every draw is coupled to the previous
// check), so this sampler is a NEW stream with the same semantics: draw k of problem p is Philox4x32-10(key = seed,
// counter = (k, stream id of p)) -> two doubles uniform in [-1, 1) (53-bit), and the result is the PREFIX of that fixed
// sequence up to its n-th free draw -- independent of how many lanes evaluate it.  One warp per problem, 32 draws per step,
// ballot + prefix popcount keep the draw order.  Parity with the reference is distributional (tests: free / collided
// classification exact, prefix property, free fraction = free area of the map).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(32) maze_sample_kernel(const uint8_t* __restrict__ maps, const int32_t* __restrict__ problem_of_slot,
                                                         const int64_t* __restrict__ stream_of_slot, const int64_t* __restrict__ first_draw,
                                                         int n_slots, int n_want, int cap_collided, uint64_t seed, double* __restrict__ free_out,
                                                         double* __restrict__ collided_out, int32_t* __restrict__ n_collided_out,
                                                         int64_t* __restrict__ n_draws_out) {
  const int sidx = blockIdx.x;
  if (sidx >= n_slots) return;
  const int lane = threadIdx.x;
  const uint8_t* map = maps + (int64_t)problem_of_slot[sidx] * (kW * kW);
  const uint64_t stream = (uint64_t)stream_of_slot[sidx];
  uint64_t k = first_draw ? (uint64_t)first_draw[sidx] : 0;     // continue a stream across resampling rounds
  const uint64_t k_begin = k;
  double* fo = free_out + (size_t)sidx * n_want * 2;
  double* co = collided_out + (size_t)sidx * cap_collided * 2;
  int nf = 0, nc = 0;
  while (nf < n_want) {
    const uint64_t kk = k + lane;
    uint32_t r[4];
    philox4x32_10((uint32_t)kk, (uint32_t)(kk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double x = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    const double y = (double)((((uint64_t)r[2] << 32) | r[3]) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    const bool ok = cell_free(map, x, y);                       // every draw is inside the limits: one counted lookup
    const uint32_t fm = __ballot_sync(0xffffffffu, ok);
    const int before_f = __popc(fm & ((1u << lane) - 1u));
    // draws after the one that completes the n_want-th free state are not consumed
    const int need = n_want - nf;
    int last = 31;                                              // last consumed lane of this step
    if (__popc(fm) >= need) {
      uint32_t m = fm;
      for (int i = 1; i < need; ++i) m &= m - 1;                // clear the first need-1 set bits
      last = __ffs(m) - 1;
    }
    if (lane <= last) {
      if (ok) { fo[2 * (nf + before_f)] = x; fo[2 * (nf + before_f) + 1] = y; }
      else {
        const int j = nc + (lane - before_f);
        if (j < cap_collided) { co[2 * j] = x; co[2 * j + 1] = y; }
      }
    }
    const uint32_t used = last == 31 ? 0xffffffffu : ((2u << last) - 1u);
    nf += __popc(fm & used);
    nc += __popc(~fm & used);
    k += (uint64_t)(last + 1);
  }
  if (lane == 0) {
    n_collided_out[sidx] = nc;
    n_draws_out[sidx] = (int64_t)(k - k_begin);
  }
}

// =====================================================================================================================
// 3-D stick maze: MazeEnv(dim=3)                                    (maze_env.py:245-264, 279-291, 327-347)
// =====================================================================================================================
// state = (x, y, theta): a stick of length STICK_LENGTH = 0.2 centred at (x, y) at angle theta / 0.4 * pi.  The stick's end
// points and everything derived from them are float64 in the reference whatever the state dtype (a float32 scalar divided by
// the float64 LIMITS[2] promotes); poses along an edge are interpolated in the state dtype.  One thread per state / edge,
// same DFS order and the same collision_check_count / env.k side effects as oracle/maze.c (pinned by
// tests/golden/maze3_collision.npz, produced by the reference module itself).
constexpr double kLim2 = 8. * 5e-2;                 // LIMITS[2] = 8 * RRT_EPS (env_config.py)
constexpr double kStickHalf = (1.5 * 2 / 15) / 2.;  // STICK_LENGTH / 2

// 2-D float64 _edge_fp that also counts its midpoints (env.k); same traversal as edge_free<double>
__device__ bool edge2_k(const uint8_t* __restrict__ map, double ax, double ay, double bx, double by, int& cnt, int& k, bool endpoints) {
  if (endpoints) {
    k = 0;
    if (!in_range(ax, ay) || !in_range(bx, by)) return false;
    cnt += 1;
    if (!cell_free(map, ax, ay)) return false;
    cnt += 1;
    if (!cell_free(map, bx, by)) return false;
  }
  double lx = ax, ly = ay, rx = bx, ry = by;
  double sx[kStack], sy[kStack], tx[kStack], ty[kStack];
  int sp = 0;
  while (true) {
    const int dc = abs(cell_of(lx) - cell_of(rx)) + abs(cell_of(ly) - cell_of(ry));
    const double l1 = fabs(lx - rx) + fabs(ly - ry);
    if (dc > 1 && l1 > 5e-2 && sp < kStack) {
      const double mx = (lx + rx) / 2.0, my = (ly + ry) / 2.0;
      ++k;
      if (!in_range(mx, my)) return false;     // (a midpoint of two in-range points is in range; kept for symmetry with the reference)
      ++cnt;
      if (!cell_free(map, mx, my)) return false;
      sx[sp] = mx; sy[sp] = my; tx[sp] = rx; ty[sp] = ry;
      ++sp;
      rx = mx; ry = my;
      continue;
    }
    if (sp == 0) return true;
    --sp;
    lx = sx[sp]; ly = sy[sp]; rx = tx[sp]; ry = ty[sp];
  }
}

// explicit roundings: NumPy evaluates a * b + c as two operations; the compiler must not contract them into an FMA here
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T>
__device__ __forceinline__ bool valid3(T x, T y, T th) {
  return in_range(x, y) && (double)th >= -kLim2 && (double)th <= kLim2;
}
template <typename T>
__device__ __forceinline__ void end_points(T x, T y, T th, double& ax, double& ay, double& bx, double& by) {
  const double theta = (double)th / kLim2 * 3.141592653589793;
  const double ox = cos(theta), oy = sin(theta);
  const double hx = __dmul_rn(kStickHalf, ox), hy = __dmul_rn(kStickHalf, oy);      // l / 2. * orient, then centre -/+ that
  ax = __dsub_rn((double)x, hx); ay = __dsub_rn((double)y, hy);
  bx = __dadd_rn((double)x, hx); by = __dadd_rn((double)y, hy);
}
// _stick_in_free_space (:279-291)
template <typename T>
__device__ bool stick_free(const uint8_t* __restrict__ map, T x, T y, T th, int& cnt, int& k) {
  k = 0;
  if (!valid3(x, y, th)) return false;
  double ax, ay, bx, by;
  end_points(x, y, th, ax, ay, bx, by);
  if (!in_range(ax, ay)) return false;
  cnt += 1;
  if (!cell_free(map, ax, ay)) return false;
  if (!in_range(bx, by)) return false;
  cnt += 1;
  if (!cell_free(map, bx, by)) return false;
  return edge2_k(map, ax, ay, bx, by, cnt, k, false);
}

template <typename T> __device__ __forceinline__ T sqrt_rn(T x);
template <> __device__ __forceinline__ float sqrt_rn<float>(float x) { return __fsqrt_rn(x); }
template <> __device__ __forceinline__ double sqrt_rn<double>(double x) { return __dsqrt_rn(x); }

template <typename T>
__global__ void __launch_bounds__(128) maze3_state_kernel(const T* __restrict__ states, const uint8_t* __restrict__ maps,
                                                          const int32_t* __restrict__ problem, int64_t n, uint8_t* __restrict__ free_out,
                                                          int32_t* __restrict__ n_checks_out, int32_t* __restrict__ k_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * (kW * kW);
    int cnt = 0, k = 0;
    const bool ok = stick_free<T>(map, states[3 * i], states[3 * i + 1], states[3 * i + 2], cnt, k);
    free_out[i] = ok ? 1 : 0;
    if (n_checks_out) n_checks_out[i] = cnt;
    if (k_out) k_out[i] = k;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) maze3_edge_kernel(const T* __restrict__ a, const T* __restrict__ b, const uint8_t* __restrict__ maps,
                                                         const int32_t* __restrict__ problem, int64_t n, uint8_t* __restrict__ free_out,
                                                         int32_t* __restrict__ n_checks_out, int32_t* __restrict__ k_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * (kW * kW);
    const T s0 = a[3 * i], s1 = a[3 * i + 1], s2 = a[3 * i + 2], t0 = b[3 * i], t1 = b[3 * i + 1], t2 = b[3 * i + 2];
    int cnt = 0, k = 0;
    bool ok = valid3(s0, s1, s2) && valid3(t0, t1, t2);                                            // :320
    if (ok) ok = stick_free<T>(map, s0, s1, s2, cnt, k) && stick_free<T>(map, t0, t1, t2, cnt, k);   // :322
    if (ok) {
      const T d0 = t0 - s0, d1 = t1 - s1;
      T d2 = t2 - s2;
      if (fabs((double)d2) > kLim2) d2 = (T)((double)d2 > 0 ? (double)d2 - 2 * kLim2 : (double)d2 + 2 * kLim2);   // :329-333
      // distance() (:137-149): |diff| per component, the theta component wrapped (computed in float64, stored in T)
      const T a0 = d0 < 0 ? -d0 : d0, a1 = d1 < 0 ? -d1 : d1;
      T a2 = (t2 - s2) < 0 ? -(t2 - s2) : (t2 - s2);
      { const double w = fabs((double)a2 - 2 * kLim2); a2 = (T)((double)a2 < w ? (double)a2 : w); }
      const T dist = sqrt_rn<T>(add_rn(add_rn(mul_rn(a0, a0), mul_rn(a1, a1)), mul_rn(a2, a2)));
      const int K = (int)(dist / (T)0.015);                                                          // :337
      for (int kk = 1; kk < K && ok; ++kk) {
        const T ratio = (T)((double)kk * 1. / (double)K);
        const T c0 = add_rn(s0, mul_rn(ratio, d0)), c1 = add_rn(s1, mul_rn(ratio, d1)), c2 = add_rn(s2, mul_rn(ratio, d2));
        double ax, ay, bx, by;
        end_points(c0, c1, c2, ax, ay, bx, by);
        ok = edge2_k(map, ax, ay, bx, by, cnt, k, true);                                             // :344-345
      }
    }
    free_out[i] = ok ? 1 : 0;
    if (n_checks_out) n_checks_out[i] = cnt;
    if (k_out) k_out[i] = k;
  }
}

// per-problem rows of the final reduction (eval_gnn.py:120-134): (problem id, success, path cost, collision checks of the search,
// speculative checks never committed, explored nodes) -- the payload of the multi-GPU all-gather
__global__ void search_rows_kernel(const int32_t* __restrict__ status, const float* __restrict__ path_cost, const int32_t* __restrict__ n_checks,
                                   const int32_t* __restrict__ n_spec, const int32_t* __restrict__ n_explored, int first_problem, int n,
                                   float* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rows[6 * i + 0] = (float)(first_problem + i);
  rows[6 * i + 1] = status[i] == 1 ? 1.f : 0.f;
  rows[6 * i + 2] = path_cost ? path_cost[i] : 0.f;
  rows[6 * i + 3] = (float)n_checks[i];
  rows[6 * i + 4] = (float)n_spec[i];
  rows[6 * i + 5] = (float)n_explored[i];
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

extern "C" int64_t gmp_tree_search_workspace_bytes(int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total) {
  if (n_graphs < 0 || n_nodes_total < 0 || n_edges_total < 0) return -1;
  Carver cv(nullptr);
  SearchArgs A;
  return tree_search_carve(cv, A, n_graphs, n_nodes_total, n_edges_total) + 256;
}

extern "C" int gmp_maze_tree_search(const float* v, const int32_t* node_ptr, const int32_t* n_free, const int64_t* edge_index,
                                    int64_t edge_row_stride, const int32_t* edge_ptr, const float* edge_logits, const double* goal,
                                    const uint8_t* maps, const int32_t* problem_of_graph, const int32_t* slot_of_graph,
                                    int64_t n_graphs, int64_t n_nodes_total, int64_t n_edges_total, int spec_k, int first_round,
                                    int32_t* explored, int32_t* n_explored, int32_t* prev, int32_t* explored_edges,
                                    int32_t* n_explored_edges, int32_t* n_checks, int32_t* n_spec_checks, int32_t* status,
                                    int32_t* path, int32_t* path_len, float* path_cost, int32_t cap_nodes,
                                    int32_t cap_explored_edges, void* workspace, int64_t workspace_bytes, void* stream) {
  GMP_REQUIRE(n_graphs >= 0 && n_nodes_total >= 0 && n_edges_total >= 0, "negative size");
  if (n_graphs == 0) return GMP_OK;
  GMP_REQUIRE(v && node_ptr && n_free && edge_ptr && goal && maps && (edge_index || n_edges_total == 0) && (edge_logits || n_edges_total == 0),
              "null pointer");
  GMP_REQUIRE(explored && n_explored && prev && explored_edges && n_explored_edges && n_checks && n_spec_checks && status && path && path_len,
              "null search-state pointer");
  GMP_REQUIRE(spec_k >= 1 && spec_k <= 32, "spec_k must be in [1, 32]");
  GMP_REQUIRE(cap_nodes >= 2 && cap_explored_edges >= 6, "search-state capacity too small");
  GMP_REQUIRE(workspace && workspace_bytes >= gmp_tree_search_workspace_bytes(n_graphs, n_nodes_total, n_edges_total),
              "workspace too small (gmp_tree_search_workspace_bytes)");
  Carver cv(workspace);
  SearchArgs A;
  tree_search_carve(cv, A, n_graphs, n_nodes_total, n_edges_total);
  A.v = v; A.dim = 2; A.node_ptr = node_ptr; A.n_free = n_free; A.edge_index = edge_index; A.row_stride = edge_row_stride; A.edge_ptr = edge_ptr;
  A.logits = edge_logits; A.goal = goal; A.slot_of_graph = slot_of_graph;
  A.spec_k = spec_k; A.first_round = first_round;
  A.explored = explored; A.n_explored = n_explored; A.prev = prev; A.elist = explored_edges; A.n_elist = n_explored_edges;
  A.n_checks = n_checks; A.n_spec = n_spec_checks; A.status = status; A.path = path; A.path_len = path_len; A.path_cost = path_cost;
  A.cap_nodes = cap_nodes; A.cap_elist = cap_explored_edges;
  tree_search_kernel<MazeSearchEnv><<<(unsigned)n_graphs, kSearchThreads, 0, static_cast<cudaStream_t>(stream)>>>(A, MazeSearchEnv::Args{maps, problem_of_graph});
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_search_result_rows(const int32_t* status, const float* path_cost, const int32_t* n_checks, const int32_t* n_spec_checks,
                                      const int32_t* n_explored, int64_t n_problems, int32_t first_problem_id, float* rows_out, void* stream) {
  GMP_REQUIRE(n_problems >= 0, "n_problems < 0");
  if (n_problems == 0) return GMP_OK;
  GMP_REQUIRE(status && n_checks && n_spec_checks && n_explored && rows_out, "null pointer");
  search_rows_kernel<<<(unsigned)((n_problems + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      status, path_cost, n_checks, n_spec_checks, n_explored, first_problem_id, (int)n_problems, rows_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze_steer_rounds(const float* old_path, const float* new_path, const int32_t* path_ptr, const uint8_t* maps,
                                     const int32_t* problem_of_path, int64_t n_paths, double rrt_eps, float* path_out,
                                     int32_t* n_checks_out, int32_t* n_rounds_out, float* path_cost_out, void* stream) {
  GMP_REQUIRE(n_paths >= 0 && rrt_eps > 0, "n_paths < 0 or rrt_eps <= 0");
  if (n_paths == 0) return GMP_OK;
  GMP_REQUIRE(old_path && new_path && path_ptr && maps && path_out && n_checks_out, "null pointer");
  maze_steer_kernel<<<(unsigned)((n_paths + 31) / 32), 32, 0, static_cast<cudaStream_t>(stream)>>>(
      old_path, new_path, path_ptr, maps, problem_of_path, (int)n_paths, (float)rrt_eps, path_out, n_checks_out, n_rounds_out, path_cost_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze_sample_points(const uint8_t* maps, const int32_t* problem_of_slot, const int64_t* stream_of_slot,
                                      const int64_t* first_draw, int64_t n_slots, int32_t n_points, int32_t cap_collided, uint64_t seed,
                                      double* free_out, double* collided_out, int32_t* n_collided_out, int64_t* n_draws_out, void* stream) {
  GMP_REQUIRE(n_slots >= 0 && n_points >= 1 && cap_collided >= 0, "bad size");
  if (n_slots == 0) return GMP_OK;
  GMP_REQUIRE(maps && problem_of_slot && stream_of_slot && free_out && collided_out && n_collided_out && n_draws_out, "null pointer");
  maze_sample_kernel<<<(unsigned)n_slots, 32, 0, static_cast<cudaStream_t>(stream)>>>(maps, problem_of_slot, stream_of_slot, first_draw, (int)n_slots,
                                                                                     n_points, cap_collided, seed, free_out, collided_out,
                                                                                     n_collided_out, n_draws_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze3_state_fp(const void* states, int dtype, const uint8_t* maps, const int32_t* problem_of_state, int64_t n,
                                  uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out, void* stream) {
  GMP_REQUIRE(n >= 0, "n < 0");
  GMP_REQUIRE(dtype == GMP_DTYPE_F32 || dtype == GMP_DTYPE_F64, "dtype must be GMP_DTYPE_F32 or GMP_DTYPE_F64");
  if (n == 0) return GMP_OK;
  GMP_REQUIRE(states && maps && free_out, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GMP_DTYPE_F32)
    maze3_state_kernel<float><<<grid_for(n, 128), 128, 0, st>>>(static_cast<const float*>(states), maps, problem_of_state, n, free_out,
                                                                n_checks_out, k_out);
  else
    maze3_state_kernel<double><<<grid_for(n, 128), 128, 0, st>>>(static_cast<const double*>(states), maps, problem_of_state, n, free_out,
                                                                 n_checks_out, k_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}

extern "C" int gmp_maze3_edge_fp(const void* a, const void* b, int dtype, const uint8_t* maps, const int32_t* problem_of_edge, int64_t n,
                                 uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out, void* stream) {
  GMP_REQUIRE(n >= 0, "n < 0");
  GMP_REQUIRE(dtype == GMP_DTYPE_F32 || dtype == GMP_DTYPE_F64, "dtype must be GMP_DTYPE_F32 or GMP_DTYPE_F64");
  if (n == 0) return GMP_OK;
  GMP_REQUIRE(a && b && maps && free_out, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GMP_DTYPE_F32)
    maze3_edge_kernel<float><<<grid_for(n, 128), 128, 0, st>>>(static_cast<const float*>(a), static_cast<const float*>(b), maps,
                                                               problem_of_edge, n, free_out, n_checks_out, k_out);
  else
    maze3_edge_kernel<double><<<grid_for(n, 128), 128, 0, st>>>(static_cast<const double*>(a), static_cast<const double*>(b), maps,
                                                                problem_of_edge, n, free_out, n_checks_out, k_out);
  GMP_LAUNCH_CHECK();
  return GMP_OK;
}
