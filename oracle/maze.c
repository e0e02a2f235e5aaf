/* ORACLE -- test infrastructure, never the product path.
 *
 * Plain-C restatement of the reference 2-D maze collision check, reference
 * environment/maze_env.py (MazeEnv, dim == 2):
 *   _transform                :236-239   cell = trunc((x + 1.0) * w / 2.0), clamped to w-1, in the input dtype
 *   _valid_state              :266-268   -LIMITS <= x <= LIMITS, LIMITS = [1, 1] (env_config.py:5)
 *   _point_in_free_space      :270-277   out of range -> False (not counted); else count += 1; map[cx][cy] == 0
 *   _state_fp                 :293-299
 *   _iterative_check_segment  :301-314   recursive bisection, left half first, short-circuit `and`
 *   _edge_fp                  :316-325   (2-D branch)
 *
 * dtype T in {f32, f64}: the planner re-tests float32 tensors (eval_gnn.py:215) while the sampler
 * tests float64 draws (maze_env.py:131), so both instantiations exist.  NumPy-2 (NEP 50) scalar
 * semantics are followed at maze_env.py:306: the L1 length (a T scalar) is compared with RRT_EPS
 * rounded to T.  Build with -ffp-contract=off so no FMA changes a rounding.
 *
 * Pinned against the reference's own maze_env.py (importable here) by tests/golden/make_golden.py
 * -> tests/golden/maze_*.npz and tests/test_oracle_maze.py.
 *
 * n_checks mirrors the increments of MazeEnv.collision_check_count (in-range lookups only,
 * DFS order, early exit on the first blocked lookup).
 */
#include <stdint.h>
#include <math.h>

#define MAZE_W 15
#define RRT_EPS_D 5e-2

#define DEFINE_MAZE(T, SUF, ONE, HALFDIV, WCONST, EPS)                                              \
  static inline int cell_##SUF(T x) {                                                               \
    T t = ((x + ONE) * WCONST) / HALFDIV;                                                           \
    int c = (int)t; /* C truncation toward zero == ndarray.astype(int) for in-range values */       \
    return c > MAZE_W - 1 ? MAZE_W - 1 : c;                                                         \
  }                                                                                                 \
  static inline int valid_##SUF(const T* s) {                                                       \
    return (double)s[0] >= -1.0 && (double)s[1] >= -1.0 && (double)s[0] <= 1.0 && (double)s[1] <= 1.0; \
  }                                                                                                 \
  /* returns 1 free / 0 blocked-or-invalid; *cnt incremented iff the state was in range */          \
  static inline int point_free_##SUF(const T* s, const uint8_t* map, int* cnt) {                    \
    if (!valid_##SUF(s)) return 0;                                                                  \
    *cnt += 1;                                                                                      \
    return map[cell_##SUF(s[0]) * MAZE_W + cell_##SUF(s[1])] == 0;                                  \
  }                                                                                                 \
  static int segment_##SUF(const T* l, const T* r, const uint8_t* map, int* cnt) {                  \
    int lc0 = cell_##SUF(l[0]), lc1 = cell_##SUF(l[1]);                                             \
    int rc0 = cell_##SUF(r[0]), rc1 = cell_##SUF(r[1]);                                             \
    int dc = (lc0 > rc0 ? lc0 - rc0 : rc0 - lc0) + (lc1 > rc1 ? lc1 - rc1 : rc1 - lc1);             \
    T d0 = l[0] - r[0], d1 = l[1] - r[1];                                                           \
    d0 = d0 < 0 ? -d0 : d0;                                                                         \
    d1 = d1 < 0 ? -d1 : d1;                                                                         \
    T l1 = d0 + d1;                                                                                 \
    if (dc > 1 && l1 > EPS) {                                                                       \
      T mid[2];                                                                                     \
      mid[0] = (l[0] + r[0]) / HALFDIV;                                                             \
      mid[1] = (l[1] + r[1]) / HALFDIV;                                                             \
      if (!point_free_##SUF(mid, map, cnt)) return 0;                                               \
      return segment_##SUF(l, mid, map, cnt) && segment_##SUF(mid, r, map, cnt);                    \
    }                                                                                               \
    return 1;                                                                                       \
  }                                                                                                 \
  void oracle_maze_state_fp_##SUF(const T* states, const uint8_t* maps, const int32_t* problem,     \
                                  int64_t n, uint8_t* free_out, uint8_t* counted_out) {             \
    for (int64_t i = 0; i < n; ++i) {                                                               \
      int cnt = 0;                                                                                  \
      const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * MAZE_W * MAZE_W;            \
      free_out[i] = (uint8_t)point_free_##SUF(states + 2 * i, map, &cnt);                           \
      if (counted_out) counted_out[i] = (uint8_t)cnt;                                               \
    }                                                                                               \
  }                                                                                                 \
  void oracle_maze_edge_fp_##SUF(const T* a, const T* b, const uint8_t* maps,                       \
                                 const int32_t* problem, int64_t n, uint8_t* free_out,              \
                                 int32_t* n_checks_out) {                                           \
    for (int64_t i = 0; i < n; ++i) {                                                               \
      int cnt = 0, ok;                                                                              \
      const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * MAZE_W * MAZE_W;            \
      const T* s = a + 2 * i;                                                                       \
      const T* t = b + 2 * i;                                                                       \
      if (!valid_##SUF(s) || !valid_##SUF(t)) ok = 0;                       /* :320 */              \
      else if (!point_free_##SUF(s, map, &cnt) || !point_free_##SUF(t, map, &cnt)) ok = 0; /* :322 */ \
      else ok = segment_##SUF(s, t, map, &cnt);                             /* :325 */              \
      free_out[i] = (uint8_t)ok;                                                                    \
      if (n_checks_out) n_checks_out[i] = cnt;                                                      \
    }                                                                                               \
  }

DEFINE_MAZE(float, f32, 1.0f, 2.0f, 15.0f, ((float)RRT_EPS_D))
DEFINE_MAZE(double, f64, 1.0, 2.0, 15.0, RRT_EPS_D)
