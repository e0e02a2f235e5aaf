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

/* ------------------------------------------------------------------------------------------------------------------
 * 3-D stick maze (MazeEnv(dim=3)): state = (x, y, theta), a stick of length STICK_LENGTH = 0.2 centred at (x, y)
 *   _end_points            maze_env.py:245-264   theta' = theta / LIMITS[2] * pi (float64: a float32 scalar divided by a float64
 *                                                scalar promotes), end points centre -/+ 0.1 (cos, sin)(theta') in float64
 *   _stick_in_free_space   :279-291              valid state, both end points free (2-D float64 point checks), bisection of the stick
 *   _edge_fp (3-D branch)  :327-347              theta displacement wrapped into [-0.4, 0.4]; d = distance() (:137-149, theta
 *                                                component wrapped); K = int(d / 0.015); poses k = 1..K-1 at state + k/K * disp in
 *                                                the input dtype; each pose's stick checked as a 2-D float64 edge
 * env.k: _edge_fp / _stick_in_free_space reset it; the nested 2-D _edge_fp call of every pose resets it again, so after a 3-D
 * edge check it holds the midpoint count of the LAST stick checked.
 * Pinned by tests/golden/maze3_collision.npz (the reference module itself, float32 and float64 states). */
#define LIM2_D (8. * RRT_EPS_D)
#define STICK_HALF_D ((1.5 * 2 / 15) / 2.)

static int seg_k_f64(const double* l, const double* r, const uint8_t* map, int* cnt, int* k) {
  int lc0 = cell_f64(l[0]), lc1 = cell_f64(l[1]), rc0 = cell_f64(r[0]), rc1 = cell_f64(r[1]);
  int dc = (lc0 > rc0 ? lc0 - rc0 : rc0 - lc0) + (lc1 > rc1 ? lc1 - rc1 : rc1 - lc1);
  double d0 = fabs(l[0] - r[0]), d1 = fabs(l[1] - r[1]);
  if (dc > 1 && d0 + d1 > RRT_EPS_D) {
    double mid[2] = {(l[0] + r[0]) / 2.0, (l[1] + r[1]) / 2.0};
    *k += 1;
    if (!point_free_f64(mid, map, cnt)) return 0;
    return seg_k_f64(l, mid, map, cnt, k) && seg_k_f64(mid, r, map, cnt, k);
  }
  return 1;
}
/* 2-D float64 _edge_fp with its own k */
static int edge2_k_f64(const double* a, const double* b, const uint8_t* map, int* cnt, int* k) {
  *k = 0;
  if (!valid_f64(a) || !valid_f64(b)) return 0;
  if (!point_free_f64(a, map, cnt) || !point_free_f64(b, map, cnt)) return 0;
  return seg_k_f64(a, b, map, cnt, k);
}

#define DEFINE_MAZE3(T, SUF, SQRT)                                                                                   \
  static inline int valid3_##SUF(const T* s) {                                                                       \
    return (double)s[0] >= -1.0 && (double)s[0] <= 1.0 && (double)s[1] >= -1.0 && (double)s[1] <= 1.0 &&             \
           (double)s[2] >= -LIM2_D && (double)s[2] <= LIM2_D;                                                        \
  }                                                                                                                  \
  static inline void end_points_##SUF(const T* c, double* a, double* b) {                                            \
    const double theta = (double)c[2] / LIM2_D * M_PI;                                                               \
    const double ox = cos(theta), oy = sin(theta);                                                                   \
    a[0] = (double)c[0] - STICK_HALF_D * ox; a[1] = (double)c[1] - STICK_HALF_D * oy;                                \
    b[0] = (double)c[0] + STICK_HALF_D * ox; b[1] = (double)c[1] + STICK_HALF_D * oy;                                \
  }                                                                                                                  \
  static int stick_free_##SUF(const T* s, const uint8_t* map, int* cnt, int* k) {                                    \
    double a[2], b[2];                                                                                               \
    *k = 0;                                                                                                          \
    if (!valid3_##SUF(s)) return 0;                                                                                  \
    end_points_##SUF(s, a, b);                                                                                       \
    if (!point_free_f64(a, map, cnt) || !point_free_f64(b, map, cnt)) return 0;                                      \
    return seg_k_f64(a, b, map, cnt, k);                                                                             \
  }                                                                                                                  \
  void oracle_maze3_state_fp_##SUF(const T* states, const uint8_t* maps, const int32_t* problem, int64_t n,          \
                                   uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out) {                       \
    for (int64_t i = 0; i < n; ++i) {                                                                                \
      int cnt = 0, k = 0;                                                                                            \
      const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * MAZE_W * MAZE_W;                             \
      free_out[i] = (uint8_t)stick_free_##SUF(states + 3 * i, map, &cnt, &k);                                        \
      if (n_checks_out) n_checks_out[i] = cnt;                                                                       \
      if (k_out) k_out[i] = k;                                                                                       \
    }                                                                                                                \
  }                                                                                                                  \
  void oracle_maze3_edge_fp_##SUF(const T* A, const T* B, const uint8_t* maps, const int32_t* problem, int64_t n,    \
                                  uint8_t* free_out, int32_t* n_checks_out, int32_t* k_out) {                        \
    for (int64_t i = 0; i < n; ++i) {                                                                                \
      int cnt = 0, k = 0, ok = 1;                                                                                    \
      const uint8_t* map = maps + (int64_t)(problem ? problem[i] : 0) * MAZE_W * MAZE_W;                             \
      const T* s = A + 3 * i;                                                                                        \
      const T* t = B + 3 * i;                                                                                        \
      if (!valid3_##SUF(s) || !valid3_##SUF(t)) ok = 0;                                      /* :320 */              \
      else if (!stick_free_##SUF(s, map, &cnt, &k) || !stick_free_##SUF(t, map, &cnt, &k)) ok = 0;   /* :322 */      \
      else {                                                                                                         \
        T disp[3] = {(T)(t[0] - s[0]), (T)(t[1] - s[1]), (T)(t[2] - s[2])};                                          \
        if (fabs((double)disp[2]) > LIM2_D)                                                  /* :329-333 */          \
          disp[2] = (T)((double)disp[2] > 0 ? (double)disp[2] - 2 * LIM2_D : (double)disp[2] + 2 * LIM2_D);          \
        T diff[3];                                                                           /* distance(), :137-149 */ \
        for (int j = 0; j < 3; ++j) { T d_ = (T)(t[j] - s[j]); diff[j] = d_ < 0 ? -d_ : d_; }                        \
        {                                                                                                            \
          const double w = fabs((double)diff[2] - 2 * LIM2_D);                                                       \
          diff[2] = (T)((double)diff[2] < w ? (double)diff[2] : w);                                                  \
        }                                                                                                            \
        const T d = SQRT((T)((T)((T)(diff[0] * diff[0]) + (T)(diff[1] * diff[1])) + (T)(diff[2] * diff[2])));        \
        const int K = (int)(d / (T)0.015);                                                   /* :337 */              \
        for (int kk = 1; kk < K && ok; ++kk) {                                                                       \
          const T ratio = (T)((double)kk * 1. / (double)K);                                                          \
          T c[3];                                                                                                    \
          for (int j = 0; j < 3; ++j) { const T step = (T)(ratio * disp[j]); c[j] = (T)(s[j] + step); }              \
          double ca[2], cb[2];                                                                                       \
          end_points_##SUF(c, ca, cb);                                                                               \
          if (!edge2_k_f64(ca, cb, map, &cnt, &k)) ok = 0;                                   /* :344-345 */          \
        }                                                                                                            \
      }                                                                                                              \
      free_out[i] = (uint8_t)ok;                                                                                     \
      if (n_checks_out) n_checks_out[i] = cnt;                                                                       \
      if (k_out) k_out[i] = k;                                                                                       \
    }                                                                                                                \
  }

DEFINE_MAZE3(float, f32, sqrtf)
DEFINE_MAZE3(double, f64, sqrt)
