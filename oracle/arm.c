/* ORACLE -- test infrastructure, never the product path.
 *
 * Plain-C restatement of the ARM collision check as THIS repository specifies it ("arm collision").
 * The reference decides arm collisions with PyBullet (environment/kuka_env.py:354-370,
 * environment/kuka_2arm_env.py:357-369): resetJointState x DoF, performCollisionDetection, "free iff no
 * contact point".  PyBullet is absent from /root/reference, not installable here, and its version is unpinned
 * by the reference => PARITY UNPINNED for the booleans themselves.  What IS restated from the reference, line by
 * line, is everything around the contact query:
 *   _valid_state            kuka_env.py:350-352     joint limits, compared in double
 *   _point_in_free_space    kuka_env.py:354-370     out of limits -> False, not counted; else count += 1
 *   _edge_fp                kuka_env.py:389-411     endpoints valid, endpoints free (short circuit), d = ||b - a||_2
 *                                                   in the input dtype, K = int(d / RRT_EPS), states
 *                                                   a + (k/K)(b - a), k = 0..K-1, first collision -> False
 *   Kuka2Env.set_config     kuka_2arm_env.py:167-174  config[0:7] -> arm at x=-0.5, config[7:14] -> arm at x=+0.5
 * and the contact query is replaced by the sphere model of csrc/arm_models_data.h (joint chain from the
 * reference URDFs, links covered by spheres fitted to the reference STL meshes; free iff every sphere is farther
 * than GMP_ARM_MARGIN from every box and from every sphere of the other arm).
 * All model arithmetic is IEEE double without FMA (build with -ffp-contract=off) so that the CUDA kernel can be
 * compared bit for bit.
 */
#include <math.h>
#include <stdint.h>

#include "../gnn_motion_planning_b200/csrc/arm_models_data.h"
#include "../include/gmp_arm_math.h"

typedef struct {
  int n_arms, dof_per_arm, n_spheres;
  int n_joints;            /* chain length per arm (== dof_per_arm except for the snake's planar base) */
  int self_min_diff;       /* self collision between links whose frames differ by at least this much */
  double lo[GMP_ARM_MAX_JOINTS], hi[GMP_ARM_MAX_JOINTS]; /* limits of the STATE components of one arm */
  const GmpJoint* joints;
  const GmpSphere* spheres;
  double base_x[2];
  int self_collision;      /* URDF_USE_SELF_COLLISION (ur5_env.py:107): links not directly connected may collide */
  int plane_exempt_frame;  /* ground plane z = 0 (ur5_env.py:108-111), -2 = no plane; this frame's link is filtered out */
} ArmModel;

static ArmModel get_model(int id) {
  ArmModel m;
  m.base_x[0] = m.base_x[1] = 0.0;
  m.self_collision = 0;
  m.self_min_diff = 2;
  m.plane_exempt_frame = -2;
  if (id == GMP_ARM_SNAKE7) {
    /* snake_env.py:90: URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_INCLUDE_PARENT */
    m.n_arms = 1; m.dof_per_arm = 7; m.joints = gmp_snake7_joints; m.spheres = gmp_snake7_spheres;
    m.n_spheres = (int)(sizeof(gmp_snake7_spheres) / sizeof(GmpSphere));
    m.self_collision = 1;
    m.self_min_diff = 1;
  } else if (id == GMP_ARM_UR5) {
    m.n_arms = 1; m.dof_per_arm = 6; m.joints = gmp_ur5_joints; m.spheres = gmp_ur5_spheres;
    m.n_spheres = (int)(sizeof(gmp_ur5_spheres) / sizeof(GmpSphere));
    m.self_collision = 1;
    m.plane_exempt_frame = 0;
  } else if (id == GMP_ARM_KUKA13) {
    m.n_arms = 1; m.dof_per_arm = 13; m.joints = gmp_kuka13_joints; m.spheres = gmp_kuka13_spheres;
    m.n_spheres = (int)(sizeof(gmp_kuka13_spheres) / sizeof(GmpSphere));
  } else {
    m.n_arms = id == GMP_ARM_KUKA14 ? 2 : 1; m.dof_per_arm = 7; m.joints = gmp_kuka7_joints; m.spheres = gmp_kuka7_spheres;
    m.n_spheres = (int)(sizeof(gmp_kuka7_spheres) / sizeof(GmpSphere));
    if (id == GMP_ARM_KUKA14) { m.base_x[0] = -0.5; m.base_x[1] = 0.5; }
  }
  m.n_joints = m.dof_per_arm;
  for (int j = 0; j < m.dof_per_arm; ++j) {
    m.lo[j] = id == GMP_ARM_SNAKE7 ? gmp_snake7_lo[j] : m.joints[j].lo;
    m.hi[j] = id == GMP_ARM_SNAKE7 ? gmp_snake7_hi[j] : m.joints[j].hi;
  }
  return m;
}

int oracle_arm_dof(int id) { ArmModel m = get_model(id); return m.n_arms * m.dof_per_arm; }

void oracle_arm_limits(int id, double* lo, double* hi) {
  ArmModel m = get_model(id);
  for (int a = 0; a < m.n_arms; ++a)
    for (int j = 0; j < m.dof_per_arm; ++j) { lo[a * m.dof_per_arm + j] = m.lo[j]; hi[a * m.dof_per_arm + j] = m.hi[j]; }
}

/* world sphere centres of one arm: out[3 * s] */
static void arm_spheres_world(const ArmModel* m, const double* q, double base_x, double* out) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {base_x, 0.0, 0.0};
  int s = 0;
  for (int f = 0; f <= m->n_joints; ++f) {
    if (f > 0) {
      const GmpJoint* J = &m->joints[f - 1];
      const double qj = q[J->qidx];
      /* p += R * t ; R = R * Rj */
      double np_[3], R1[9], Rq[9], R2[9];
      for (int i = 0; i < 3; ++i) np_[i] = p[i] + ((R[3 * i] * J->t[0] + R[3 * i + 1] * J->t[1]) + R[3 * i + 2] * J->t[2]);
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k)
          R1[3 * i + k] = (R[3 * i] * J->R[k] + R[3 * i + 1] * J->R[3 + k]) + R[3 * i + 2] * J->R[6 + k];
      if (J->type == 1) { /* prismatic: translate along the joint axis, orientation unchanged */
        const double v0 = J->axis[0] * qj, v1 = J->axis[1] * qj, v2 = J->axis[2] * qj;
        for (int i = 0; i < 3; ++i) p[i] = np_[i] + ((R1[3 * i] * v0 + R1[3 * i + 1] * v1) + R1[3 * i + 2] * v2);
        for (int i = 0; i < 9; ++i) R[i] = R1[i];
      } else {
        double sn, cs;
        gmp_sincos(qj, &sn, &cs);
        const double ax = J->axis[0], ay = J->axis[1], az = J->axis[2], oc = 1.0 - cs;
        Rq[0] = cs + oc * (ax * ax); Rq[1] = oc * (ax * ay) - sn * az; Rq[2] = oc * (ax * az) + sn * ay;
        Rq[3] = oc * (ay * ax) + sn * az; Rq[4] = cs + oc * (ay * ay); Rq[5] = oc * (ay * az) - sn * ax;
        Rq[6] = oc * (az * ax) - sn * ay; Rq[7] = oc * (az * ay) + sn * ax; Rq[8] = cs + oc * (az * az);
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            R2[3 * i + k] = (R1[3 * i] * Rq[k] + R1[3 * i + 1] * Rq[3 + k]) + R1[3 * i + 2] * Rq[6 + k];
        for (int i = 0; i < 9; ++i) R[i] = R2[i];
        for (int i = 0; i < 3; ++i) p[i] = np_[i];
      }
    }
    while (s < m->n_spheres && m->spheres[s].frame == f) {
      const double* c = m->spheres[s].c;
      for (int i = 0; i < 3; ++i) out[3 * s + i] = p[i] + ((R[3 * i] * c[0] + R[3 * i + 1] * c[1]) + R[3 * i + 2] * c[2]);
      ++s;
    }
  }
}

static int config_collides(const ArmModel* m, const double* q, const double* boxes, int n_boxes) {
  double w[2][3 * GMP_ARM_MAX_SPHERES];
  for (int a = 0; a < m->n_arms; ++a) arm_spheres_world(m, q + a * m->dof_per_arm, m->base_x[a], w[a]);
  for (int a = 0; a < m->n_arms; ++a)
    for (int s = 0; s < m->n_spheres; ++s) {
      const double rr = m->spheres[s].r + GMP_ARM_MARGIN;
      for (int b = 0; b < n_boxes; ++b) {
        const double* bx = boxes + 6 * b; /* half extents, centre */
        double d2 = 0.0;
        for (int i = 0; i < 3; ++i) {
          double d = fabs(w[a][3 * s + i] - bx[3 + i]) - bx[i];
          if (d < 0.0) d = 0.0;
          d2 = d2 + d * d;
        }
        if (d2 <= rr * rr) return 1;
      }
    }
  if (m->plane_exempt_frame != -2)
    for (int s = 0; s < m->n_spheres; ++s) {
      if (m->spheres[s].frame == m->plane_exempt_frame) continue;
      if (w[0][3 * s + 2] <= m->spheres[s].r + GMP_ARM_MARGIN) return 1;
    }
  if (m->self_collision)
    for (int s = 0; s < m->n_spheres; ++s)
      for (int t = s + 1; t < m->n_spheres; ++t) {
        if (m->spheres[t].frame - m->spheres[s].frame < m->self_min_diff) continue; /* same / (ur5) directly connected links */
        const double rr = (m->spheres[s].r + m->spheres[t].r) + GMP_ARM_MARGIN;
        double d2 = 0.0;
        for (int i = 0; i < 3; ++i) {
          const double d = w[0][3 * s + i] - w[0][3 * t + i];
          d2 = d2 + d * d;
        }
        if (d2 <= rr * rr) return 1;
      }
  if (m->n_arms == 2)
    for (int s = 0; s < m->n_spheres; ++s)
      for (int t = 0; t < m->n_spheres; ++t) {
        const double rr = (m->spheres[s].r + m->spheres[t].r) + GMP_ARM_MARGIN;
        double d2 = 0.0;
        for (int i = 0; i < 3; ++i) {
          const double d = w[0][3 * s + i] - w[1][3 * t + i];
          d2 = d2 + d * d;
        }
        if (d2 <= rr * rr) return 1;
      }
  return 0;
}

static int state_valid(const ArmModel* m, const double* q) {
  for (int a = 0; a < m->n_arms; ++a)
    for (int j = 0; j < m->dof_per_arm; ++j) {
      const double x = q[a * m->dof_per_arm + j];
      if (!(x >= m->lo[j] && x <= m->hi[j])) return 0;
    }
  return 1;
}

/* _point_in_free_space: 1 free / 0 not; *cnt += 1 iff within limits */
static int point_free(const ArmModel* m, const double* q, const double* boxes, int n_boxes, int* cnt) {
  if (!state_valid(m, q)) return 0;
  *cnt += 1;
  return !config_collides(m, q, boxes, n_boxes);
}

#define DEFINE_ARM(T, SUF, SQRT)                                                                                     \
  void oracle_arm_state_fp_##SUF(int model, const T* states, const double* boxes, const int32_t* box_ptr,            \
                                 const int32_t* problem, int64_t n, uint8_t* free_out, uint8_t* counted_out) {       \
    ArmModel m = get_model(model);                                                                                   \
    const int dof = m.n_arms * m.dof_per_arm;                                                                        \
    for (int64_t i = 0; i < n; ++i) {                                                                                \
      double q[GMP_ARM_MAX_JOINTS];                                                                                  \
      for (int j = 0; j < dof; ++j) q[j] = (double)states[i * dof + j];                                              \
      const int pr = problem ? problem[i] : 0;                                                                       \
      int cnt = 0;                                                                                                   \
      free_out[i] = (uint8_t)point_free(&m, q, boxes + 6 * (int64_t)box_ptr[pr], box_ptr[pr + 1] - box_ptr[pr], &cnt); \
      if (counted_out) counted_out[i] = (uint8_t)cnt;                                                                \
    }                                                                                                                \
  }                                                                                                                  \
  void oracle_arm_edge_fp_##SUF(int model, const T* a, const T* b, const double* boxes, const int32_t* box_ptr,      \
                                const int32_t* problem, int64_t n, double rrt_eps, uint8_t* free_out,                \
                                int32_t* n_checks_out) {                                                             \
    ArmModel m = get_model(model);                                                                                   \
    const int dof = m.n_arms * m.dof_per_arm;                                                                        \
    for (int64_t i = 0; i < n; ++i) {                                                                                \
      const T* s = a + i * dof;                                                                                      \
      const T* t = b + i * dof;                                                                                      \
      const int pr = problem ? problem[i] : 0;                                                                       \
      const double* bx = boxes + 6 * (int64_t)box_ptr[pr];                                                           \
      const int nb = box_ptr[pr + 1] - box_ptr[pr];                                                                  \
      double qs[GMP_ARM_MAX_JOINTS], qt[GMP_ARM_MAX_JOINTS], qc[GMP_ARM_MAX_JOINTS];                                 \
      for (int j = 0; j < dof; ++j) { qs[j] = (double)s[j]; qt[j] = (double)t[j]; }                                  \
      int cnt = 0, ok = 1;                                                                                           \
      if (!state_valid(&m, qs) || !state_valid(&m, qt)) ok = 0;                           /* kuka_env.py:394 */      \
      else if (!point_free(&m, qs, bx, nb, &cnt) || !point_free(&m, qt, bx, nb, &cnt)) ok = 0; /* :397 */             \
      else {                                                                                                         \
        /* d = KukaEnv.distance (kuka_env.py:224-233): to_state is first clamped against the FLOAT64 pose_range, which promotes   */ \
        /* the whole expression to float64 whatever the input dtype (a no-op in value: both states are within the limits);    */ \
        /* numpy's summation order (8-way unrolled pairwise for n >= 8); K = int(d / RRT_EPS) in float64          (:403)      */ \
        double sq[GMP_ARM_MAX_JOINTS], d2;                                                                           \
        for (int j = 0; j < dof; ++j) { double df = (double)t[j] - (double)s[j]; df = df < 0 ? -df : df; sq[j] = df * df; } \
        if (dof < 8) { d2 = 0; for (int j = 0; j < dof; ++j) d2 = d2 + sq[j]; }                                      \
        else {                                                                                                       \
          d2 = ((sq[0] + sq[1]) + (sq[2] + sq[3])) + ((sq[4] + sq[5]) + (sq[6] + sq[7]));                            \
          for (int j = 8; j < dof; ++j) d2 = d2 + sq[j];                                                             \
        }                                                                                                            \
        const double d = sqrt(d2);                                                                                   \
        const int K = (int)(d / rrt_eps);                                                                            \
        for (int k = 0; k < K && ok; ++k) {                                               /* :404-409 */             \
          const T ratio = (T)((double)k * 1. / (double)K);                                                           \
          for (int j = 0; j < dof; ++j) { const T step = ratio * (t[j] - s[j]); qc[j] = (double)(T)(s[j] + step); }  \
          if (!point_free(&m, qc, bx, nb, &cnt)) ok = 0;                                                             \
        }                                                                                                            \
      }                                                                                                              \
      free_out[i] = (uint8_t)ok;                                                                                     \
      if (n_checks_out) n_checks_out[i] = cnt;                                                                       \
    }                                                                                                                \
  }

DEFINE_ARM(float, f32, sqrtf)
DEFINE_ARM(double, f64, sqrt)
