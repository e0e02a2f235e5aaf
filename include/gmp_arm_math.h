/* gmp_arm_math.h -- the trigonometric function of the arm collision MODEL (part of its specification).
 *
 * The arm model ("arm collision") must give bit-identical booleans on the GPU and in the CPU oracle,
 * so sin/cos cannot come from two different math libraries.  This header defines them with plain IEEE double
 * add / multiply only (fdlibm-style argument reduction by pi/2 and the classic degree-13/14 kernels), written so
 * that no fused multiply-add can be formed: the CUDA side is compiled with -fmad=false and the C side with
 * -ffp-contract=off.  Valid for |x| <= ~1e4, far beyond any joint limit.
 */
#ifndef GMP_ARM_MATH_H_
#define GMP_ARM_MATH_H_

#if defined(__CUDACC__)
#define GMP_HD __host__ __device__ __forceinline__
#else
#define GMP_HD static inline
#endif

GMP_HD double gmp_k_sin(double x) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x + v * (S1 + z * r);
}

GMP_HD double gmp_k_cos(double x) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return (1.0 - 0.5 * z) + z * r;
}

GMP_HD void gmp_sincos(double x, double* s, double* c) {
  const double INV_PIO2 = 6.36619772367581382433e-01;
  const double PIO2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double PIO2_1T = 6.07710050650619224932e-11; /* pi/2 - PIO2_1 */
  const double t = x * INV_PIO2;
  const double kf = (double)(long long)(t >= 0.0 ? t + 0.5 : t - 0.5);  /* round half away from zero */
  const double r = (x - kf * PIO2_1) - kf * PIO2_1T;
  const long long k = (long long)kf;
  const double sr = gmp_k_sin(r), cr = gmp_k_cos(r);
  switch ((int)(k & 3)) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

#endif /* GMP_ARM_MATH_H_ */
