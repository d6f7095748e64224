/*
 * pcr_detmath.h — deterministic fp64 elementary functions (part of the arithmetic SPECIFICATION).
 *
 * Why this exists: the registration path needs sin/cos (ICP Euler update, Open3D
 * utility/Eigen.cpp TransformVector6dToMatrix4d), acos/cos (Open3D EstimateNormals.cpp
 * FastEigen3x3) and atan2/acos (Open3D Feature.cpp ComputePairFeatures).  glibc's libm and
 * CUDA's libdevice differ in the last ulp, which would make histogram bins, normals and ICP
 * transforms differ between the CPU oracle and the GPU in rare cases.  The functions below use
 * only IEEE-754 +,-,*,/,sqrt,rint in a fixed order, so gcc (-ffp-contract=off) and nvcc
 * (-fmad=false) produce bit-identical results.  Accuracy is <= 2 ulp vs libm (tests/test_detmath.py).
 *
 * Coefficients: tools/gen_detmath.py (Chebyshev interpolation, 80-digit arithmetic).
 * Both the oracle (oracle/pcr_oracle.c) and the product (3d-matching_b200/csrc) include this
 * header; neither includes the other.
 */
#ifndef PCR_DETMATH_H
#define PCR_DETMATH_H

#include <math.h>

#if defined(__CUDACC__)
#define PCR_HD __host__ __device__ __forceinline__
#else
#define PCR_HD static inline
#endif

#define PCR_PI 0x1.921fb54442d18p+1
#define PCR_PIO2 0x1.921fb54442d18p+0
#define PCR_PIO4 0x1.921fb54442d18p-1
#define PCR_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define PCR_PIO2_HI 0x1.921fb54400000p+0 /* 33 significant bits: k*HI exact for |k| < 2^20 */
#define PCR_PIO2_LO 0x1.0b4611a626331p-34
#define PCR_TAN_PIO8 0x1.a827999fcef32p-2 /* sqrt(2)-1 */

/* sin(r) for |r| <= pi/4 (+ slack): r + r^3*S(r^2) */
PCR_HD double pcr_sin_kernel(double r) {
    const double z = r * r;
    double s = -0x1.ab17a79237a19p-41;
    s = s * z + 0x1.61217ec01749dp-33;
    s = s * z + -0x1.ae64541266378p-26;
    s = s * z + 0x1.71de3a54605eep-19;
    s = s * z + -0x1.a01a01a019936p-13;
    s = s * z + 0x1.1111111111110p-7;
    s = s * z + -0x1.5555555555555p-3;
    return r + (r * z) * s;
}

/* cos(r) for |r| <= pi/4 (+ slack): 1 - z/2 + z^2*C(z) */
PCR_HD double pcr_cos_kernel(double r) {
    const double z = r * r;
    double c = 0x1.ab783376962cfp-45;
    c = c * z + -0x1.9394b9c9c20a4p-37;
    c = c * z + 0x1.1eed8deb6d561p-29;
    c = c * z + -0x1.27e4fb7712bdfp-22;
    c = c * z + 0x1.a01a01a019d0ap-16;
    c = c * z + -0x1.6c16c16c16c16p-10;
    c = c * z + 0x1.5555555555555p-5;
    return (1.0 - 0.5 * z) + (z * z) * c;
}

/* sin and cos of x (intended range |x| < 1e5; Cody-Waite two-term reduction). */
PCR_HD void pcr_sincos(double x, double *s, double *c) {
    const double kd = rint(x * PCR_TWO_OVER_PI);
    const double r = (x - kd * PCR_PIO2_HI) - kd * PCR_PIO2_LO;
    const long long k = (long long)kd;
    const double sk = pcr_sin_kernel(r);
    const double ck = pcr_cos_kernel(r);
    switch ((int)(k & 3)) {
        case 0: *s = sk; *c = ck; break;
        case 1: *s = ck; *c = -sk; break;
        case 2: *s = -sk; *c = -ck; break;
        default: *s = -ck; *c = sk; break;
    }
}
PCR_HD double pcr_sin(double x) { double s, c; pcr_sincos(x, &s, &c); return s; }
PCR_HD double pcr_cos(double x) { double s, c; pcr_sincos(x, &s, &c); return c; }

/* atan(t) for |t| <= tan(pi/8) (+ slack): t + t*z*A(z) */
PCR_HD double pcr_atan_kernel(double t) {
    const double z = t * t;
    double a = 0x1.87d01190ba280p-7;
    a = a * z + -0x1.bed41c47c5a20p-6;
    a = a * z + 0x1.31038a2d9a394p-5;
    a = a * z + -0x1.5fc8a61e87785p-5;
    a = a * z + 0x1.857f8346cd3fdp-5;
    a = a * z + -0x1.af19a4df9619cp-5;
    a = a * z + 0x1.e1e0de4023a93p-5;
    a = a * z + -0x1.11110ad3458eep-4;
    a = a * z + 0x1.3b13b106be071p-4;
    a = a * z + -0x1.745d1744b460ap-4;
    a = a * z + 0x1.c71c71c718cd9p-4;
    a = a * z + -0x1.2492492492460p-3;
    a = a * z + 0x1.9999999999999p-3;
    a = a * z + -0x1.5555555555555p-2;
    return t + (t * z) * a;
}

/* atan2(y, x).  atan2(0,0) = 0; signed zeros are treated as +0; NaN in -> NaN out. */
PCR_HD double pcr_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    if (ax == 0.0 && ay == 0.0) return 0.0;
    const int swap = ay > ax;
    double t = swap ? ax / ay : ay / ax; /* in [0,1], NaN propagates */
    double base = 0.0;
    if (t > PCR_TAN_PIO8) {
        t = (t - 1.0) / (t + 1.0);
        base = PCR_PIO4;
    }
    double a = base + pcr_atan_kernel(t);
    if (swap) a = PCR_PIO2 - a;
    if (x < 0.0) a = PCR_PI - a;
    if (y < 0.0) a = -a;
    return a;
}

/* acos(x) = 2*atan2(sqrt(1-x), sqrt(1+x)); |x| > 1 -> NaN (as libm). */
PCR_HD double pcr_acos(double x) {
    return 2.0 * pcr_atan2(sqrt(1.0 - x), sqrt(1.0 + x));
}

#endif /* PCR_DETMATH_H */
