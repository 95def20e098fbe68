// cosf / sinf exactly as the reference's host libm evaluates them.  computeOrbDescriptor (src/ORBextractor.cc:109)
// calls `(float)cos(angle)` / `(float)sin(angle)` with a float argument, i.e. glibc cosf / sinf.  glibc >= 2.28 uses the
// ARM optimized-routines algorithm (sysdeps/ieee754/flt-32/s_sincosf.h): argument reduction by n = round(x * 2/pi) in
// double, then an odd / even minimax polynomial in double, one rounding to float at the end.  On x86-64 hosts with
// FMA+AVX2 the ifunc selects the FMA build, whose contraction pattern (read from the shipped libm.so.6 2.39) is
// reproduced below with explicit __fma_rn / __dmul_rn so that the device result is bit-identical:
//   reduce : r = x * hpi_inv(2^24 * 2/pi); n = ((int32)r + 0x800000) >> 24; x = fma(-n, pi/2, x)
//   sin    : s1v = fma(x2, s3, s2); x3 = x2 * xs; x5 = x2 * x3; s = fma(x3, s1, xs); ret = fma(s1v, x5, s)
//   cos    : x4 = x2 * x2; c1v = fma(x2, c1, c0); c2v = fma(x2, c4, c3); x6 = x2 * x4; c = fma(x4, c2, c1v); ret = fma(c2v, x6, c)
// Valid for |x| < 120 (the descriptor angle lies in [0, 2*pi]).  tests/test_trig_gpu.py compares every float angle in
// [0, 360] degrees against the host libm.
#pragma once
namespace ygz_b200 {

__device__ __forceinline__ double glibc_sin_poly(double xs, double x2) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double s1v = __fma_rn(x2, s3, s2);
    const double x3 = __dmul_rn(x2, xs);
    const double x5 = __dmul_rn(x2, x3);
    const double s = __fma_rn(x3, s1, xs);
    return __fma_rn(s1v, x5, s);
}

__device__ __forceinline__ double glibc_cos_poly(double x2, double sgn) {
    // sgn = +1 for table 0, -1 for table 1 (all cosine coefficients negated: exact sign flips)
    const double c0 = 1.0 * sgn, c1 = -0x1.ffffffd0c621cp-2 * sgn, c2 = 0x1.55553e1068f19p-5 * sgn,
                 c3 = -0x1.6c087e89a359dp-10 * sgn, c4 = 0x1.99343027bf8c3p-16 * sgn;
    const double x4 = __dmul_rn(x2, x2);
    const double c1v = __fma_rn(x2, c1, c0);
    const double c2v = __fma_rn(x2, c4, c3);
    const double x6 = __dmul_rn(x2, x4);
    const double c = __fma_rn(x4, c2, c1v);
    return __fma_rn(c2v, x6, c);
}

__device__ inline void ref_cos_sin(float y, float* cos_out, float* sin_out) {
    const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
    const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top <= 0x3f3) {                       // |y| < pi/4
        const double x2 = __dmul_rn(x, x);
        if (top <= 0x397) { *cos_out = 1.0f; *sin_out = y; return; }      // |y| < 2^-12
        *cos_out = __double2float_rn(glibc_cos_poly(x2, 1.0));
        *sin_out = __double2float_rn(glibc_sin_poly(x, x2));
        return;
    }
    const double r = __dmul_rn(x, hpi_inv);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, hpi, x);
    const double x2 = __dmul_rn(x, x);
    const double tsgn = (n & 2) ? -1.0 : 1.0;                             // __sincosf_table[1] when (n & 2)
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;      // sign[n & 3] = {1, -1, -1, 1}
    const double xs = __dmul_rn(x, sign);
    if (n & 1) {
        *cos_out = __double2float_rn(glibc_sin_poly(xs, x2));             // cosf: odd quadrant -> sine polynomial
        *sin_out = __double2float_rn(glibc_cos_poly(x2, tsgn));           // sinf: odd quadrant -> cosine polynomial
    } else {
        *cos_out = __double2float_rn(glibc_cos_poly(x2, tsgn));
        *sin_out = __double2float_rn(glibc_sin_poly(xs, x2));
    }
}
}  // namespace ygz_b200
