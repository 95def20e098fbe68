// cosf / sinf as the reference's host libm evaluates them (computeOrbDescriptor, src/ORBextractor.cc:109:
// `(float)cos(angle)` with a float argument -> glibc cosf/sinf).
// v1: double-precision device cos/sin rounded to float (correctly rounded float result except when the
// double result lies within ~2 ulp_double of a float rounding boundary).
#pragma once
namespace ygz_b200 {
__device__ inline void ref_cos_sin(float angle_rad, float* c, float* s) {
    double sd, cd;
    sincos((double)angle_rad, &sd, &cd);
    *c = (float)cd;
    *s = (float)sd;
}
}  // namespace ygz_b200
