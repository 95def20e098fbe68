// SparseImgAlign::run on the device: one persistent CTA per alignment problem runs every pyramid level and
// every Gauss-Newton iteration inside ONE launch (reference: /root/reference/src/SparseImageAlign.cc,
// include/NLSSolver_impl.hpp:17-91, Sophus SE3f from Thirdparty/sophus).
//
//   per level:  zero the Jacobian cache (:41)  ->  precomputeReferencePatches (:57-128), thread per (feature, pixel)
//   per iter :  computeResiduals (:130-231): every thread accumulates H (21 upper entries), Jres (6), chi2, n_meas
//               for its (feature, pixel) items, warp-shuffle + shared-memory reduction in double, then thread 0
//               does solve (6x6 pivoted LDLT, :233-238), the chi2 rollback test (impl:55-65), update
//               T <- T * exp(-x) (:240-244) and the eps stop (impl:87-89).
// Quirks reproduced: visible_fts_ survives across levels, ref_patch_cache_ is never cleared,
// stop_ is sticky, chi2_ is carried across levels, 10 (= n_iter) iterations for every level.
// Tolerance contract: pose within 1e-4 of the CPU path (float accumulation order differs: the reduction tree
// sums in double, the reference sums sequentially in float).
#include "common.cuh"
#include <algorithm>
#include <vector>

using namespace ygz_b200;

namespace {

constexpr int SIA_THREADS = 256;
constexpr int SIA_MAX_LEVELS = 16;
constexpr int NACC = 29;            // 21 (H upper) + 6 (Jres) + chi2 + n_meas
constexpr int TRACE_CAP = 192;      // (levels * iterations) records of 8 floats per problem

struct SiaProblemDev {
    int n_levels;
    int n;
    int level_w[SIA_MAX_LEVELS], level_h[SIA_MAX_LEVELS];
    const uint8_t* img_ref[SIA_MAX_LEVELS];                      // level images on the device: the handle's own arena (ygz_sia_run) or, zero
    const uint8_t* img_cur[SIA_MAX_LEVELS];                      // copy, the pyramids two extractor handles hold (ygz_sia_run_orb)
    int stride_ref[SIA_MAX_LEVELS], stride_cur[SIA_MAX_LEVELS];  // bytes per image row (== level_w in the arena)
    float inv_scale[SIA_MAX_LEVELS];
    float fx, fy, cx, cy;
    float ref_q[4], ref_t[3], cur_q[4], cur_t[3];
};

struct SE3 { float q[4]; float t[3]; };    // unit quaternion (x, y, z, w) + translation, as Sophus::SE3f

__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* o) {
    const float ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__device__ __forceinline__ void quat_rotate(const float* q, const float* v, float* o) {
    float uvx = q[1] * v[2] - q[2] * v[1], uvy = q[2] * v[0] - q[0] * v[2], uvz = q[0] * v[1] - q[1] * v[0];
    uvx += uvx; uvy += uvy; uvz += uvz;
    o[0] = v[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
    o[1] = v[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
    o[2] = v[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
__device__ __forceinline__ void se3_apply(const SE3& T, const float* p, float* o) {
    quat_rotate(T.q, p, o);
    o[0] += T.t[0]; o[1] += T.t[1]; o[2] += T.t[2];
}
__device__ inline void quat_normalize(float* q) {
    const float len = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= len; q[1] /= len; q[2] /= len; q[3] /= len;
}
__device__ inline SE3 se3_mul(const SE3& a, const SE3& b) {        // se3.hpp:160-163,268-271
    SE3 r;
    float rt[3];
    quat_rotate(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    quat_mul(a.q, b.q, r.q);
    quat_normalize(r.q);
    return r;
}
__device__ inline SE3 se3_inv(const SE3& a) {                       // se3.hpp:169-174
    SE3 r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    quat_normalize(r.q);
    const float nt[3] = {a.t[0] * -1.f, a.t[1] * -1.f, a.t[2] * -1.f};
    quat_rotate(r.q, nt, r.t);
    return r;
}
__device__ inline SE3 se3_exp(const float* a) {                     // se3.hpp:407-428, so3.hpp:426-456
    const float* up = a;
    const float* om = a + 3;
    const float eps = 1e-5f;
    const float theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const float theta = sqrtf(theta_sq);
    const float half_theta = 0.5f * theta;
    float imag, real;
    if (theta < eps) {
        const float theta_po4 = theta_sq * theta_sq;
        imag = 0.5f - (float)(1.0 / 48.0) * theta_sq + (float)(1.0 / 3840.0) * theta_po4;
        real = 1.f - 0.5f * theta_sq + (float)(1.0 / 384.0) * theta_po4;
    } else {
        imag = sinf(half_theta) / theta;
        real = cosf(half_theta);
    }
    SE3 T;
    T.q[3] = real; T.q[0] = imag * om[0]; T.q[1] = imag * om[1]; T.q[2] = imag * om[2];
    quat_normalize(T.q);
    const float Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    float V[9];
    if (theta < eps) {
        const float* q = T.q;
        const float tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
        const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
        const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
        const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        V[0] = 1 - (tyy + tzz); V[1] = txy - twz; V[2] = txz + twy;
        V[3] = txy + twz; V[4] = 1 - (txx + tzz); V[5] = tyz - twx;
        V[6] = txz - twy; V[7] = tyz + twx; V[8] = 1 - (txx + tyy);
    } else {
        float Om2[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                Om2[r * 3 + c] = Om[r * 3 + 0] * Om[0 * 3 + c] + Om[r * 3 + 1] * Om[1 * 3 + c] + Om[r * 3 + 2] * Om[2 * 3 + c];
        const float c1 = (1.f - cosf(theta)) / theta_sq;
        const float c2 = (theta - sinf(theta)) / (theta_sq * theta);
        for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.f : 0.f) + c1 * Om[i] + c2 * Om2[i];
    }
    for (int r = 0; r < 3; r++) T.t[r] = V[r * 3 + 0] * up[0] + V[r * 3 + 1] * up[1] + V[r * 3 + 2] * up[2];
    return T;
}

// Eigen LDLT<Matrix6f>::compute + solve: symmetric pivoting on the largest diagonal entry, pseudo-inverse of D.
// Every loop is unrolled and the data-dependent transposition (k <-> big) is dispatched over the possible values of `big`, so that all
// indices are compile-time and the 6 x 6 matrix stays in registers (with run-time indices it lived in local memory: the one thread that
// solves was a chain of ~300 dependent L1 accesses per iteration, and the other 255 waited at the barrier — 42 % of the stall samples
// of the r2 profile).  Same operations in the same order as before.
__device__ __forceinline__ void swapf(float& a, float& b) { const float t = a; a = b; b = t; }
__device__ __forceinline__ void ldlt_solve6(const float* Ain, const float* b, float* x) {
    constexpr int n = 6;
    float m[36];
#pragma unroll
    for (int i = 0; i < 36; i++) m[i] = Ain[i];
    int tr[6] = {0, 1, 2, 3, 4, 5};
    bool zero_first = false;
#define M(r, c) m[(r) * 6 + (c)]
#pragma unroll
    for (int k = 0; k < n; k++) {
        if (zero_first) continue;                               // (the `break` of the rolled loop)
        int big = k;
        float bigv = fabsf(M(k, k));
#pragma unroll
        for (int i = k + 1; i < n; i++) { const float v = fabsf(M(i, i)); if (v > bigv) { bigv = v; big = i; } }
        tr[k] = big;
#pragma unroll
        for (int B = k + 1; B < n; B++) {
            if (big == B) {
#pragma unroll
                for (int c = 0; c < k; c++) swapf(M(k, c), M(B, c));
#pragma unroll
                for (int i = B + 1; i < n; i++) swapf(M(i, k), M(i, B));
                swapf(M(k, k), M(B, B));
#pragma unroll
                for (int i = k + 1; i < B; i++) swapf(M(i, k), M(B, i));
            }
        }
        if (k > 0) {
            float temp[6];
#pragma unroll
            for (int j = 0; j < k; j++) temp[j] = M(j, j) * M(k, j);
            float s = 0;
#pragma unroll
            for (int j = 0; j < k; j++) s += M(k, j) * temp[j];
            M(k, k) -= s;
#pragma unroll
            for (int i = k + 1; i < n; i++) {
                float a = 0;
#pragma unroll
                for (int j = 0; j < k; j++) a += M(i, j) * temp[j];
                M(i, k) -= a;
            }
        }
        const float akk = M(k, k);
        const bool valid = fabsf(akk) > 0;
        if (k == 0 && !valid) { zero_first = true; continue; }
        if (valid) {
#pragma unroll
            for (int i = k + 1; i < n; i++) M(i, k) /= akk;
        }
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < n; i++) y[i] = b[i];
    if (zero_first) {
#pragma unroll
        for (int i = 0; i < n; i++) x[i] = 0;
        return;
    }
#pragma unroll
    for (int k = 0; k < n; k++) {
#pragma unroll
        for (int B = k + 1; B < n; B++) if (tr[k] == B) swapf(y[k], y[B]);
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) y[i] -= M(i, j) * y[j];
    }
#pragma unroll
    for (int i = 0; i < n; i++) { if (fabsf(M(i, i)) > 1.17549435e-38f) y[i] /= M(i, i); else y[i] = 0; }
#pragma unroll
    for (int i = n - 1; i >= 0; i--) {
#pragma unroll
        for (int j = i + 1; j < n; j++) y[i] -= M(j, i) * y[j];
    }
#pragma unroll
    for (int k = n - 1; k >= 0; k--) {
#pragma unroll
        for (int B = k + 1; B < n; B++) if (tr[k] == B) swapf(y[k], y[B]);
    }
#pragma unroll
    for (int i = 0; i < n; i++) x[i] = y[i];
#undef M
}

struct SiaArgs {
    const SiaProblemDev* prob;      // [P]
    const float* ref_px;            // [P][maxN][2]
    const float* xyz;               // [P][maxN][3]
    const uint8_t* valid;           // [P][maxN]
    float* ref_patch;               // [P][maxN][16]
    float* jac;                     // [P][maxN][16][6]
    uint8_t* visible;               // [P][maxN]
    int maxN;
    int max_level, min_level, n_iter;
    float eps;
    float* out_q; float* out_t; float* out_H; int* out_ntracked;
    float* trace; int* trace_n;     // [P][TRACE_CAP][8]
};

// One CTA per problem, one THREAD per feature (patch): the projection of a feature is computed once for its 16 pixels, the 5x5 pixels
// its bilinear taps touch, the 16 cached reference values and the 96 cached Jacobian entries are independent loads issued back to back
// (one memory round trip per Gauss-Newton iteration instead of one per pixel item), and the 29 sums are accumulated over the patch in
// registers.  Per-pixel arithmetic is the reference's (same expressions as computeResiduals :130-231); partial sums are float per
// thread, double across threads.  Solver state lives in shared memory; the cross-warp totals are formed by 29 threads, thread 0 solves
// (register-resident LDLT above).  Measured on B200, C3 (592 problems of 200 patches per launch; one problem alone):
//   pixel item per thread, LDLT in local memory (round 1)   533 k alignments/s   0.407 ms
//   thread per feature                                       613 k                0.336 ms
//   + register-resident LDLT                                 803 k                0.239 ms
//   a WARP per problem instead of a CTA (8 problems per CTA, no block barrier): 565 k — the 77 KB Jacobian cache of a problem then
//   streams from L2 every iteration (16 problems per SM do not fit L1), which costs more than the overlapped solves gain; not kept.
__global__ void __launch_bounds__(SIA_THREADS, 2) k_sia_align(SiaArgs A) {
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SiaProblemDev& P = A.prob[p];
    const int N = P.n;
    const float* ref_px = A.ref_px + (size_t)p * A.maxN * 2;
    const float* xyz_w = A.xyz + (size_t)p * A.maxN * 3;
    const uint8_t* valid = A.valid + (size_t)p * A.maxN;
    float* ref_patch = A.ref_patch + (size_t)p * A.maxN * 16;
    float* jac = A.jac + (size_t)p * A.maxN * 96;
    uint8_t* visible = A.visible + (size_t)p * A.maxN;
    const int MN = A.maxN;

    __shared__ double s_red[SIA_THREADS / 32][NACC];
    __shared__ double s_tot[NACC];
    __shared__ SE3 s_T, s_old;      // current model, model before the last accepted update (old_model_)
    __shared__ float s_H[36];       // H_ of the last evaluated iteration
    __shared__ int s_ctl;           // 0 continue, 1 leave the level loop
    __shared__ int s_nmeas;

    SE3 refT;
    for (int i = 0; i < 4; i++) refT.q[i] = P.ref_q[i];
    for (int i = 0; i < 3; i++) refT.t[i] = P.ref_t[i];

    // thread-0 solver state (NLLSSolver members)
    float chi2_ = 1e10f;
    bool stop_ = false;
    int trace_count = 0;

    for (int i = tid; i < N; i += SIA_THREADS) visible[i] = 0;         // visible_fts_ = vector<bool>(N, false), :35
    if (tid < 36) s_H[tid] = 0.f;
    if (tid == 0) {
        SE3 cur;
        for (int i = 0; i < 4; i++) cur.q[i] = P.cur_q[i];
        for (int i = 0; i < 3; i++) cur.t[i] = P.cur_t[i];
        s_T = se3_mul(cur, se3_inv(refT));                             // T_cur_from_ref, :37
        s_nmeas = 0;
    }
    __syncthreads();

    for (int level = A.max_level; level >= A.min_level; level--) {
        const int cols = P.level_w[level], rows = P.level_h[level];
        const uint8_t* ref_img = P.img_ref[level];
        const uint8_t* cur_img = P.img_cur[level];
        const float scale = P.inv_scale[level];
        const int stride = P.stride_ref[level], cstride = P.stride_cur[level];
        // ---- jacobian_cache_.setZero() (:41) + precomputeReferencePatches (:57-128); each thread owns the rows of its features
        for (int f = tid; f < N; f += SIA_THREADS) {
            float4* Jf = reinterpret_cast<float4*>(jac) + f;                    // plane layout [24][maxN] float4: lanes read / write neighbours
            bool ok = valid[f] != 0;
            float u_ref = 0.f, v_ref = 0.f;
            int u_ref_i = 0, v_ref_i = 0;
            if (ok) {
                u_ref = ref_px[2 * f] * scale; v_ref = ref_px[2 * f + 1] * scale;
                u_ref_i = (int)floorf(u_ref); v_ref_i = (int)floorf(v_ref);
                ok = !(u_ref_i - 3 < 0 || v_ref_i - 3 < 0 || u_ref_i + 3 >= cols || v_ref_i + 3 >= rows);
            }
            if (!ok) {                                                  // the row keeps its zeros; ref_patch_cache_ keeps the coarser level's values
#pragma unroll
                for (int k = 0; k < 24; k++) Jf[(size_t)k * MN] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            visible[f] = 1;
            float xyz[3];
            se3_apply(refT, xyz_w + 3 * f, xyz);
            float J0[6], J1[6];                                         // JacobXYZ2Cam, include/SparseImageAlign.h:90-111
            {
                const float x = xyz[0], y = xyz[1];
                const float z_inv = (float)(1. / xyz[2]);
                const float z_inv_2 = z_inv * z_inv;
                J0[0] = -z_inv; J0[1] = 0.0f; J0[2] = x * z_inv_2; J0[3] = y * J0[2];
                J0[4] = -(1.0f + x * J0[2]); J0[5] = y * z_inv;
                J1[0] = 0.0f; J1[1] = -z_inv; J1[2] = y * z_inv_2; J1[3] = 1.0f + y * J1[2];
                J1[4] = -J0[3]; J1[5] = -x * z_inv;
            }
            const float su = u_ref - u_ref_i, sv = v_ref - v_ref_i;
            const float w_tl = (1.0f - su) * (1.0f - sv), w_tr = su * (1.0f - sv), w_bl = (1.0f - su) * sv, w_br = su * sv;
            const float fs = P.fx * scale;
            // the 7 x 7 pixels the values and the central differences of the 4 x 4 patch touch
            float R[7][7];
            const uint8_t* r0 = ref_img + (v_ref_i - 3) * stride + (u_ref_i - 3);
#pragma unroll
            for (int r = 0; r < 7; r++)
#pragma unroll
                for (int c = 0; c < 7; c++) R[r][c] = (float)r0[r * stride + c];
            float4* rpw = reinterpret_cast<float4*>(ref_patch) + f;                // [4][maxN] float4
            float rp[16];
#pragma unroll
            for (int pp = 0; pp < 8; pp++) {                            // two pixels = 12 Jacobian entries = 3 float4
                float Jr[12];
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const int pc = 2 * pp + h2, py = pc >> 2, px = pc & 3;
                    // q = pixel (v_ref_i + py - 2, u_ref_i - 2 + px) = R[py + 1][px + 1]
#define Q(dy, dx) R[py + 1 + (dy)][px + 1 + (dx)]
                    rp[pc] = w_tl * Q(0, 0) + w_tr * Q(0, 1) + w_bl * Q(1, 0) + w_br * Q(1, 1);
                    const float dx = 0.5f * ((w_tl * Q(0, 1) + w_tr * Q(0, 2) + w_bl * Q(1, 1) + w_br * Q(1, 2)) -
                                             (w_tl * Q(0, -1) + w_tr * Q(0, 0) + w_bl * Q(1, -1) + w_br * Q(1, 0)));
                    const float dy = 0.5f * ((w_tl * Q(1, 0) + w_tr * Q(1, 1) + w_bl * Q(2, 0) + w_br * Q(2, 1)) -
                                             (w_tl * Q(-1, 0) + w_tr * Q(-1, 1) + w_bl * Q(0, 0) + w_br * Q(0, 1)));
#undef Q
#pragma unroll
                    for (int k = 0; k < 6; k++) Jr[h2 * 6 + k] = (dx * J0[k] + dy * J1[k]) * fs;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) Jf[(size_t)(pp * 3 + k) * MN] = make_float4(Jr[4 * k], Jr[4 * k + 1], Jr[4 * k + 2], Jr[4 * k + 3]);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) rpw[(size_t)k * MN] = make_float4(rp[4 * k], rp[4 * k + 1], rp[4 * k + 2], rp[4 * k + 3]);
        }
        __syncthreads();
        if (tid == 0) s_old = s_T;

        // ---- optimizeGaussNewton (include/NLSSolver_impl.hpp:17-91)
        for (int iter = 0; iter < A.n_iter; ++iter) {
            const SE3 T = s_T;
            float acc[NACC];
#pragma unroll
            for (int k = 0; k < NACC; k++) acc[k] = 0.f;
            for (int f = tid; f < N; f += SIA_THREADS) {
                if (!visible[f]) continue;
                float xyz_ref[3], xyz_cur[3];
                se3_apply(refT, xyz_w + 3 * f, xyz_ref);
                se3_apply(T, xyz_ref, xyz_cur);
                const float uc = P.fx * xyz_cur[0] / xyz_cur[2] + P.cx;      // Camera2Pixel, include/Frame.h:154-159
                const float vc = P.fy * xyz_cur[1] / xyz_cur[2] + P.cy;
                const float u_cur = uc * scale, v_cur = vc * scale;
                const int u_cur_i = (int)floorf(u_cur), v_cur_i = (int)floorf(v_cur);
                if (u_cur_i < 0 || v_cur_i < 0 || u_cur_i - 3 < 0 || v_cur_i - 3 < 0 || u_cur_i + 3 >= cols || v_cur_i + 3 >= rows)
                    continue;
                const float su = u_cur - u_cur_i, sv = v_cur - v_cur_i;
                const float w_tl = (1.0f - su) * (1.0f - sv), w_tr = su * (1.0f - sv), w_bl = (1.0f - su) * sv, w_br = su * sv;
                float C[5][5];                                           // rows v-2 .. v+2, columns u-2 .. u+2
                const uint8_t* c0 = cur_img + (v_cur_i - 2) * cstride + (u_cur_i - 2);
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    // the five bytes of a row sit in two aligned words (two L1 requests per lane instead of five; the bytes outside
                    // the five belong to the same image: the patch is at least 3 pixels and one row away from its borders)
                    const uintptr_t a = reinterpret_cast<uintptr_t>(c0 + r * cstride);
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                    const uint32_t w0 = w[0], w1 = w[1];
                    const unsigned sh = ((unsigned)a & 3u) * 8u;
                    const uint32_t lo = __funnelshift_r(w0, w1, sh);
                    C[r][0] = (float)(lo & 0xFFu); C[r][1] = (float)((lo >> 8) & 0xFFu); C[r][2] = (float)((lo >> 16) & 0xFFu);
                    C[r][3] = (float)(lo >> 24); C[r][4] = (float)((w1 >> sh) & 0xFFu);
                }
                const float4* rp4 = reinterpret_cast<const float4*>(ref_patch) + f;
                const float4* J4 = reinterpret_cast<const float4*>(jac) + f;
#pragma unroll
                for (int qd = 0; qd < 4; qd++) {                        // one patch row = 4 pixels = 6 float4 of Jacobian rows at a time
                    const float4 rv = rp4[(size_t)qd * MN];
                    const float rpv[4] = {rv.x, rv.y, rv.z, rv.w};
                    float Jv[24];
#pragma unroll
                    for (int k = 0; k < 6; k++) { const float4 v = J4[(size_t)(qd * 6 + k) * MN]; Jv[4 * k] = v.x; Jv[4 * k + 1] = v.y; Jv[4 * k + 2] = v.z; Jv[4 * k + 3] = v.w; }
#pragma unroll
                    for (int px = 0; px < 4; px++) {
                        const int py = qd;
                        const float intensity = w_tl * C[py][px] + w_tr * C[py][px + 1] + w_bl * C[py + 1][px] + w_br * C[py + 1][px + 1];
                        const float res = intensity - rpv[px];
                        acc[27] += res * res;
                        acc[28] += 1.f;
                        const float j0 = Jv[px * 6], j1 = Jv[px * 6 + 1], j2 = Jv[px * 6 + 2], j3 = Jv[px * 6 + 3], j4 = Jv[px * 6 + 4], j5 = Jv[px * 6 + 5];
                        acc[0] += j0 * j0; acc[1] += j0 * j1; acc[2] += j0 * j2; acc[3] += j0 * j3; acc[4] += j0 * j4; acc[5] += j0 * j5;
                        acc[6] += j1 * j1; acc[7] += j1 * j2; acc[8] += j1 * j3; acc[9] += j1 * j4; acc[10] += j1 * j5;
                        acc[11] += j2 * j2; acc[12] += j2 * j3; acc[13] += j2 * j4; acc[14] += j2 * j5;
                        acc[15] += j3 * j3; acc[16] += j3 * j4; acc[17] += j3 * j5;
                        acc[18] += j4 * j4; acc[19] += j4 * j5; acc[20] += j5 * j5;
                        acc[21] -= j0 * res; acc[22] -= j1 * res; acc[23] -= j2 * res; acc[24] -= j3 * res; acc[25] -= j4 * res; acc[26] -= j5 * res;
                    }
                }
            }
            // block reduction in double.  Within a warp the 29 sums are reduced TOGETHER: at xor-distance o every lane hands half of the
            // values it still holds to its partner and adds the partner's other half, so the number of live values halves each step and
            // lane k ends with the warp total of sum k — 31 exchanges instead of 29 x 5 (the shuffles were 29 % of the kernel's
            // instructions and its largest MIO load).  Same additions between the same lanes as a per-value butterfly: identical bits.
            {
                double d[32];
#pragma unroll
                for (int k = 0; k < 32; k++) d[k] = k < NACC ? (double)acc[k] : 0.0;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const bool up = (lane & o) != 0;
#pragma unroll
                    for (int j = 0; j < o; j++) {
                        const double keep = up ? d[j + o] : d[j];
                        const double send = up ? d[j] : d[j + o];
                        d[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                if (lane < NACC) s_red[warp][lane] = d[0];
            }
            __syncthreads();
            if (warp == 0) {
                if (lane < NACC) {
                    double s = 0;
                    for (int w = 0; w < SIA_THREADS / 32; w++) s += s_red[w][lane];
                    s_tot[lane] = s;
                }
                __syncwarp();
            }
            if (tid == 0) {
                const double* tot = s_tot;
                float H[36], Jres[6], x[6];
                {
                    int k = 0;
                    for (int r = 0; r < 6; r++)
                        for (int c = r; c < 6; c++) { H[r * 6 + c] = (float)tot[k]; H[c * 6 + r] = (float)tot[k]; k++; }
                    for (int r = 0; r < 6; r++) Jres[r] = (float)tot[21 + r];
                }
                for (int i = 0; i < 36; i++) s_H[i] = H[i];
                const int n_meas = (int)(tot[28] + 0.5);
                s_nmeas = n_meas;
                const float new_chi2 = (float)tot[27] / (float)n_meas;         // chi2 / n_meas_ (NaN when nothing is visible)
                ldlt_solve6(H, Jres, x);                                        // solve() :233-238
                if (isnan(x[0])) stop_ = true;
                const bool rejected = (iter > 0 && new_chi2 > 1.2f * chi2_) || stop_;
                float xmax = -1.f;
                for (int i = 0; i < 6; i++) xmax = fmaxf(xmax, fabsf(x[i]));
                if (A.trace && trace_count < TRACE_CAP) {
                    float* tr = A.trace + ((size_t)p * TRACE_CAP + trace_count) * 8;
                    tr[0] = (float)level; tr[1] = (float)iter; tr[2] = new_chi2; tr[3] = (float)n_meas; tr[4] = xmax;
                    tr[5] = rejected ? 0.f : 1.f; tr[6] = 0.f; tr[7] = 0.f;
                    trace_count++;
                }
                int ctl = 0;
                if (rejected) {
                    s_T = s_old;                                                // rollback, impl:63-64
                    ctl = 1;
                } else {
                    float nx[6];
                    for (int i = 0; i < 6; i++) nx[i] = -x[i];
                    const SE3 nm = se3_mul(T, se3_exp(nx));                     // update() :240-244
                    s_old = T;
                    s_T = nm;
                    chi2_ = new_chi2;
                    if (xmax <= A.eps) ctl = 1;                                 // impl:87-89
                }
                s_ctl = ctl;
            }
            __syncthreads();
            if (s_ctl) break;
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int i = 0; i < 4; i++) A.out_q[p * 4 + i] = s_T.q[i];
        for (int i = 0; i < 3; i++) A.out_t[p * 3 + i] = s_T.t[i];
        if (A.out_H) for (int i = 0; i < 36; i++) A.out_H[p * 36 + i] = s_H[i];
        A.out_ntracked[p] = s_nmeas / 16;                                       // n_meas_ / patch_area_, :48
        if (A.trace_n) A.trace_n[p] = trace_count;
    }
}

}  // namespace

struct ygz_sia {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int maxP = 0, maxN = 0;
    size_t arena_per_problem = 0;
    SiaProblemDev* d_prob = nullptr;
    uint8_t* d_arena = nullptr;
    float *d_px = nullptr, *d_xyz = nullptr, *d_patch = nullptr, *d_jac = nullptr;
    uint8_t *d_valid = nullptr, *d_visible = nullptr;
    float *d_q = nullptr, *d_t = nullptr, *d_H = nullptr, *d_trace = nullptr;
    int *d_ntr = nullptr, *d_trace_n = nullptr;
    int last_n = 0;
    uint8_t* h_feat = nullptr;      // pinned staging of the per-problem feature arrays (px | xyz | valid), one H2D copy each
    cudaEvent_t ev_dep = nullptr;   // orders this handle's stream behind the extractor streams in ygz_sia_run_orb
};

static int sia_free(ygz_sia* h) {
    if (!h) return YGZ_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_prob); cudaFree(h->d_arena); cudaFree(h->d_px); cudaFree(h->d_xyz); cudaFree(h->d_patch); cudaFree(h->d_jac);
    cudaFree(h->d_valid); cudaFree(h->d_visible); cudaFree(h->d_q); cudaFree(h->d_t); cudaFree(h->d_H); cudaFree(h->d_trace);
    cudaFree(h->d_ntr); cudaFree(h->d_trace_n);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_dep) cudaEventDestroy(h->ev_dep);
    if (h->h_feat) cudaFreeHost(h->h_feat);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return YGZ_OK;
}

static SiaArgs make_args(ygz_sia* h, int max_level, int min_level, int n_iter, float eps) {
    SiaArgs A;
    A.prob = h->d_prob; A.ref_px = h->d_px; A.xyz = h->d_xyz; A.valid = h->d_valid;
    A.ref_patch = h->d_patch; A.jac = h->d_jac; A.visible = h->d_visible; A.maxN = h->maxN;
    A.max_level = max_level; A.min_level = min_level; A.n_iter = n_iter; A.eps = eps;
    A.out_q = h->d_q; A.out_t = h->d_t; A.out_H = h->d_H; A.out_ntracked = h->d_ntr;
    A.trace = h->d_trace; A.trace_n = h->d_trace_n;
    return A;
}

// Features of all problems through one pinned staging buffer (three H2D copies instead of 3 per problem), problem table, the
// kernel, results back.  Shared tail of ygz_sia_run / ygz_sia_run_orb.
static int sia_finish(ygz_sia* h, const ygz_sia_problem* problems, const std::vector<SiaProblemDev>& pd, int n_problems, int max_level, int min_level,
                      int n_iter, float eps, float* out_q, float* out_t, float* out_H, int* out_n_tracked) {
    const size_t N = h->maxN, P = n_problems;
    float* s_px = (float*)h->h_feat;
    float* s_xyz = (float*)(h->h_feat + (size_t)h->maxP * N * 8);
    uint8_t* s_valid = h->h_feat + (size_t)h->maxP * N * 20;
    for (int p = 0; p < n_problems; p++) {
        const ygz_sia_problem& S = problems[p];
        if (S.n > 0) {
            YGZ_ARG_CHECK(S.ref_px && S.world_xyz && S.valid);
            std::memcpy(s_px + p * N * 2, S.ref_px, (size_t)S.n * 8);
            std::memcpy(s_xyz + p * N * 3, S.world_xyz, (size_t)S.n * 12);
            std::memcpy(s_valid + p * N, S.valid, (size_t)S.n);
        }
    }
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_px, s_px, P * N * 8, cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_xyz, s_xyz, P * N * 12, cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_valid, s_valid, P * N, cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_prob, pd.data(), P * sizeof(SiaProblemDev), cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));      // pd / pageable sources must outlive the copies
    SiaArgs A = make_args(h, max_level, min_level, n_iter, eps);
    { NvtxRange nvtx_range("ygz/SparseImgAlign::run");
    k_sia_align<<<n_problems, SIA_THREADS, 0, h->stream>>>(A);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    h->last_n = n_problems;
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_q, h->d_q, P * 16, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_t, h->d_t, P * 12, cudaMemcpyDeviceToHost, h->stream));
    if (out_H) YGZ_CUDA_TRY(cudaMemcpyAsync(out_H, h->d_H, P * 144, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_n_tracked, h->d_ntr, P * 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    // run() returns 0 and leaves TCR untouched when the reference frame has no keys (:24-27)
    for (int p = 0; p < n_problems; p++)
        if (problems[p].n == 0) out_n_tracked[p] = 0;
    return YGZ_OK;
}

extern "C" {

int ygz_sia_create(int device, int max_problems, int max_features, int max_level_pixels, ygz_sia** out) {
    YGZ_ARG_CHECK(out && max_problems >= 1 && max_features >= 1 && max_level_pixels >= 1);
    if (ygz_device_count() <= device) { set_last_error("no such CUDA device"); return YGZ_ERR_NO_DEVICE; }
    YGZ_CUDA_TRY(cudaSetDevice(device));
    ygz_sia* h = new ygz_sia();
    h->device = device; h->maxP = max_problems; h->maxN = max_features;
    h->arena_per_problem = (size_t)2 * max_level_pixels + 64;
    const size_t P = max_problems, N = max_features;
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); sia_free(h); return YGZ_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY_OR_FREE(cudaEventCreate(&h->ev0));
    TRY_OR_FREE(cudaEventCreate(&h->ev1));
    TRY_OR_FREE(cudaEventCreateWithFlags(&h->ev_dep, cudaEventDisableTiming));
    TRY_OR_FREE(cudaMallocHost(&h->h_feat, P * N * 21));
    TRY_OR_FREE(cudaMalloc(&h->d_prob, P * sizeof(SiaProblemDev)));
    TRY_OR_FREE(cudaMalloc(&h->d_arena, P * h->arena_per_problem));
    TRY_OR_FREE(cudaMalloc(&h->d_px, P * N * 8));
    TRY_OR_FREE(cudaMalloc(&h->d_xyz, P * N * 12));
    TRY_OR_FREE(cudaMalloc(&h->d_valid, P * N));
    TRY_OR_FREE(cudaMalloc(&h->d_visible, P * N));
    TRY_OR_FREE(cudaMalloc(&h->d_patch, P * N * 64));
    TRY_OR_FREE(cudaMemset(h->d_patch, 0, P * N * 64));
    TRY_OR_FREE(cudaMalloc(&h->d_jac, P * N * 384));
    TRY_OR_FREE(cudaMalloc(&h->d_q, P * 16));
    TRY_OR_FREE(cudaMalloc(&h->d_t, P * 12));
    TRY_OR_FREE(cudaMalloc(&h->d_H, P * 144));
    TRY_OR_FREE(cudaMalloc(&h->d_ntr, P * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_trace, P * TRACE_CAP * 32));
    TRY_OR_FREE(cudaMalloc(&h->d_trace_n, P * 4));
#undef TRY_OR_FREE
    *out = h;
    return YGZ_OK;
}

int ygz_sia_destroy(ygz_sia* h) { return sia_free(h); }

int ygz_sia_run(ygz_sia* h, const ygz_sia_problem* problems, int n_problems, int max_level, int min_level, int n_iter, float eps,
                float* out_q, float* out_t, float* out_H, int* out_n_tracked) {
    YGZ_ARG_CHECK(h && problems && n_problems >= 1 && n_problems <= h->maxP && out_q && out_t && out_n_tracked);
    YGZ_ARG_CHECK(min_level >= 0 && max_level >= min_level && max_level < SIA_MAX_LEVELS && n_iter >= 0);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    std::vector<SiaProblemDev> pd(n_problems);
    for (int p = 0; p < n_problems; p++) {
        const ygz_sia_problem& S = problems[p];
        YGZ_ARG_CHECK(S.n_levels > max_level && S.n_levels <= SIA_MAX_LEVELS && S.n >= 0 && S.n <= h->maxN);
        SiaProblemDev& D = pd[p];
        std::memset(&D, 0, sizeof(D));
        D.n_levels = S.n_levels; D.n = S.n;
        D.fx = S.fx; D.fy = S.fy; D.cx = S.cx; D.cy = S.cy;
        std::memcpy(D.ref_q, S.ref_q, 16); std::memcpy(D.ref_t, S.ref_t, 12);
        std::memcpy(D.cur_q, S.cur_q, 16); std::memcpy(D.cur_t, S.cur_t, 12);
        size_t off = (size_t)p * h->arena_per_problem;
        const size_t end = off + h->arena_per_problem;
        for (int l = min_level; l <= max_level; l++) {                 // only the levels run() touches
            const size_t bytes = (size_t)S.level_w[l] * S.level_h[l];
            D.level_w[l] = S.level_w[l]; D.level_h[l] = S.level_h[l]; D.inv_scale[l] = S.inv_scale_factors[l];
            if (off + 2 * bytes > end) { set_last_error("level images exceed max_level_pixels"); return YGZ_ERR_CAPACITY; }
            D.img_ref[l] = h->d_arena + off; D.stride_ref[l] = S.level_w[l];
            YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_arena + off, S.ref_levels[l], bytes, cudaMemcpyHostToDevice, h->stream));
            off += bytes;
            D.img_cur[l] = h->d_arena + off; D.stride_cur[l] = S.level_w[l];
            YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_arena + off, S.cur_levels[l], bytes, cudaMemcpyHostToDevice, h->stream));
            off += bytes;
        }
    }
    return sia_finish(h, problems, pd, n_problems, max_level, min_level, n_iter, eps, out_q, out_t, out_H, out_n_tracked);
}

int ygz_sia_run_orb(ygz_sia* h, ygz_orb* orb_ref, ygz_orb* orb_cur, const ygz_sia_problem* problems, const int* ref_frame, const int* cur_frame,
                    int n_problems, int max_level, int min_level, int n_iter, float eps,
                    float* out_q, float* out_t, float* out_H, int* out_n_tracked) {
    YGZ_ARG_CHECK(h && orb_ref && orb_cur && problems && ref_frame && cur_frame && n_problems >= 1 && n_problems <= h->maxP && out_q && out_t && out_n_tracked);
    YGZ_ARG_CHECK(min_level >= 0 && max_level >= min_level && max_level < SIA_MAX_LEVELS && n_iter >= 0);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const int Bref = ygz_orb_max_batch(orb_ref), Bcur = ygz_orb_max_batch(orb_cur);
    float scale[16], inv_scale[16], sg[16], isg[16];
    int fpl[16];
    int rc = ygz_orb_tables(orb_ref, scale, inv_scale, sg, isg, fpl);      // ref_frame->mvInvScaleFactors (src/SparseImageAlign.cc:86)
    if (rc) return rc;
    const uint8_t *base_ref[SIA_MAX_LEVELS] = {}, *base_cur[SIA_MAX_LEVELS] = {};
    int pit_ref[SIA_MAX_LEVELS] = {}, pit_cur[SIA_MAX_LEVELS] = {}, lw[SIA_MAX_LEVELS] = {}, lh[SIA_MAX_LEVELS] = {};
    size_t fs_ref[SIA_MAX_LEVELS] = {}, fs_cur[SIA_MAX_LEVELS] = {};
    for (int l = min_level; l <= max_level; l++) {
        int w2 = 0, h2 = 0;
        void *pr = nullptr, *pc = nullptr;
        if (ygz_orb_level_size(orb_ref, l, &lw[l], &lh[l]) != YGZ_OK || ygz_orb_level_size(orb_cur, l, &w2, &h2) != YGZ_OK) {
            set_last_error("max_level exceeds the extractor's pyramid"); return YGZ_ERR_ARG;
        }
        if (w2 != lw[l] || h2 != lh[l]) { set_last_error("reference / current extractors differ"); return YGZ_ERR_ARG; }
        ygz_orb_device_level(orb_ref, l, &pr, &pit_ref[l], &fs_ref[l]);
        ygz_orb_device_level(orb_cur, l, &pc, &pit_cur[l], &fs_cur[l]);
        base_ref[l] = (const uint8_t*)pr; base_cur[l] = (const uint8_t*)pc;
    }
    std::vector<SiaProblemDev> pd(n_problems);
    for (int p = 0; p < n_problems; p++) {
        const ygz_sia_problem& S = problems[p];
        YGZ_ARG_CHECK(S.n >= 0 && S.n <= h->maxN && ref_frame[p] >= 0 && ref_frame[p] < Bref && cur_frame[p] >= 0 && cur_frame[p] < Bcur);
        SiaProblemDev& D = pd[p];
        std::memset(&D, 0, sizeof(D));
        D.n_levels = max_level + 1; D.n = S.n;
        D.fx = S.fx; D.fy = S.fy; D.cx = S.cx; D.cy = S.cy;
        std::memcpy(D.ref_q, S.ref_q, 16); std::memcpy(D.ref_t, S.ref_t, 12);
        std::memcpy(D.cur_q, S.cur_q, 16); std::memcpy(D.cur_t, S.cur_t, 12);
        for (int l = min_level; l <= max_level; l++) {
            D.level_w[l] = lw[l]; D.level_h[l] = lh[l]; D.inv_scale[l] = inv_scale[l];
            D.img_ref[l] = base_ref[l] + (size_t)ref_frame[p] * fs_ref[l]; D.stride_ref[l] = pit_ref[l];
            D.img_cur[l] = base_cur[l] + (size_t)cur_frame[p] * fs_cur[l]; D.stride_cur[l] = pit_cur[l];
        }
    }
    // the pyramids are written on the extractors' streams: order this handle's stream behind both
    YGZ_CUDA_TRY(cudaEventRecord(h->ev_dep, (cudaStream_t)ygz_orb_stream(orb_ref)));
    YGZ_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_dep, 0));
    if (orb_cur != orb_ref) {
        YGZ_CUDA_TRY(cudaEventRecord(h->ev_dep, (cudaStream_t)ygz_orb_stream(orb_cur)));
        YGZ_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_dep, 0));
    }
    return sia_finish(h, problems, pd, n_problems, max_level, min_level, n_iter, eps, out_q, out_t, out_H, out_n_tracked);
}

int ygz_sia_last_trace(ygz_sia* h, int problem, float* out, int cap, int* n) {
    YGZ_ARG_CHECK(h && out && n && problem >= 0 && problem < h->last_n);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    int cnt = 0;
    YGZ_CUDA_TRY(cudaMemcpy(&cnt, h->d_trace_n + problem, 4, cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, cap);
    if (cnt) YGZ_CUDA_TRY(cudaMemcpy(out, h->d_trace + (size_t)problem * TRACE_CAP * 8, (size_t)cnt * 32, cudaMemcpyDeviceToHost));
    *n = cnt;
    return YGZ_OK;
}

int ygz_sia_time(ygz_sia* h, int n_problems, int max_level, int min_level, int n_iter, float eps, int iters, float* ms) {
    YGZ_ARG_CHECK(h && ms && iters >= 1 && n_problems >= 1 && n_problems <= h->last_n);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    SiaArgs A = make_args(h, max_level, min_level, n_iter, eps);
    YGZ_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    { NvtxRange nvtx_range("ygz/SparseImgAlign::run");
    for (int i = 0; i < iters; i++) k_sia_align<<<n_problems, SIA_THREADS, 0, h->stream>>>(A);
    }
    YGZ_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    YGZ_CUDA_TRY(cudaEventSynchronize(h->ev1));
    YGZ_CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return YGZ_OK;
}

}  // extern "C"
