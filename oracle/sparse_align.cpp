// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of ygz::SparseImgAlign
// (/root/reference/src/SparseImageAlign.cc), its Gauss-Newton driver
// (/root/reference/include/NLSSolver_impl.hpp:17-91, 288-298) and the Sophus SE3f pieces it uses
// (/root/reference/Thirdparty/sophus/sophus/se3.hpp:160-175,235-271,407-428; so3.hpp:234-276,
// 300-348,426-456).  Eigen is not available here: the 6x6 H.ldlt().solve() is restated as
// Eigen's pivoted LDLT^T (Eigen/src/Cholesky/LDLT.h, unblocked + pseudo-inverse of D in solve).
// All arithmetic in float, sequential accumulation order feature-major / pixel row-major.
// Quirks kept on purpose: visible_fts_ cleared once per run, ref_patch_cache_
// never cleared, jacobian_cache_ zeroed per level, stop_ sticky, chi2_ carried across levels.
// Oracle choice where the reference reads out of bounds (iterations[level_] for level_ >= 6):
// n_iter for every level.
#include "ygz_oracle.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct SE3 {
    float q[4];   // x, y, z, w
    float t[3];
};

inline void quat_mul(const float* a, const float* b, float* o) {   // Eigen quaternion product a*b
    const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}

inline void quat_rotate(const float* q, const float* v, float* o) {   // Eigen _transformVector
    // uv = 2 * q.vec x v ; o = v + w*uv + q.vec x uv
    float uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
    o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
    o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

inline void se3_apply(const SE3& T, const float* p, float* o) {   // se3.hpp:257-259
    quat_rotate(T.q, p, o);
    o[0] += T.t[0]; o[1] += T.t[1]; o[2] += T.t[2];
}

inline SE3 se3_mul(const SE3& a, const SE3& b) {   // se3.hpp:160-163 fastMultiply + so3 normalize :268-276
    SE3 r;
    float rt[3];
    quat_rotate(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    quat_mul(a.q, b.q, r.q);
    float len = std::sqrt(r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3]);
    for (int i = 0; i < 4; i++) r.q[i] /= len;
    return r;
}

inline SE3 se3_inv(const SE3& a) {   // se3.hpp:169-174
    SE3 r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    {   // SO3Group(Quaternion) ctor normalises (so3.hpp:741-744)
        float len = std::sqrt(r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3]);
        for (int i = 0; i < 4; i++) r.q[i] /= len;
    }
    float nt[3] = {a.t[0] * -1.f, a.t[1] * -1.f, a.t[2] * -1.f};
    quat_rotate(r.q, nt, r.t);
    return r;
}

inline void quat_to_matrix(const float* q, float* R) {   // Eigen toRotationMatrix
    const float tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

SE3 se3_exp(const float* a) {   // se3.hpp:407-428 ; so3.hpp:426-456 ; twist = (upsilon, omega)
    const float* up = a;
    const float* om = a + 3;
    const float eps = 1e-5f;   // SophusConstants<float>::epsilon, sophus.hpp:55-60
    const float theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const float theta = std::sqrt(theta_sq);
    const float half_theta = 0.5f * theta;
    float imag, real;
    if (theta < eps) {
        const float theta_po4 = theta_sq * theta_sq;
        imag = 0.5f - (float)(1.0 / 48.0) * theta_sq + (float)(1.0 / 3840.0) * theta_po4;
        real = 1.f - 0.5f * theta_sq + (float)(1.0 / 384.0) * theta_po4;
    } else {
        const float s = std::sin(half_theta);
        imag = s / theta;
        real = std::cos(half_theta);
    }
    SE3 T;
    T.q[3] = real; T.q[0] = imag * om[0]; T.q[1] = imag * om[1]; T.q[2] = imag * om[2];
    // the SO3Group(Quaternion) ctor normalises (so3.hpp: "normalize()" in ctor)
    float len = std::sqrt(T.q[0] * T.q[0] + T.q[1] * T.q[1] + T.q[2] * T.q[2] + T.q[3] * T.q[3]);
    for (int i = 0; i < 4; i++) T.q[i] /= len;
    float Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    float Om2[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            Om2[r * 3 + c] = Om[r * 3 + 0] * Om[0 * 3 + c] + Om[r * 3 + 1] * Om[1 * 3 + c] + Om[r * 3 + 2] * Om[2 * 3 + c];
    float V[9];
    if (theta < eps) {
        quat_to_matrix(T.q, V);
    } else {
        const float c1 = (1.f - std::cos(theta)) / theta_sq;
        const float c2 = (theta - std::sin(theta)) / (theta_sq * theta);
        for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.f : 0.f) + c1 * Om[i] + c2 * Om2[i];
    }
    for (int r = 0; r < 3; r++) T.t[r] = V[r * 3 + 0] * up[0] + V[r * 3 + 1] * up[1] + V[r * 3 + 2] * up[2];
    return T;
}

// Eigen LDLT<Matrix6f, Lower>::compute + solve (pivoted, pseudo-inverse of D).  A is 6x6 symmetric.
void ldlt_solve6(const float* Ain, const float* b, float* x) {
    const int n = 6;
    float m[36];
    std::memcpy(m, Ain, sizeof(m));          // row-major; only lower triangle used
    int tr[6];
    auto M = [&](int r, int c) -> float& { return m[r * n + c]; };
    bool zero_first = false;
    for (int k = 0; k < n; k++) {
        int big = k;
        float bigv = std::fabs(M(k, k));
        for (int i = k + 1; i < n; i++) if (std::fabs(M(i, i)) > bigv) { bigv = std::fabs(M(i, i)); big = i; }
        tr[k] = big;
        if (k != big) {
            int s = n - big - 1;
            for (int c = 0; c < k; c++) std::swap(M(k, c), M(big, c));
            for (int i = 0; i < s; i++) std::swap(M(big + 1 + i, k), M(big + 1 + i, big));
            std::swap(M(k, k), M(big, big));
            for (int i = k + 1; i < big; i++) std::swap(M(i, k), M(big, i));
        }
        int rs = n - k - 1;
        if (k > 0) {
            float temp[6];
            for (int j = 0; j < k; j++) temp[j] = M(j, j) * M(k, j);
            float s = 0;
            for (int j = 0; j < k; j++) s += M(k, j) * temp[j];
            M(k, k) -= s;
            for (int i = 0; i < rs; i++) {
                float a = 0;
                for (int j = 0; j < k; j++) a += M(k + 1 + i, j) * temp[j];
                M(k + 1 + i, k) -= a;
            }
        }
        float akk = M(k, k);
        bool valid = std::fabs(akk) > 0;
        if (k == 0 && !valid) {
            for (int j = 0; j < n; j++) tr[j] = j;
            zero_first = true;
            break;
        }
        if (rs > 0 && valid) for (int i = 0; i < rs; i++) M(k + 1 + i, k) /= akk;
    }
    float y[6];
    for (int i = 0; i < n; i++) y[i] = b[i];
    if (zero_first) { for (int i = 0; i < n; i++) x[i] = 0; return; }
    for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);                // P b
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) y[i] -= M(i, j) * y[j];   // L^-1
    const float tol = 1.17549435e-38f;                                     // numeric_limits<float>::min()
    for (int i = 0; i < n; i++) { if (std::fabs(M(i, i)) > tol) y[i] /= M(i, i); else y[i] = 0; }
    for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) y[i] -= M(j, i) * y[j];   // L^-T
    for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);           // P^T
    for (int i = 0; i < n; i++) x[i] = y[i];
}

thread_local bool g_wide_sums = false;

struct Aligner {
    const orc_sia_problem* P;
    SE3 ref_T;
    int level_ = 0;
    std::vector<float> ref_patch_cache_;    // N x 16
    std::vector<float> jacobian_cache_;     // 6 x (16N), column-major (col = 6 floats)
    std::vector<uint8_t> visible_fts_;
    bool have_ref_patch_cache_ = false;
    float H_[36], Jres_[6], x_[6];
    float chi2_ = 1e10f;
    size_t n_meas_ = 0;
    bool stop_ = false;
    int n_iter_ = 10;
    float eps_ = 1e-6f;
    std::vector<float>* trace = nullptr;

    // src/SparseImageAlign.cc:57-128
    void precomputeReferencePatches() {
        const int border = 3;
        const uint8_t* ref_img = P->ref_levels[level_];
        const int cols = P->level_w[level_], rows = P->level_h[level_];
        const int stride = cols;
        const float scale = P->inv_scale_factors[level_];
        const float focal_length = P->fx;
        for (int i = 0; i < P->n; i++) {
            if (!P->valid[i]) continue;
            const float u_ref = P->ref_px[2 * i] * scale;
            const float v_ref = P->ref_px[2 * i + 1] * scale;
            const int u_ref_i = (int)floorf(u_ref);
            const int v_ref_i = (int)floorf(v_ref);
            if (u_ref_i - border < 0 || v_ref_i - border < 0 || u_ref_i + border >= cols || v_ref_i + border >= rows)
                continue;
            visible_fts_[i] = 1;
            float xyz[3];
            se3_apply(ref_T, P->world_xyz + 3 * (size_t)i, xyz);
            // include/SparseImageAlign.h:90-111 JacobXYZ2Cam
            float J0[6], J1[6];
            {
                const float x = xyz[0], y = xyz[1];
                const float z_inv = (float)(1. / xyz[2]);
                const float z_inv_2 = z_inv * z_inv;
                J0[0] = -z_inv; J0[1] = 0.0f; J0[2] = x * z_inv_2; J0[3] = y * J0[2];
                J0[4] = (float)-(1.0 + x * J0[2]); J0[5] = y * z_inv;
                J1[0] = 0.0f; J1[1] = -z_inv; J1[2] = y * z_inv_2; J1[3] = (float)(1.0 + y * J1[2]);
                J1[4] = -J0[3]; J1[5] = -x * z_inv;
            }
            const float su = u_ref - u_ref_i, sv = v_ref - v_ref_i;
            const float w_tl = (float)((1.0 - su) * (1.0 - sv));
            const float w_tr = (float)(su * (1.0 - sv));
            const float w_bl = (float)((1.0 - su) * sv);
            const float w_br = su * sv;
            size_t pixel_counter = 0;
            float* cache_ptr = ref_patch_cache_.data() + 16 * (size_t)i;
            for (int y = 0; y < 4; ++y) {
                const uint8_t* p = ref_img + (v_ref_i + y - 2) * stride + (u_ref_i - 2);
                for (int x = 0; x < 4; ++x, ++p, ++cache_ptr, ++pixel_counter) {
                    *cache_ptr = w_tl * p[0] + w_tr * p[1] + w_bl * p[stride] + w_br * p[stride + 1];
                    float dx = 0.5f * ((w_tl * p[1] + w_tr * p[2] + w_bl * p[stride + 1] + w_br * p[stride + 2]) -
                                       (w_tl * p[-1] + w_tr * p[0] + w_bl * p[stride - 1] + w_br * p[stride]));
                    float dy = 0.5f * ((w_tl * p[stride] + w_tr * p[1 + stride] + w_bl * p[stride * 2] + w_br * p[stride * 2 + 1]) -
                                       (w_tl * p[-stride] + w_tr * p[1 - stride] + w_bl * p[0] + w_br * p[1]));
                    float* Jc = jacobian_cache_.data() + 6 * ((size_t)i * 16 + pixel_counter);
                    const float fs = focal_length * scale;
                    for (int k = 0; k < 6; k++) Jc[k] = (dx * J0[k] + dy * J1[k]) * fs;
                }
            }
        }
        have_ref_patch_cache_ = true;
    }

    // src/SparseImageAlign.cc:130-231 (use_weights_ == false, display_ == false)
    float computeResiduals(const SE3& T, bool linearize) {
        const uint8_t* cur_img = P->cur_levels[level_];
        const int cols = P->level_w[level_], rows = P->level_h[level_];
        if (!have_ref_patch_cache_) precomputeReferencePatches();
        const int stride = cols;
        const int border = 3;
        const float scale = P->inv_scale_factors[level_];
        float chi2 = 0.0f;
        // wide_sums (test switch orc_sia_wide_sums): accumulate H, Jres and chi2 in double and round once — the value the
        // reference's float sums approximate, independent of summation order / FMA contraction.  Used to state how far the
        // reference's own float accumulation is from exact arithmetic on ill-conditioned problems (tests/test_sia_gpu.py).
        double Hd[36] = {0}, Jd[6] = {0}, chi2d = 0.0;
        for (int i = 0; i < P->n; i++) {
            if (!visible_fts_[i]) continue;
            float xyz_ref[3], xyz_cur[3];
            se3_apply(ref_T, P->world_xyz + 3 * (size_t)i, xyz_ref);
            se3_apply(T, xyz_ref, xyz_cur);
            // include/Frame.h:154-159 Camera2Pixel
            const float uc = P->fx * xyz_cur[0] / xyz_cur[2] + P->cx;
            const float vc = P->fy * xyz_cur[1] / xyz_cur[2] + P->cy;
            const float u_cur = uc * scale, v_cur = vc * scale;
            const int u_cur_i = (int)floorf(u_cur), v_cur_i = (int)floorf(v_cur);
            if (u_cur_i < 0 || v_cur_i < 0 || u_cur_i - border < 0 || v_cur_i - border < 0 ||
                u_cur_i + border >= cols || v_cur_i + border >= rows)
                continue;
            const float su = u_cur - u_cur_i, sv = v_cur - v_cur_i;
            const float w_tl = (float)((1.0 - su) * (1.0 - sv));
            const float w_tr = (float)(su * (1.0 - sv));
            const float w_bl = (float)((1.0 - su) * sv);
            const float w_br = su * sv;
            const float* ref_ptr = ref_patch_cache_.data() + 16 * (size_t)i;
            size_t pixel_counter = 0;
            for (int y = 0; y < 4; ++y) {
                const uint8_t* p = cur_img + (v_cur_i + y - 2) * stride + (u_cur_i - 2);
                for (int x = 0; x < 4; ++x, ++pixel_counter, ++p, ++ref_ptr) {
                    const float intensity_cur = w_tl * p[0] + w_tr * p[1] + w_bl * p[stride] + w_br * p[stride + 1];
                    const float res = intensity_cur - (*ref_ptr);
                    chi2 += res * res * 1.0f;
                    chi2d += (double)(res * res);
                    n_meas_++;
                    if (linearize) {
                        const float* J = jacobian_cache_.data() + 6 * ((size_t)i * 16 + pixel_counter);
                        if (g_wide_sums) {
                            for (int r = 0; r < 6; r++)
                                for (int c = 0; c < 6; c++) Hd[r * 6 + c] += (double)(J[r] * J[c]);
                            for (int r = 0; r < 6; r++) Jd[r] -= (double)(J[r] * res);
                        } else {
                            for (int r = 0; r < 6; r++)
                                for (int c = 0; c < 6; c++) H_[r * 6 + c] += J[r] * J[c];
                            for (int r = 0; r < 6; r++) Jres_[r] -= J[r] * res;
                        }
                    }
                }
            }
        }
        if (g_wide_sums) {
            if (linearize) {
                for (int k = 0; k < 36; k++) H_[k] = (float)Hd[k];
                for (int k = 0; k < 6; k++) Jres_[k] = (float)Jd[k];
            }
            chi2 = (float)chi2d;
        }
        return chi2 / n_meas_;
    }

    // include/NLSSolver_impl.hpp:17-91
    void optimizeGaussNewton(SE3& model) {
        SE3 old_model = model;
        for (int iter = 0; iter < n_iter_; ++iter) {
            std::memset(H_, 0, sizeof(H_));
            std::memset(Jres_, 0, sizeof(Jres_));
            n_meas_ = 0;
            float new_chi2 = computeResiduals(model, true);
            ldlt_solve6(H_, Jres_, x_);                       // solve() :233-238
            if (std::isnan(x_[0])) stop_ = true;
            bool rejected = (iter > 0 && new_chi2 > 1.2 * chi2_) || stop_;
            float xmax = -1;
            for (int i = 0; i < 6; i++) xmax = std::fmax(xmax, std::fabs(x_[i]));
            if (trace) {
                float rec[8] = {(float)level_, (float)iter, new_chi2, (float)n_meas_, xmax, rejected ? 0.f : 1.f, 0, 0};
                trace->insert(trace->end(), rec, rec + 8);
            }
            if (rejected) { model = old_model; break; }
            float nx[6];
            for (int i = 0; i < 6; i++) nx[i] = -x_[i];
            SE3 new_model = se3_mul(model, se3_exp(nx));       // update() :240-244
            old_model = model;
            model = new_model;
            chi2_ = new_chi2;
            if (xmax <= eps_) break;
        }
    }
};

}  // namespace

extern "C" {

void orc_se3_exp(const float* x6, float* q, float* t) {
    SE3 T = se3_exp(x6);
    std::memcpy(q, T.q, 16); std::memcpy(t, T.t, 12);
}
void orc_se3_mul(const float* qa, const float* ta, const float* qb, const float* tb, float* q, float* t) {
    SE3 a, b;
    std::memcpy(a.q, qa, 16); std::memcpy(a.t, ta, 12); std::memcpy(b.q, qb, 16); std::memcpy(b.t, tb, 12);
    SE3 r = se3_mul(a, b);
    std::memcpy(q, r.q, 16); std::memcpy(t, r.t, 12);
}
void orc_se3_inv(const float* q, const float* t, float* qo, float* to) {
    SE3 a;
    std::memcpy(a.q, q, 16); std::memcpy(a.t, t, 12);
    SE3 r = se3_inv(a);
    std::memcpy(qo, r.q, 16); std::memcpy(to, r.t, 12);
}

/* the third-party arithmetic, exposed for oracle/ref_mock (the reference's src/SparseImageAlign.cc compiled against the stand-in
 * Eigen / Sophus forwards its ldlt().solve / SE3 operations here: see ref_mock/mock_eigen.hpp) */
void orc_ldlt_solve6(const float* A, const float* b, float* x) { ldlt_solve6(A, b, x); }
void orc_se3_apply(const float* q, const float* t, const float* p, float* o) {
    SE3 T; std::memcpy(T.q, q, 16); std::memcpy(T.t, t, 12);
    se3_apply(T, p, o);
}
void orc_quat_to_matrix(const float* q, float* R) { quat_to_matrix(q, R); }

/* test switch: 1 = H / Jres / chi2 summed in double (see computeResiduals); per calling thread */
void orc_sia_wide_sums(int on) { g_wide_sums = on != 0; }

// src/SparseImageAlign.cc:20-49
int orc_sia_run(const orc_sia_problem* p, int max_level, int min_level, int n_iter, float eps,
                float* out_q, float* out_t, float* H_out, float* trace, int trace_cap, int* trace_n) {
    if (trace_n) *trace_n = 0;
    SE3 ref, cur;
    std::memcpy(ref.q, p->ref_q, 16); std::memcpy(ref.t, p->ref_t, 12);
    std::memcpy(cur.q, p->cur_q, 16); std::memcpy(cur.t, p->cur_t, 12);
    SE3 T = se3_mul(cur, se3_inv(ref));
    if (p->n <= 0) {                       // :24-27 "no features to track" -> 0, TCR untouched
        return 0;
    }
    Aligner A;
    A.P = p;
    A.ref_T = ref;
    A.ref_patch_cache_.assign((size_t)p->n * 16, 0.f);
    A.jacobian_cache_.assign((size_t)p->n * 16 * 6, 0.f);
    A.visible_fts_.assign(p->n, 0);
    A.n_iter_ = n_iter;
    A.eps_ = eps;
    std::vector<float> tr;
    if (trace) A.trace = &tr;
    for (int level = max_level; level >= min_level; level -= 1) {
        A.level_ = level;
        std::fill(A.jacobian_cache_.begin(), A.jacobian_cache_.end(), 0.f);
        A.have_ref_patch_cache_ = false;
        A.optimizeGaussNewton(T);
    }
    std::memcpy(out_q, T.q, 16);
    std::memcpy(out_t, T.t, 12);
    if (H_out) std::memcpy(H_out, A.H_, sizeof(A.H_));
    if (trace) {
        int n = (int)tr.size() / 8;
        if (n > trace_cap) n = trace_cap;
        std::memcpy(trace, tr.data(), (size_t)n * 8 * sizeof(float));
        if (trace_n) *trace_n = n;
    }
    return (int)(A.n_meas_ / 16);
}

}  // extern "C"
