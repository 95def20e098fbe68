// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of the direct (patch) projection search
// of ORB-YGZ-SLAM, SURVEY.md §8(f) rank 1:
//   ORBmatcher::GetWarpAffineMatrix   /root/reference/src/ORBmatcher.cc:1525-1547
//   ORBmatcher::GetBestSearchLevel    /root/reference/include/ORBmatcher.h:185-197
//   ORBmatcher::WarpAffine            /root/reference/src/ORBmatcher.cc:1549-1573
//   GetBilateralInterpUchar           /root/reference/include/ORBmatcher.h:200-211
//   ORBmatcher::FindDirectProjection  /root/reference/src/ORBmatcher.cc:1575-1602
//   ygz::Align2D                      /root/reference/src/Align.cc:8-103
//   KeyFrame::Pixel2Camera / Camera2Pixel / World2Pixel   /root/reference/include/KeyFrame.h:173-194 (float depth)
//   Sophus SE3f product / inverse / point transform       Thirdparty/sophus/sophus/se3.hpp:160-175,239-259
// Eigen is not available: Matrix2f::inverse() and Matrix3f::inverse() are restated as Eigen's closed forms
// (Eigen/src/LU/InverseImpl.h: compute_inverse_size2_helper, compute_inverse<.,3> with cofactor_3x3).
// Arithmetic types follow the reference line by line: float everywhere except the bilinear warp
// (GetBilateralInterpUchar takes doubles) and the interpolation weights of Align2D ((1.0 - subpix) products are
// double, stored to float).  No fma contraction (-ffp-contract=off), accumulation in source order.
// PARITY UNPINNED: the reference has no test or golden vector for this path and cannot be built here; the oracle is
// pinned only by its own fixture (tests/golden/direct_v1.npz) and by recovering known sub-pixel shifts.
#include "ygz_oracle.h"
#include <cmath>
#include <cstring>

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

inline SE3 se3_mul(const SE3& a, const SE3& b) {   // se3.hpp:239-243 -> operator*= :268-276 (normalises)
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
    float len = std::sqrt(r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3]);
    for (int i = 0; i < 4; i++) r.q[i] /= len;
    float nt[3] = {a.t[0] * -1.f, a.t[1] * -1.f, a.t[2] * -1.f};
    quat_rotate(r.q, nt, r.t);
    return r;
}

struct Cam { float fx, fy, cx, cy; };

inline void camera2pixel(const Cam& K, const float* p, float* px) {   // KeyFrame.h:173-178
    px[0] = K.fx * p[0] / p[2] + K.cx;
    px[1] = K.fy * p[1] / p[2] + K.cy;
}

inline void pixel2camera(const Cam& K, const float* px, float depth, float* p) {   // KeyFrame.h:180-186
    p[0] = (px[0] - K.cx) * depth / K.fx;
    p[1] = (px[1] - K.cy) * depth / K.fy;
    p[2] = depth;
}

// include/ORBmatcher.h:200-211
inline uint8_t bilinear_uchar(double x, double y, const uint8_t* data0, int step) {
    const double xx = x - std::floor(x);
    const double yy = y - std::floor(y);
    const uint8_t* data = data0 + (int)y * step + (int)x;
    return (uint8_t)((1 - xx) * (1 - yy) * data[0] + xx * (1 - yy) * data[1] + (1 - xx) * yy * data[step] +
                     xx * yy * data[step + 1]);
}

// Eigen compute_inverse<Matrix3f, 3>: cofactor_3x3<i,j> = m((i+1)%3,(j+1)%3) * m((i+2)%3,(j+2)%3)
//                                                      - m((i+1)%3,(j+2)%3) * m((i+2)%3,(j+1)%3)
inline float cof3(const float m[3][3], int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
}

inline void inverse3(const float m[3][3], float r[3][3]) {
    const float c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
    const float det = c00 * m[0][0] + (c10 * m[0][1] + c20 * m[0][2]);      // (cofactors_col0 .* matrix.col(0)).sum(), Eigen's unrolled redux: e0 + (e1 + e2) — see note
    const float invdet = 1.f / det;
    r[0][0] = c00 * invdet; r[0][1] = c10 * invdet; r[0][2] = c20 * invdet;
    r[1][0] = cof3(m, 0, 1) * invdet; r[1][1] = cof3(m, 1, 1) * invdet; r[1][2] = cof3(m, 2, 1) * invdet;
    r[2][0] = cof3(m, 0, 2) * invdet; r[2][1] = cof3(m, 1, 2) * invdet; r[2][2] = cof3(m, 2, 2) * invdet;
}
// note: Eigen pairs cofactors_col0 with matrix.col(0) = (m00, m10, m20); H is symmetric here (J J^T sums), so
// m[0][k] == m[k][0] and the restatement reads row 0.

// src/Align.cc:8-103
bool align2d(const uint8_t* cur, int cols, int rows, int step, const uint8_t* ref_patch_with_border,
             const uint8_t* ref_patch, int n_iter, float* cur_px) {
    const int halfpatch_size_ = 4, patch_size_ = 8;
    bool converged = false;
    float ref_patch_dx[64], ref_patch_dy[64];
    float H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int ref_step = patch_size_ + 2;
    float* it_dx = ref_patch_dx;
    float* it_dy = ref_patch_dy;
    for (int y = 0; y < patch_size_; ++y) {
        const uint8_t* it = ref_patch_with_border + (y + 1) * ref_step + 1;
        for (int x = 0; x < patch_size_; ++x, ++it, ++it_dx, ++it_dy) {
            float J[3];
            J[0] = (float)(0.5 * (it[1] - it[-1]));
            J[1] = (float)(0.5 * (it[ref_step] - it[-ref_step]));
            J[2] = 1;
            *it_dx = J[0];
            *it_dy = J[1];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) H[a][b] += J[a] * J[b];
        }
    }
    float Hinv[3][3];
    inverse3(H, Hinv);
    float mean_diff = 0;
    float u = cur_px[0], v = cur_px[1];
    const float min_update_squared = (float)(0.03 * 0.03);
    float update[3] = {0, 0, 0};
    for (int iter = 0; iter < n_iter; ++iter) {
        int u_r = (int)std::floor(u);
        int v_r = (int)std::floor(v);
        if (u_r < halfpatch_size_ || v_r < halfpatch_size_ || u_r >= cols - halfpatch_size_ || v_r >= rows - halfpatch_size_)
            break;
        if (std::isnan(u) || std::isnan(v)) return false;
        float subpix_x = u - u_r;
        float subpix_y = v - v_r;
        float wTL = (float)((1.0 - subpix_x) * (1.0 - subpix_y));
        float wTR = (float)(subpix_x * (1.0 - subpix_y));
        float wBL = (float)((1.0 - subpix_x) * subpix_y);
        float wBR = subpix_x * subpix_y;
        const uint8_t* it_ref = ref_patch;
        const float* it_ref_dx = ref_patch_dx;
        const float* it_ref_dy = ref_patch_dy;
        float Jres[3] = {0, 0, 0};
        for (int y = 0; y < patch_size_; ++y) {
            const uint8_t* it = cur + (v_r + y - halfpatch_size_) * step + u_r - halfpatch_size_;
            for (int x = 0; x < patch_size_; ++x, ++it, ++it_ref, ++it_ref_dx, ++it_ref_dy) {
                float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[step] + wBR * it[step + 1];
                float res = search_pixel - *it_ref + mean_diff;
                Jres[0] -= res * (*it_ref_dx);
                Jres[1] -= res * (*it_ref_dy);
                Jres[2] -= res;
            }
        }
        // Matrix3f * Vector3f, coefficient-based product: row . vector summed by the unrolled redux e0 + (e1 + e2)
        for (int a = 0; a < 3; a++) update[a] = Hinv[a][0] * Jres[0] + (Hinv[a][1] * Jres[1] + Hinv[a][2] * Jres[2]);
        u += update[0];
        v += update[1];
        mean_diff += update[2];
        if (update[0] * update[0] + update[1] * update[1] < min_update_squared) {
            converged = true;
            break;
        }
    }
    cur_px[0] = u;
    cur_px[1] = v;
    return converged;
}

}  // namespace

extern "C" int orc_find_direct_projection(const orc_direct_frame* refs, int n_refs, const orc_direct_frame* cur,
                                          const float* K4, const float* scale, const float* inv_scale,
                                          float inv_level_sigma2_1, int nlevels, const orc_direct_item* items,
                                          int n_items, float* out_px, int* out_level, uint8_t* out_ok) {
    const Cam K{K4[0], K4[1], K4[2], K4[3]};
    const int WarpHalfPatchSize = 4, WarpPatchSize = 8;
    for (int n = 0; n < n_items; n++) {
        const orc_direct_item& it = items[n];
        if (it.ref < 0 || it.ref >= n_refs || it.octave < 0 || it.octave >= nlevels) return -1;
        const orc_direct_frame& R = refs[it.ref];
        SE3 pose_ref, pose_cur;
        std::memcpy(pose_ref.q, R.q, 16); std::memcpy(pose_ref.t, R.t, 12);
        std::memcpy(pose_cur.q, cur->q, 16); std::memcpy(pose_cur.t, cur->t, 12);
        const SE3 TCR = se3_mul(pose_cur, se3_inv(pose_ref));                 // :1584
        // ---- GetWarpAffineMatrix (:1525-1547)
        const float px_ref[2] = {it.px_ref[0], it.px_ref[1]};
        float pt_ref[3];
        se3_apply(pose_ref, it.pw, pt_ref);
        const float depth = pt_ref[2];
        const float sl = scale[it.octave];
        const float pdu[2] = {px_ref[0] + (float)WarpHalfPatchSize * sl, px_ref[1] + 0.f * sl};
        const float pdv[2] = {px_ref[0] + 0.f * sl, px_ref[1] + (float)WarpHalfPatchSize * sl};
        float pt_du_ref[3], pt_dv_ref[3];
        pixel2camera(K, pdu, depth, pt_du_ref);
        pixel2camera(K, pdv, depth, pt_dv_ref);
        float pc[3], px_cur[2], px_du[2], px_dv[2];
        se3_apply(TCR, pt_ref, pc); camera2pixel(K, pc, px_cur);
        se3_apply(TCR, pt_du_ref, pc); camera2pixel(K, pc, px_du);
        se3_apply(TCR, pt_dv_ref, pc); camera2pixel(K, pc, px_dv);
        float ACR[2][2];     // columns: (px_du - px_cur)/4, (px_dv - px_cur)/4
        ACR[0][0] = (px_du[0] - px_cur[0]) / (float)WarpHalfPatchSize;
        ACR[1][0] = (px_du[1] - px_cur[1]) / (float)WarpHalfPatchSize;
        ACR[0][1] = (px_dv[0] - px_cur[0]) / (float)WarpHalfPatchSize;
        ACR[1][1] = (px_dv[1] - px_cur[1]) / (float)WarpHalfPatchSize;
        // ---- GetBestSearchLevel (ORBmatcher.h:185-197)
        int search_level = 0;
        float D = ACR[0][0] * ACR[1][1] - ACR[1][0] * ACR[0][1];
        while (D > 3.0 && search_level < nlevels - 1) {
            search_level += 1;
            D *= inv_level_sigma2_1;
        }
        // ---- WarpAffine (:1549-1573), half_patch_size = 5
        uint8_t patch_with_border[100], patch[64];
        {
            const int half_patch_size = WarpHalfPatchSize + 1, patch_size = 2 * half_patch_size;
            const float det = ACR[0][0] * ACR[1][1] - ACR[1][0] * ACR[0][1];
            const float invdet = 1.f / det;                                   // compute_inverse_size2_helper
            const float A00 = ACR[1][1] * invdet, A10 = -ACR[1][0] * invdet, A01 = -ACR[0][1] * invdet, A11 = ACR[0][0] * invdet;
            const float prx = px_ref[0] / scale[it.octave], pry = px_ref[1] / scale[it.octave];
            const uint8_t* img = R.level[it.octave];
            const int cols = R.w[it.octave], rows = R.h[it.octave], step = R.pitch[it.octave];
            uint8_t* pp = patch_with_border;
            for (int y = 0; y < patch_size; y++)
                for (int x = 0; x < patch_size; x++, ++pp) {
                    float qx = (float)(x - half_patch_size), qy = (float)(y - half_patch_size);
                    qx *= scale[search_level]; qy *= scale[search_level];
                    const float px0 = (A00 * qx + A01 * qy) + prx;
                    const float px1 = (A10 * qx + A11 * qy) + pry;
                    if (px0 < 0 || px1 < 0 || px0 >= cols - 1 || px1 >= rows - 1) *pp = 0;
                    else *pp = bilinear_uchar(px0, px1, img, step);
                }
        }
        for (int y = 1; y < WarpPatchSize + 1; ++y)
            for (int x = 0; x < WarpPatchSize; ++x) patch[(y - 1) * WarpPatchSize + x] = patch_with_border[y * (WarpPatchSize + 2) + 1 + x];
        float px_scaled[2] = {it.px_cur[0] * inv_scale[search_level], it.px_cur[1] * inv_scale[search_level]};
        const bool ok = align2d(cur->level[search_level], cur->w[search_level], cur->h[search_level], cur->pitch[search_level],
                                patch_with_border, patch, 10, px_scaled);
        out_px[2 * n] = px_scaled[0] * scale[search_level];
        out_px[2 * n + 1] = px_scaled[1] * scale[search_level];
        out_level[n] = search_level;
        out_ok[n] = ok ? 1 : 0;
    }
    return 0;
}
