// Direct (patch) projection search of the YGZ front end — SURVEY.md §8(f) rank 1:
//   ORBmatcher::FindDirectProjection  /root/reference/src/ORBmatcher.cc:1575-1602
//   ORBmatcher::GetWarpAffineMatrix   :1525-1547      ORBmatcher::WarpAffine :1549-1573
//   GetBestSearchLevel / GetBilateralInterpUchar      /root/reference/include/ORBmatcher.h:185-211
//   ygz::Align2D                      /root/reference/src/Align.cc:8-103
// called once per (map point, observing keyframe) by Tracking::SearchLocalPointsDirect (src/Tracking.cc:2174-2326).
//
// One THREAD per item: Align2D accumulates its 3-vector Jres over the 64 patch pixels sequentially in float, and
// the convergence test (|update|^2 < 0.03^2) decides the return value, so the accumulation order is part of the
// result.  A thread-per-item mapping keeps the reference's order bit for bit; the parallelism is the hundreds of
// map points per frame times the frames in flight.  Keyframe and current-frame pyramids stay resident in device
// slots.  All float arithmetic is written with explicit round-to-nearest intrinsics (no fma contraction), doubles
// where the reference uses doubles (bilinear warp, interpolation weights).
#include "common.cuh"
#include <vector>

namespace {
using namespace ygz_b200;

constexpr int DIR_MAX_LEVELS = 16;

struct DirectFrameDev {
    const uint8_t* level[DIR_MAX_LEVELS];     // tight rows (pitch == w)
    int w[DIR_MAX_LEVELS], h[DIR_MAX_LEVELS];
    float q[4], t[3];
    int nlevels;
};

struct DirectParams {
    float fx, fy, cx, cy;
    float scale[DIR_MAX_LEVELS], inv_scale[DIR_MAX_LEVELS];
    float inv_level_sigma2_1;
    int nlevels;
    int cur_slot;
    int n_slots;
};

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }

struct SE3d { float q[4]; float t[3]; };

__device__ inline void quat_mul(const float* a, const float* b, float* o) {   // Eigen quaternion product a*b
    const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = fs(fs(fs(fm(aw, bw), fm(ax, bx)), fm(ay, by)), fm(az, bz));
    o[0] = fs(fa(fa(fm(aw, bx), fm(ax, bw)), fm(ay, bz)), fm(az, by));
    o[1] = fs(fa(fa(fm(aw, by), fm(ay, bw)), fm(az, bx)), fm(ax, bz));
    o[2] = fs(fa(fa(fm(aw, bz), fm(az, bw)), fm(ax, by)), fm(ay, bx));
}

__device__ inline void quat_rotate(const float* q, const float* v, float* o) {   // Eigen _transformVector
    float uv[3] = {fs(fm(q[1], v[2]), fm(q[2], v[1])), fs(fm(q[2], v[0]), fm(q[0], v[2])), fs(fm(q[0], v[1]), fm(q[1], v[0]))};
    uv[0] = fa(uv[0], uv[0]); uv[1] = fa(uv[1], uv[1]); uv[2] = fa(uv[2], uv[2]);
    o[0] = fa(fa(v[0], fm(q[3], uv[0])), fs(fm(q[1], uv[2]), fm(q[2], uv[1])));
    o[1] = fa(fa(v[1], fm(q[3], uv[1])), fs(fm(q[2], uv[0]), fm(q[0], uv[2])));
    o[2] = fa(fa(v[2], fm(q[3], uv[2])), fs(fm(q[0], uv[1]), fm(q[1], uv[0])));
}

__device__ inline void se3_apply(const SE3d& T, const float* p, float* o) {   // se3.hpp:257-259
    quat_rotate(T.q, p, o);
    o[0] = fa(o[0], T.t[0]); o[1] = fa(o[1], T.t[1]); o[2] = fa(o[2], T.t[2]);
}

__device__ inline void quat_normalize(float* q) {
    const float len = __fsqrt_rn(fa(fa(fa(fm(q[0], q[0]), fm(q[1], q[1])), fm(q[2], q[2])), fm(q[3], q[3])));
    for (int i = 0; i < 4; i++) q[i] = fd(q[i], len);
}

__device__ inline SE3d se3_mul(const SE3d& a, const SE3d& b) {   // se3.hpp:239-243, 268-276
    SE3d r;
    float rt[3];
    quat_rotate(a.q, b.t, rt);
    r.t[0] = fa(a.t[0], rt[0]); r.t[1] = fa(a.t[1], rt[1]); r.t[2] = fa(a.t[2], rt[2]);
    quat_mul(a.q, b.q, r.q);
    quat_normalize(r.q);
    return r;
}

__device__ inline SE3d se3_inv(const SE3d& a) {   // se3.hpp:169-174
    SE3d r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    quat_normalize(r.q);
    float nt[3] = {fm(a.t[0], -1.f), fm(a.t[1], -1.f), fm(a.t[2], -1.f)};
    quat_rotate(r.q, nt, r.t);
    return r;
}

__device__ inline void camera2pixel(const DirectParams& P, const float* p, float* px) {   // KeyFrame.h:173-178
    px[0] = fa(fd(fm(P.fx, p[0]), p[2]), P.cx);
    px[1] = fa(fd(fm(P.fy, p[1]), p[2]), P.cy);
}

__device__ inline void pixel2camera(const DirectParams& P, const float* px, float depth, float* p) {   // KeyFrame.h:180-186
    p[0] = fd(fm(fs(px[0], P.cx), depth), P.fx);
    p[1] = fd(fm(fs(px[1], P.cy), depth), P.fy);
    p[2] = depth;
}

// include/ORBmatcher.h:200-211 (double arithmetic, truncation to uchar)
__device__ inline uint8_t bilinear_uchar(double x, double y, const uint8_t* data0, int step) {
    const double xx = __dsub_rn(x, floor(x));
    const double yy = __dsub_rn(y, floor(y));
    const uint8_t* data = data0 + (int)y * step + (int)x;
    const double ox = __dsub_rn(1.0, xx), oy = __dsub_rn(1.0, yy);
    double s = __dmul_rn(__dmul_rn(ox, oy), (double)data[0]);
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(xx, oy), (double)data[1]));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(ox, yy), (double)data[step]));
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(xx, yy), (double)data[step + 1]));
    return (uint8_t)(int)s;
}

__device__ inline float cof3(const float (*m)[3], int i, int j) {   // Eigen cofactor_3x3<i,j>
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return fs(fm(m[i1][j1], m[i2][j2]), fm(m[i1][j2], m[i2][j1]));
}

__global__ void __launch_bounds__(128) k_direct_projection(const DirectFrameDev* __restrict__ frames, DirectParams P,
                                                           const ygz_direct_item* __restrict__ items, int n_items,
                                                           float* __restrict__ out_px, int* __restrict__ out_level,
                                                           uint8_t* __restrict__ out_ok) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_items) return;
    const ygz_direct_item it = items[n];
    const DirectFrameDev& R = frames[it.ref];
    const DirectFrameDev& C = frames[P.cur_slot];
    SE3d pose_ref, pose_cur;
    for (int i = 0; i < 4; i++) { pose_ref.q[i] = R.q[i]; pose_cur.q[i] = C.q[i]; }
    for (int i = 0; i < 3; i++) { pose_ref.t[i] = R.t[i]; pose_cur.t[i] = C.t[i]; }
    const SE3d TCR = se3_mul(pose_cur, se3_inv(pose_ref));                        // :1584
    // ---- GetWarpAffineMatrix (:1525-1547)
    const float WH = 4.f;                                                          // WarpHalfPatchSize
    float pt_ref[3];
    se3_apply(pose_ref, it.pw, pt_ref);
    const float depth = pt_ref[2];
    const float sl = P.scale[it.octave];
    const float pdu[2] = {fa(it.px_ref[0], fm(WH, sl)), fa(it.px_ref[1], fm(0.f, sl))};
    const float pdv[2] = {fa(it.px_ref[0], fm(0.f, sl)), fa(it.px_ref[1], fm(WH, sl))};
    float pt_du[3], pt_dv[3], pc[3], px_cur[2], px_du[2], px_dv[2];
    pixel2camera(P, pdu, depth, pt_du);
    pixel2camera(P, pdv, depth, pt_dv);
    se3_apply(TCR, pt_ref, pc); camera2pixel(P, pc, px_cur);
    se3_apply(TCR, pt_du, pc); camera2pixel(P, pc, px_du);
    se3_apply(TCR, pt_dv, pc); camera2pixel(P, pc, px_dv);
    const float A00c = fd(fs(px_du[0], px_cur[0]), WH), A10c = fd(fs(px_du[1], px_cur[1]), WH);
    const float A01c = fd(fs(px_dv[0], px_cur[0]), WH), A11c = fd(fs(px_dv[1], px_cur[1]), WH);
    // ---- GetBestSearchLevel (ORBmatcher.h:185-197)
    int search_level = 0;
    const float det = fs(fm(A00c, A11c), fm(A10c, A01c));
    float D = det;
    while ((double)D > 3.0 && search_level < P.nlevels - 1) {
        search_level += 1;
        D = fm(D, P.inv_level_sigma2_1);
    }
    // ---- WarpAffine (:1549-1573) with half_patch_size 5: 10 x 10 patch with border
    uint8_t pwb[100];
    {
        const float invdet = fd(1.f, det);                                         // Matrix2f::inverse()
        const float A00 = fm(A11c, invdet), A10 = fm(-A10c, invdet), A01 = fm(-A01c, invdet), A11 = fm(A00c, invdet);
        const float prx = fd(it.px_ref[0], sl), pry = fd(it.px_ref[1], sl);
        const uint8_t* img = R.level[it.octave];
        const int cols = R.w[it.octave], rows = R.h[it.octave];
        const float ss = P.scale[search_level];
        for (int y = 0; y < 10; y++)
            for (int x = 0; x < 10; x++) {
                const float qx = fm((float)(x - 5), ss), qy = fm((float)(y - 5), ss);
                const float px0 = fa(fa(fm(A00, qx), fm(A01, qy)), prx);
                const float px1 = fa(fa(fm(A10, qx), fm(A11, qy)), pry);
                uint8_t v = 0;
                if (!(px0 < 0 || px1 < 0 || px0 >= (float)(cols - 1) || px1 >= (float)(rows - 1)))
                    v = bilinear_uchar((double)px0, (double)px1, img, cols);
                pwb[y * 10 + x] = v;
            }
    }
    // ---- Align2D (src/Align.cc:8-103) on curr->mvImagePyramid[search_level]
    float u = fm(it.px_cur[0], P.inv_scale[search_level]), v = fm(it.px_cur[1], P.inv_scale[search_level]);
    bool converged = false;
    {
        float dxs[64], dys[64];
        float H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                const uint8_t* p = pwb + (y + 1) * 10 + 1 + x;
                const float J0 = 0.5f * (float)((int)p[1] - (int)p[-1]);          // exact: half-integers
                const float J1 = 0.5f * (float)((int)p[10] - (int)p[-10]);
                dxs[y * 8 + x] = J0; dys[y * 8 + x] = J1;
                // H += J J^T: all partial sums are multiples of 0.25 below 2^22 — exact in float in any order
                H[0][0] = fa(H[0][0], fm(J0, J0)); H[0][1] = fa(H[0][1], fm(J0, J1)); H[0][2] = fa(H[0][2], J0);
                H[1][1] = fa(H[1][1], fm(J1, J1)); H[1][2] = fa(H[1][2], J1); H[2][2] = fa(H[2][2], 1.f);
            }
        H[1][0] = H[0][1]; H[2][0] = H[0][2]; H[2][1] = H[1][2];
        float Hi[3][3];
        {   // Matrix3f::inverse(): cofactors, det = c00*m00 + (c10*m10 + c20*m20) (Eigen's unrolled redux)
            const float c00 = cof3(H, 0, 0), c10 = cof3(H, 1, 0), c20 = cof3(H, 2, 0);
            const float d3 = fa(fm(c00, H[0][0]), fa(fm(c10, H[0][1]), fm(c20, H[0][2])));
            const float id = fd(1.f, d3);
            Hi[0][0] = fm(c00, id); Hi[0][1] = fm(c10, id); Hi[0][2] = fm(c20, id);
            Hi[1][0] = fm(cof3(H, 0, 1), id); Hi[1][1] = fm(cof3(H, 1, 1), id); Hi[1][2] = fm(cof3(H, 2, 1), id);
            Hi[2][0] = fm(cof3(H, 0, 2), id); Hi[2][1] = fm(cof3(H, 1, 2), id); Hi[2][2] = fm(cof3(H, 2, 2), id);
        }
        const uint8_t* cur = C.level[search_level];
        const int cols = C.w[search_level], rows = C.h[search_level];
        float mean_diff = 0.f;
        const float min_update_squared = (float)(0.03 * 0.03);
        bool nan_exit = false;
        const float u0 = u, v0 = v;
        for (int iter = 0; iter < 10; iter++) {
            const int u_r = (int)floorf(u), v_r = (int)floorf(v);
            if (u_r < 4 || v_r < 4 || u_r >= cols - 4 || v_r >= rows - 4) break;
            if (isnan(u) || isnan(v)) { nan_exit = true; break; }
            const float sx = fs(u, (float)u_r), sy = fs(v, (float)v_r);
            const double osx = __dsub_rn(1.0, (double)sx), osy = __dsub_rn(1.0, (double)sy);
            const float wTL = __double2float_rn(__dmul_rn(osx, osy));
            const float wTR = __double2float_rn(__dmul_rn((double)sx, osy));
            const float wBL = __double2float_rn(__dmul_rn(osx, (double)sy));
            const float wBR = fm(sx, sy);
            float J0 = 0.f, J1 = 0.f, J2 = 0.f;
            for (int y = 0; y < 8; y++) {
                const uint8_t* p = cur + (size_t)(v_r + y - 4) * cols + u_r - 4;
                for (int x = 0; x < 8; x++, p++) {
                    const float sp = fa(fa(fa(fm(wTL, (float)p[0]), fm(wTR, (float)p[1])), fm(wBL, (float)p[cols])), fm(wBR, (float)p[cols + 1]));
                    const float res = fa(fs(sp, (float)pwb[(y + 1) * 10 + 1 + x]), mean_diff);
                    J0 = fs(J0, fm(res, dxs[y * 8 + x]));
                    J1 = fs(J1, fm(res, dys[y * 8 + x]));
                    J2 = fs(J2, res);
                }
            }
            const float up0 = fa(fm(Hi[0][0], J0), fa(fm(Hi[0][1], J1), fm(Hi[0][2], J2)));
            const float up1 = fa(fm(Hi[1][0], J0), fa(fm(Hi[1][1], J1), fm(Hi[1][2], J2)));
            const float up2 = fa(fm(Hi[2][0], J0), fa(fm(Hi[2][1], J1), fm(Hi[2][2], J2)));
            u = fa(u, up0); v = fa(v, up1); mean_diff = fa(mean_diff, up2);
            if (fa(fm(up0, up0), fm(up1, up1)) < min_update_squared) { converged = true; break; }
        }
        if (nan_exit) { u = u0; v = v0; converged = false; }     // `return false` leaves cur_px_estimate untouched
    }
    out_px[2 * n] = fm(u, P.scale[search_level]);
    out_px[2 * n + 1] = fm(v, P.scale[search_level]);
    out_level[n] = search_level;
    out_ok[n] = converged ? 1 : 0;
}

}  // namespace

struct ygz_direct {
    int device = 0;
    cudaStream_t stream = nullptr;
    int max_slots = 0, max_items = 0;
    size_t slot_bytes = 0;
    uint8_t* d_arena = nullptr;
    DirectFrameDev* d_frames = nullptr;
    std::vector<DirectFrameDev> frames;
    std::vector<int> valid;
    ygz_direct_item* d_items = nullptr;
    float* d_px = nullptr;
    int* d_level = nullptr;
    uint8_t* d_ok = nullptr;
    DirectParams P;
    int last_n = 0;
};

static int direct_free(ygz_direct* h) {
    if (!h) return YGZ_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_arena); cudaFree(h->d_frames); cudaFree(h->d_items); cudaFree(h->d_px); cudaFree(h->d_level); cudaFree(h->d_ok);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return YGZ_OK;
}

extern "C" {

int ygz_direct_create(int device, int max_slots, int max_items, size_t max_pyramid_bytes, ygz_direct** out) {
    YGZ_ARG_CHECK(out && max_slots >= 2 && max_items >= 1 && max_pyramid_bytes >= 64);
    if (ygz_device_count() <= device) { set_last_error("no such CUDA device"); return YGZ_ERR_NO_DEVICE; }
    YGZ_CUDA_TRY(cudaSetDevice(device));
    ygz_direct* h = new ygz_direct();
    h->device = device; h->max_slots = max_slots; h->max_items = max_items;
    h->slot_bytes = (max_pyramid_bytes + 255) / 256 * 256;
    h->frames.resize(max_slots);
    h->valid.assign(max_slots, 0);
    std::memset(h->frames.data(), 0, sizeof(DirectFrameDev) * max_slots);
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); direct_free(h); return YGZ_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY_OR_FREE(cudaMalloc(&h->d_arena, h->slot_bytes * max_slots));
    TRY_OR_FREE(cudaMalloc(&h->d_frames, sizeof(DirectFrameDev) * max_slots));
    TRY_OR_FREE(cudaMalloc(&h->d_items, sizeof(ygz_direct_item) * max_items));
    TRY_OR_FREE(cudaMalloc(&h->d_px, sizeof(float) * 2 * max_items));
    TRY_OR_FREE(cudaMalloc(&h->d_level, sizeof(int) * max_items));
    TRY_OR_FREE(cudaMalloc(&h->d_ok, max_items));
#undef TRY_OR_FREE
    *out = h;
    return YGZ_OK;
}

int ygz_direct_destroy(ygz_direct* h) { return direct_free(h); }

int ygz_direct_upload_frame(ygz_direct* h, int slot, const ygz_direct_frame* f, int nlevels) {
    YGZ_ARG_CHECK(h && f && slot >= 0 && slot < h->max_slots && nlevels >= 1 && nlevels <= DIR_MAX_LEVELS);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    size_t need = 0;
    for (int l = 0; l < nlevels; l++) {
        YGZ_ARG_CHECK(f->level[l] && f->w[l] >= 1 && f->h[l] >= 1 && f->pitch[l] >= f->w[l]);
        need += ((size_t)f->w[l] * f->h[l] + 15) / 16 * 16;
    }
    if (need > h->slot_bytes) { set_last_error("pyramid larger than max_pyramid_bytes"); return YGZ_ERR_ARG; }
    DirectFrameDev& D = h->frames[slot];
    uint8_t* base = h->d_arena + h->slot_bytes * slot;
    size_t off = 0;
    for (int l = 0; l < nlevels; l++) {
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(base + off, f->w[l], f->level[l], f->pitch[l], f->w[l], f->h[l], cudaMemcpyHostToDevice, h->stream));
        D.level[l] = base + off; D.w[l] = f->w[l]; D.h[l] = f->h[l];
        off += ((size_t)f->w[l] * f->h[l] + 15) / 16 * 16;
    }
    for (int i = 0; i < 4; i++) D.q[i] = f->q[i];
    for (int i = 0; i < 3; i++) D.t[i] = f->t[i];
    D.nlevels = nlevels;
    h->valid[slot] = 1;
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_frames + slot, &D, sizeof(D), cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));        // D and the host images may go out of scope
    return YGZ_OK;
}

int ygz_direct_set_pose(ygz_direct* h, int slot, const float* q, const float* t) {
    YGZ_ARG_CHECK(h && q && t && slot >= 0 && slot < h->max_slots && h->valid[slot]);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    DirectFrameDev& D = h->frames[slot];
    for (int i = 0; i < 4; i++) D.q[i] = q[i];
    for (int i = 0; i < 3; i++) D.t[i] = t[i];
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_frames + slot, &D, sizeof(D), cudaMemcpyHostToDevice, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

static int direct_launch(ygz_direct* h, int n) {
    { NvtxRange nvtx_range("ygz/FindDirectProjection");
    k_direct_projection<<<ceil_div(n, 128), 128, 0, h->stream>>>(h->d_frames, h->P, h->d_items, n, h->d_px, h->d_level, h->d_ok);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

int ygz_find_direct_projection(ygz_direct* h, int cur_slot, const float* K4, const float* scale, const float* inv_scale,
                               float inv_level_sigma2_1, int nlevels, const ygz_direct_item* items, int n_items,
                               float* out_px, int* out_level, uint8_t* out_ok) {
    YGZ_ARG_CHECK(h && K4 && scale && inv_scale && items && out_px && out_level && out_ok);
    YGZ_ARG_CHECK(n_items >= 0 && n_items <= h->max_items && nlevels >= 1 && nlevels <= DIR_MAX_LEVELS);
    YGZ_ARG_CHECK(cur_slot >= 0 && cur_slot < h->max_slots && h->valid[cur_slot] && h->frames[cur_slot].nlevels >= nlevels);
    if (n_items == 0) return YGZ_OK;
    for (int i = 0; i < n_items; i++) {
        const ygz_direct_item& it = items[i];
        YGZ_ARG_CHECK(it.ref >= 0 && it.ref < h->max_slots && h->valid[it.ref] && it.octave >= 0 && it.octave < nlevels &&
                      it.octave < h->frames[it.ref].nlevels);
    }
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    DirectParams& P = h->P;
    P.fx = K4[0]; P.fy = K4[1]; P.cx = K4[2]; P.cy = K4[3];
    for (int l = 0; l < DIR_MAX_LEVELS; l++) { P.scale[l] = l < nlevels ? scale[l] : 1.f; P.inv_scale[l] = l < nlevels ? inv_scale[l] : 1.f; }
    P.inv_level_sigma2_1 = inv_level_sigma2_1;
    P.nlevels = nlevels; P.cur_slot = cur_slot; P.n_slots = h->max_slots;
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_items, items, sizeof(ygz_direct_item) * n_items, cudaMemcpyHostToDevice, h->stream));
    int rc = direct_launch(h, n_items);
    if (rc) return rc;
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_px, h->d_px, sizeof(float) * 2 * n_items, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_level, h->d_level, sizeof(int) * n_items, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(out_ok, h->d_ok, n_items, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->last_n = n_items;
    return YGZ_OK;
}

/* device-timed repeats of the last ygz_find_direct_projection launch (items resident) for bench.py */
int ygz_direct_time(ygz_direct* h, int iters, float* ms) {
    YGZ_ARG_CHECK(h && ms && iters >= 1 && h->last_n >= 1);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    YGZ_CUDA_TRY(cudaEventCreate(&e0));
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); set_last_error("cudaEventCreate failed"); return YGZ_ERR_CUDA; }
    int rc = YGZ_OK;
    cudaError_t ce = cudaEventRecord(e0, h->stream);
    for (int i = 0; i < iters && rc == YGZ_OK && ce == cudaSuccess; i++) rc = direct_launch(h, h->last_n);
    if (ce == cudaSuccess) ce = cudaEventRecord(e1, h->stream);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc == YGZ_OK && ce != cudaSuccess) { set_last_error(std::string("ygz_direct_time: ") + cudaGetErrorString(ce)); rc = YGZ_ERR_CUDA; }
    return rc;
}

}  // extern "C"
