// ORACLE — TEST INFRASTRUCTURE ONLY (see cvprims.h header comment).
// Restatement of the published OpenCV 4.x algorithms (imgproc/resize.cpp fixed-point
// bilinear, imgproc/smooth fixed-point Gaussian, features2d/fast.cpp + fast_score.cpp,
// core/mathfuncs_core fastAtan2).  OpenCV sources are not vendored in /root/reference;
// behaviour pinned to cv2 4.13.0 by tests/test_oracle_vs_cv2.py.
#include "cvprims.h"
#include <algorithm>
#include <cfloat>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------- resize (INTER_LINEAR, 8U)
static const int COEF_BITS = 11;
static const int COEF_ONE = 1 << COEF_BITS;

void resize_linear_tables(int ssize, int dsize, std::vector<int>& ofs, std::vector<short>& coef) {
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1.0 / inv_scale;
    ofs.resize(dsize);
    coef.resize(2 * dsize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        // saturate_cast<short>(float) == cvRound
        coef[2 * d] = (short)cv_round(c0 * COEF_ONE);
        coef[2 * d + 1] = (short)cv_round(c1 * COEF_ONE);
    }
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride) {
    std::vector<int> xofs;
    std::vector<short> ialpha;
    resize_linear_tables(sw, dw, xofs, ialpha);
    // vertical: cv keeps the fractional weights and clips the row indices instead
    const double scale_y = 1.0 / ((double)dh / sh);
    static thread_local std::vector<int> r0, r1;
    if (r0.size() < (size_t)dw) { r0.resize(dw); r1.resize(dw); }
    int prev0 = -1, prev1 = -1;
    auto hrow = [&](int sy, std::vector<int>& out) {
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            if (sx >= sw - 1) out[dx] = S[sx] * COEF_ONE;
            else out[dx] = S[sx] * ialpha[2 * dx] + S[sx + 1] * ialpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        short b0 = (short)cv_round((1.f - fy) * COEF_ONE);
        short b1 = (short)cv_round(fy * COEF_ONE);
        int y0 = std::min(std::max(sy, 0), sh - 1);
        int y1 = std::min(std::max(sy + 1, 0), sh - 1);
        if (y0 != prev0) { hrow(y0, r0); prev0 = y0; }
        if (y1 != prev1) { hrow(y1, r1); prev1 = y1; }
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int x = 0; x < dw; x++)
            D[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------- border
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

void border_reflect101(const uint8_t* src, int w, int h, int sstride,
                       uint8_t* dst, int border, int dstride) {
    for (int y = -border; y < h + border; y++) {
        const uint8_t* S = src + (size_t)reflect101(y, h) * sstride;
        uint8_t* D = dst + (size_t)(y + border) * dstride;
        for (int x = -border; x < w + border; x++) D[x + border] = S[reflect101(x, w)];
    }
}

// ---------------------------------------------------------------- GaussianBlur 7x7 sigma 2 (8U fixed point)
// getGaussianKernelBitExact(7, 2) in 8.8 fixed point = {18,34,48,56,48,34,18}/256.
static const int G7[7] = {18, 34, 48, 56, 48, 34, 18};

void gauss7_sigma2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    // horizontal pass into 8.8 fixed point (u16), rows padded by reflect-101 so the inner loop is branch-free
    // scratch reused per thread: fresh multi-100KB vectors per call go through mmap and serialise many-thread runs
    static thread_local std::vector<uint16_t> hbuf;
    static thread_local std::vector<uint8_t> prow;
    if (hbuf.size() < (size_t)w * h) hbuf.resize((size_t)w * h);
    if (prow.size() < (size_t)w + 6) prow.resize((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = -3; x < w + 3; x++) prow[x + 3] = S[reflect101(x, w)];
        uint16_t* H = hbuf.data() + (size_t)y * w;
        const uint8_t* P = prow.data();
        for (int x = 0; x < w; x++)
            H[x] = (uint16_t)(G7[0] * (P[x] + P[x + 6]) + G7[1] * (P[x + 1] + P[x + 5]) +
                              G7[2] * (P[x + 2] + P[x + 4]) + G7[3] * P[x + 3]);   // <= 255*256
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = -3; k <= 3; k++) R[k + 3] = hbuf.data() + (size_t)reflect101(y + k, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = (uint32_t)G7[0] * ((uint32_t)R[0][x] + R[6][x]) + (uint32_t)G7[1] * ((uint32_t)R[1][x] + R[5][x]) +
                           (uint32_t)G7[2] * ((uint32_t)R[2][x] + R[4][x]) + (uint32_t)G7[3] * R[3][x];
            D[x] = (uint8_t)((acc + (1u << 15)) >> 16);
        }
    }
}

// ---------------------------------------------------------------- FAST 9/16
static const int RING[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

int fast9_16_score(const uint8_t* p, int stride, int threshold) {
    const int K = 8, N = K * 3 + 1;
    int d[N];
    const int v = p[0];
    for (int k = 0; k < N; k++) d[k] = v - p[RING[k & 15][0] + RING[k & 15][1] * stride];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]);
        a = std::min(a, d[k + 5]);
        a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]);
        a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]);
        b = std::max(b, d[k + 4]);
        b = std::max(b, d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 6]);
        b = std::max(b, d[k + 7]);
        b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

// corner test of cv::FAST_t<16>: threshold table + the 8 opposite-pair pre-tests, then the
// contiguous-run count (K = 8, N = 25).
static inline bool fast9_is_corner(const uint8_t* p, const int* pixel, const uint8_t* tab0, int t) {
    const int v = p[0];
    const uint8_t* tab = tab0 - v + 255;
    int d = tab[p[pixel[0]]] | tab[p[pixel[8]]];
    if (d == 0) return false;
    d &= tab[p[pixel[2]]] | tab[p[pixel[10]]];
    d &= tab[p[pixel[4]]] | tab[p[pixel[12]]];
    d &= tab[p[pixel[6]]] | tab[p[pixel[14]]];
    if (d == 0) return false;
    d &= tab[p[pixel[1]]] | tab[p[pixel[9]]];
    d &= tab[p[pixel[3]]] | tab[p[pixel[11]]];
    d &= tab[p[pixel[5]]] | tab[p[pixel[13]]];
    d &= tab[p[pixel[7]]] | tab[p[pixel[15]]];
    if (d & 1) {
        const int vt = v - t;
        int count = 0;
        for (int k = 0; k < 25; k++) {
            if (p[pixel[k]] < vt) { if (++count > 8) return true; }
            else count = 0;
        }
    }
    if (d & 2) {
        const int vt = v + t;
        int count = 0;
        for (int k = 0; k < 25; k++) {
            if (p[pixel[k]] > vt) { if (++count > 8) return true; }
            else count = 0;
        }
    }
    return false;
}

void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nms,
              std::vector<KeyPoint>& out) {
    if (w < 7 || h < 7) return;
    threshold = std::min(std::max(threshold, 0), 255);
    // three rolling rows of scores, zero where no corner (as cv::FAST_t<16>)
    int pixel[25];
    for (int k = 0; k < 25; k++) pixel[k] = RING[k & 15][0] + RING[k & 15][1] * stride;
    uint8_t tab0[512];
    for (int i = -255; i <= 255; i++) tab0[i + 255] = (uint8_t)(i < -threshold ? 1 : i > threshold ? 2 : 0);
    static thread_local std::vector<uint8_t> sbuf;
    static thread_local std::vector<int> cbuf;
    sbuf.assign((size_t)3 * w, 0);
    cbuf.assign((size_t)3 * (w + 1), 0);
    uint8_t* buf[3] = {sbuf.data(), sbuf.data() + w, sbuf.data() + 2 * w};
    int* cpbuf[3] = {cbuf.data() + 1, cbuf.data() + (w + 1) + 1, cbuf.data() + 2 * (w + 1) + 1};
    for (int i = 3; i < h - 2; i++) {
        const uint8_t* ptr = img + (size_t)i * stride;
        uint8_t* curr = buf[(i - 3) % 3];
        int* cornerpos = cpbuf[(i - 3) % 3];
        std::memset(curr, 0, w);
        int ncorners = 0;
        if (i < h - 3) {
            for (int j = 3; j < w - 3; j++) {
                if (fast9_is_corner(ptr + j, pixel, tab0, threshold)) {
                    cornerpos[ncorners++] = j;
                    if (nms) curr[j] = (uint8_t)fast9_16_score(ptr + j, stride, threshold);
                }
            }
        }
        cornerpos[-1] = ncorners;
        if (i == 3) continue;
        const uint8_t* prev = buf[(i - 4 + 3) % 3];
        const uint8_t* pprev = buf[(i - 5 + 3) % 3];
        cornerpos = cpbuf[(i - 4 + 3) % 3];
        ncorners = cornerpos[-1];
        for (int k = 0; k < ncorners; k++) {
            int j = cornerpos[k];
            int score = prev[j];
            if (!nms || (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] &&
                         score > pprev[j] && score > pprev[j + 1] && score > curr[j - 1] &&
                         score > curr[j] && score > curr[j + 1])) {
                KeyPoint kp;
                kp.x = (float)j; kp.y = (float)(i - 1); kp.size = 7.f; kp.angle = -1.f;
                kp.response = (float)score; kp.octave = 0; kp.class_id = -1;
                out.push_back(kp);
            }
        }
    }
}

// ---------------------------------------------------------------- fastAtan2 (degrees)
float fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace orc
