// ORACLE — TEST INFRASTRUCTURE ONLY (see ygz_oracle.h).  CPU restatement of the YGZ "DSO" key-point selection that
// Frame::ExtractORB uses for frames that already carry direct-tracked points (src/Frame.cc:335-337), SURVEY.md §8(f)
// rank 2:
//   ORBextractor::ComputeKeyPointsDSOSingleLevel   /root/reference/src/ORBextractor.cc:1277-1381
//   ORBextractor::ShiTomasiScore                   :1152-1186
//   fast::fast_corner_detect_10 / _sse2            /root/reference/Thirdparty/fast/src/fast_10.cpp:9-3167,
//                                                  faster_corner_10_sse.cpp:15-201
//   ORBextractor::operator()(Frame*, ..., DSO_KEYPOINT)   :1031-1127
// FAST-10 is restated from its definition (>= 10 contiguous ring pixels all brighter than p + b or all darker than
// p - b; the upstream sources are machine-generated decision trees of that predicate) and PINNED against the
// reference's own code compiled from where it lies (oracle/_ref/libfast_ref.so, tests/test_fast10_oracle.py), including
// the region quirk: widths below 22 px go through the plain detector, which tests EVERY pixel of the w x h window
// (reading 3 px around it), while the SSE2 detector tests only [3, w-3) x [3, h-3).
// Oracle determinism rule: the per-cell `std::sort` by score (unstable for more than 16 corners) is a stable sort —
// equal Shi-Tomasi scores keep raster order.  No fma contraction.
#include "orb_extractor.h"
#include <algorithm>
#include <cmath>

namespace orc {

static const int RING[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

bool fast10_is_corner(const uint8_t* p, int stride, int b) {
    const int c = *p;
    unsigned brighter = 0, darker = 0;
    for (int k = 0; k < 16; k++) {
        const int v = p[RING[k][0] + RING[k][1] * stride];
        if (v > c + b) brighter |= 1u << k;
        if (v < c - b) darker |= 1u << k;
    }
    auto run10 = [](unsigned m) {
        m |= m << 16;                                  // circular
        unsigned r = m;
        for (int i = 1; i < 10; i++) r &= m >> i;
        return (r & 0xFFFFu) != 0;
    };
    return run10(brighter) || run10(darker);
}

// fast::fast_corner_detect_10_sse2 as the reference calls it on a cell (faster_corner_10_sse.cpp:188-201)
void fast10_detect_cell(const uint8_t* img, int w, int h, int stride, int barrier, std::vector<std::pair<int, int>>& out) {
    if (w < 22) {                                      // plain detector: every pixel of the window
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                if (fast10_is_corner(img + y * stride + x, stride, barrier)) out.emplace_back(x, y);
        return;
    }
    if (h < 7) return;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++)
            if (fast10_is_corner(img + y * stride + x, stride, barrier)) out.emplace_back(x, y);
}

// src/ORBextractor.cc:1152-1186
float Extractor::ShiTomasiScore(const Image& img, int u, int v) const {
    float dXX = 0.0, dYY = 0.0, dXY = 0.0;
    const int halfbox_size = 4, box_size = 8, box_area = 64;
    const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
    if (x_min < 1 || x_max >= img.w - 1 || y_min < 1 || y_max >= img.h - 1) return 0.0;
    const int stride = img.w;
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t* ptr_left = img.px.data() + stride * y + x_min - 1;
        const uint8_t* ptr_right = img.px.data() + stride * y + x_min + 1;
        const uint8_t* ptr_top = img.px.data() + stride * (y - 1) + x_min;
        const uint8_t* ptr_bottom = img.px.data() + stride * (y + 1) + x_min;
        for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
            float dx = (float)(*ptr_right - *ptr_left);
            float dy = (float)(*ptr_bottom - *ptr_top);
            dXX += dx * dx;
            dYY += dy * dy;
            dXY += dx * dy;
        }
    }
    dXX = (float)(dXX / (2.0 * box_area));
    dYY = (float)(dYY / (2.0 * box_area));
    dXY = (float)(dXY / (2.0 * box_area));
    return (float)(0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY))));
}

// src/ORBextractor.cc:1277-1381
void Extractor::ComputeKeyPointsDSOSingleLevel(std::vector<KeyPoint>& allKeypoints, std::vector<KeyPoint>& exist_kps) {
    const Image& img = pyramid[0];
    const int h = img.h, w = img.w, n = nfeatures;
    std::vector<uint8_t> occ((size_t)w * h, 0);
    for (const KeyPoint& kp : exist_kps) occ[(size_t)cv_round(kp.y) * w + cv_round(kp.x)] = 255;
    if (mnGridSize < 0) mnGridSize = static_cast<int>(std::sqrt(1.0 * h * w / n));
    allKeypoints.clear();
    int cnt = 0;
    while (cnt < n) {
        if (cnt > 0) {
            mnGridSize -= 5;
            if (mnGridSize < 7) { mnGridSize = 7; break; }
        }
        allKeypoints.clear();
        const int grid_n_rows = h / mnGridSize, grid_n_cols = w / mnGridSize;
        cnt = 0;
        for (int k = 0; k < grid_n_rows * grid_n_cols; k++) {
            const int nn = k / grid_n_cols;
            if (nn == 0 || nn == grid_n_rows - 1 || (k % grid_n_cols) == 0 || (k + 1) % grid_n_cols == 0) continue;
            const int x_start = (k - nn * grid_n_cols) * mnGridSize, y_start = nn * mnGridSize;
            const uint8_t* data = img.row(y_start) + x_start;
            std::vector<std::pair<int, int>> fast_corners;
            fast10_detect_cell(data, mnGridSize, mnGridSize, w, 20, fast_corners);
            if (fast_corners.empty()) fast10_detect_cell(data, mnGridSize, mnGridSize, w, 5, fast_corners);
            if (fast_corners.empty()) continue;
            std::vector<std::pair<std::pair<int, int>, float>> corner_score;
            for (auto& xy : fast_corners) {
                xy.first += x_start;
                xy.second += y_start;
                if (xy.first < 20 || xy.second < 20 || xy.first >= img.w - 20 || xy.second >= img.h - 20) continue;
                if (occ[(size_t)xy.second * w + xy.first] == 255) continue;
                corner_score.push_back(std::make_pair(xy, ShiTomasiScore(img, xy.first, xy.second)));
            }
            std::stable_sort(corner_score.begin(), corner_score.end(),
                             [](const std::pair<std::pair<int, int>, float>& c1, const std::pair<std::pair<int, int>, float>& c2) {
                                 return c1.second > c2.second;
                             });
            const int n_feature_in_grid = corner_score.size() > 3 ? 3 : (int)corner_score.size();
            for (int i = 0; i < n_feature_in_grid; i++) {
                KeyPoint kp;
                kp.x = (float)corner_score[i].first.first; kp.y = (float)corner_score[i].first.second;
                kp.size = 7; kp.response = 0; kp.octave = 0; kp.class_id = -1;
                kp.angle = IC_Angle(img, kp.x, kp.y);
                allKeypoints.push_back(kp);
                cnt++;
            }
        }
        if (cnt == 0) break;       // ORACLE RULE: the reference loops forever on a frame without a single corner (cnt stays 0)
    }
    if (cnt > n) mnGridSize += 5;
    for (KeyPoint& kp : exist_kps)
        kp.angle = IC_Angle(pyramid[kp.octave], kp.x * mvInvScaleFactor[kp.octave], kp.y * mvInvScaleFactor[kp.octave]);
}

// src/ORBextractor.cc:1031-1127 with method == DSO_KEYPOINT, left eye, pyramid already built
void Extractor::extract_frame_dso(std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors) {
    std::vector<KeyPoint> kps;
    ComputeKeyPointsDSOSingleLevel(kps, keypoints);
    const int nExisting = (int)keypoints.size();       // oracle's definition of Frame::N
    const int n = nExisting + (int)kps.size();
    descriptors.assign((size_t)n * 32, 0);
    blur_all_public();
    for (int i = 0; i < nExisting; i++) {
        KeyPoint tmp = keypoints[i];
        tmp.x *= mvInvScaleFactor[tmp.octave];
        tmp.y *= mvInvScaleFactor[tmp.octave];
        computeOrbDescriptor(tmp, blurred[tmp.octave], &descriptors[(size_t)i * 32]);
    }
    for (size_t i = 0; i < kps.size(); i++) computeOrbDescriptor(kps[i], blurred[0], &descriptors[(size_t)(nExisting + i) * 32]);
    keypoints.insert(keypoints.end(), kps.begin(), kps.end());
}

}  // namespace orc
