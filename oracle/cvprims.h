// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// CPU restatements of the OpenCV primitives the reference hot path calls.  OpenCV is a
// third-party dependency whose sources are NOT under /root/reference; the reference only
// pins a floor (find_package(OpenCV 3.0) else 2.4.3, CMakeLists.txt:39-45).  Parity is
// pinned to the behaviour of opencv-python-headless 4.13.0 (cv2) in the build container;
// tests/test_oracle_vs_cv2.py checks every primitive here against cv2 bit-for-bit.
//
//   resize_linear_u8   <- cv::resize(..., INTER_LINEAR) on CV_8UC1   (src/ORBextractor.cc:1139)
//   border_reflect101  <- cv::copyMakeBorder(..., BORDER_REFLECT_101) (src/ORBextractor.cc:1141-1146)
//   gauss7_sigma2_u8   <- cv::GaussianBlur(7x7, sigma 2, REFLECT_101) (src/ORBextractor.cc:1010,1083)
//   fast9_16           <- cv::FAST(img, kps, threshold, true)         (src/ORBextractor.cc:765-769)
//   fast_atan2         <- cv::fastAtan2                               (src/ORBextractor.cc:100)
//   cv_round           <- cvRound (round-half-even)                   (src/ORBextractor.cc:80,111,115)
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>

namespace orc {

struct KeyPoint {           // field order and size (28 B) of cv::KeyPoint
    float x, y, size, angle, response;
    int octave, class_id;
};

static inline int cv_round(float v) { return (int)lrintf(v); }   // cvtss2si, half-to-even
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(float v) { int i = (int)v; return i - (v < (float)i); }

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride);

// Per-axis tables exactly as cv::resize builds them (xofs/ialpha): also used by the GPU
// side tests to check the host-side table builder.
void resize_linear_tables(int ssize, int dsize, std::vector<int>& ofs, std::vector<short>& coef);

void border_reflect101(const uint8_t* src, int w, int h, int sstride,
                       uint8_t* dst, int border, int dstride);

void gauss7_sigma2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);

// cv::FAST 9/16 on a (sub)image.  Appends keypoints (x, y local to the sub-image,
// size 7, angle -1, response = score) in raster order.
void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nms,
              std::vector<KeyPoint>& out);

// score of one pixel (max threshold for which it is still a corner), cornerScore<16>.
int fast9_16_score(const uint8_t* p, int stride, int threshold);

float fast_atan2(float y, float x);

}  // namespace orc
