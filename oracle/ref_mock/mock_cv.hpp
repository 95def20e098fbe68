// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// A minimal stand-in for the slice of the OpenCV C++ API that the reference's hot-path sources use, so that
// /root/reference/src/ORBextractor.cc (and the matcher / aligner sources) can be compiled UNMODIFIED, from where they
// lie, into oracle/_ref/*.so and serve as the pin for the oracle restatement (oracle/Makefile: _ref/libextractor_ref.so).
// OpenCV's headers are not in this image; its *behaviour* is what cv2 4.13.0 does, and the image
// primitives here forward to oracle/cvprims.cpp, which tests/test_oracle_vs_cv2.py pins to cv2 bit for bit:
//   cv::resize(INTER_LINEAR, 8U)  -> orc::resize_linear_u8        cv::GaussianBlur(7x7, 2, 2) -> orc::gauss7_sigma2_u8
//   cv::FAST(img, kps, t, true)   -> orc::fast9_16                cv::fastAtan2              -> orc::fast_atan2
//   cv::copyMakeBorder(REFLECT_101 [+ISOLATED]) : index arithmetic, written out below (in place safe)
// Everything else (Mat headers / ROI / reference counting, Point_, Size, Rect, KeyPoint, Input/OutputArray) is
// container plumbing with OpenCV's documented semantics; in particular `m = Mat::zeros(r, c, t)` on a Mat that already
// has that geometry clears it IN PLACE (MatExpr assignment = create() + setTo(0)), which computeDescriptors
// (src/ORBextractor.cc:964) relies on to write through a rowRange() header into the caller's descriptor matrix.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#if defined(__SSE4_2__) || defined(__POPCNT__)
#include <nmmintrin.h>      // OpenCV's cvdef.h pulls the SSE intrinsics in; src/ORBmatcher.cc:1515 relies on it for _mm_popcnt_u64
#endif

#include "../cvprims.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_32FC1 5

static inline int cvRound(double v) { return (int)lrint(v); }       // cvtsd2si: round half to even
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (v < i); }
static inline int cvFloor(float v) { int i = (int)v; return i - (v < i); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int)v; return i + (v > i); }
static inline int cvCeil(float v) { int i = (int)v; return i + (v > i); }
static inline int cvCeil(int v) { return v; }

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <class T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <class T> static inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <class T> static inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <class T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Scalar { double v[4]; Scalar(double a = 0) { v[0] = a; v[1] = v[2] = v[3] = 0; } };

struct KeyPoint {       // field order and size (28 B) of cv::KeyPoint
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct MatStep {
    size_t buf[2];
    size_t* p;
    MatStep() : p(buf) { buf[0] = buf[1] = 0; }
    MatStep(const MatStep& o) : p(buf) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; }
    MatStep& operator=(const MatStep& o) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; return *this; }
    operator size_t() const { return buf[0]; }
    size_t operator[](int i) const { return buf[i]; }
};

static inline int mock_elem_size(int type) {
    switch (type & 7) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_64F: return 8; default: return 4; }
}

class Mat;
struct MatZeros { int rows, cols, type; };     // what Mat::zeros returns (OpenCV: a MatExpr)

class Mat {
public:
    int flags, dims, rows, cols;
    uchar* data;
    MatStep step;
    std::shared_ptr<uchar> holder;              // owning buffer (null for user data)

    Mat() : flags(0), dims(0), rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int t) : Mat() { create(r, c, t); }
    Mat(Size s, int t) : Mat() { create(s.height, s.width, t); }
    Mat(Size s, int t, const Scalar& v) : Mat() { create(s.height, s.width, t); setTo(v); }      // byte fill: only called with Scalar(0)
    Mat(int r, int c, int t, void* d, size_t st = 0) : flags(t), dims(2), rows(r), cols(c), data((uchar*)d) {
        step.buf[1] = mock_elem_size(t);
        step.buf[0] = st ? st : (size_t)c * step.buf[1];
    }
    Mat(const MatZeros& z) : Mat() { *this = z; }
    Mat(const Mat&) = default;
    Mat& operator=(const Mat&) = default;
    Mat& operator=(const MatZeros& z) {
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step.buf[0], 0, (size_t)cols * step.buf[1]);
        return *this;
    }
    Mat& operator=(const Scalar& s) { return setTo(s); }
    Mat& setTo(const Scalar& s) {
        for (int y = 0; y < rows; y++) std::memset(data + (size_t)y * step.buf[0], (int)s.v[0], (size_t)cols * step.buf[1]);
        return *this;
    }

    void create(int r, int c, int t) {
        if (data && rows == r && cols == c && type() == t) return;      // same geometry: keep the buffer (OpenCV)
        flags = t; dims = 2; rows = r; cols = c;
        step.buf[1] = mock_elem_size(t);
        step.buf[0] = (size_t)c * step.buf[1];
        size_t n = (size_t)r * step.buf[0];
        holder.reset((uchar*)std::malloc(n ? n : 1), std::free);
        data = holder.get();
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { holder.reset(); data = nullptr; rows = cols = 0; dims = 0; }

    static MatZeros zeros(int r, int c, int t) { return MatZeros{r, c, t}; }
    // Mat::ones / convertTo(CV_32F): the 11x11 float patches of Frame::ComputeStereoMatches (src/Frame.cc:604-627)
    static Mat ones(int r, int c, int t) {
        Mat m(r, c, t);
        for (int y = 0; y < r; y++) for (int x = 0; x < c; x++) { if ((t & 7) == CV_32F) m.at<float>(y, x) = 1.f; else m.data[(size_t)y * m.step.buf[0] + x] = 1; }
        return m;
    }
    void convertTo(Mat& dst, int rtype) const {             // 8U -> 32F (or a copy); dst may alias *this (a ROI header)
        Mat out(rows, cols, rtype);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                const double v = depth() == CV_32F ? (double)at<float>(y, x) : (double)data[(size_t)y * step.buf[0] + x];
                if ((rtype & 7) == CV_32F) out.at<float>(y, x) = (float)v; else out.data[(size_t)y * out.step.buf[0] + x] = (uchar)v;
            }
        dst = out;
    }
    static MatZeros zeros(Size s, int t) { return MatZeros{s.height, s.width, t}; }

    Mat clone() const {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, type());
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step.buf[0], data + (size_t)y * step.buf[0], (size_t)cols * step.buf[1]);
        return m;
    }
    void copyTo(Mat& m) const {
        m.create(rows, cols, type());
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step.buf[0], data + (size_t)y * step.buf[0], (size_t)cols * step.buf[1]);
    }
    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.buf[0] + (size_t)r.x * step.buf[1];
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    // CV_32F matrix algebra of the back-end matcher variants (src/ORBmatcher.cc:274-278, 897-901, 1014-1024): OpenCV's
    // generic 32F GEMM / dot / norm accumulate in double and round the result to float once.
    Mat t() const {
        Mat m(cols, rows, type());
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.at<float>(x, y) = at<float>(y, x);
        return m;
    }
    double dot(const Mat& o) const {
        double r = 0;
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) r += (double)at<float>(y, x) * o.at<float>(y, x);
        return r;
    }

    int type() const { return flags & 0xfff; }
    int depth() const { return flags & 7; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return step.buf[1]; }
    size_t step1(int i = 0) const { return step.buf[i] / mock_elem_size(flags); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step.buf[0] == (size_t)cols * step.buf[1]; }
    Size size() const { return Size(cols, rows); }

    uchar* ptr(int y = 0) { return data + (size_t)y * step.buf[0]; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step.buf[0]; }
    template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.buf[0]); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.buf[0]); }
    template <class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step.buf[0]))[x]; }
    template <class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step.buf[0]))[x]; }
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step.buf[0]); }
    template <class T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step.buf[0]); }
};

static inline Mat operator*(const Mat& a, const Mat& b) {
    Mat m(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += (double)a.at<float>(i, k) * b.at<float>(k, j);
            m.at<float>(i, j) = (float)s;
        }
    return m;
}
static inline Mat mock_ew(const Mat& a, const Mat& b, int sign) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = sign > 0 ? a.at<float>(i, j) + b.at<float>(i, j) : a.at<float>(i, j) - b.at<float>(i, j);
    return m;
}
static inline Mat mock_scale(const Mat& a, double s) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) m.at<float>(i, j) = (float)(a.at<float>(i, j) * s);
    return m;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return mock_ew(a, b, 1); }
static inline Mat operator-(const Mat& a, const Mat& b) { return mock_ew(a, b, -1); }
static inline Mat operator-(const Mat& a) { return mock_scale(a, -1.0); }
static inline Mat operator*(const Mat& a, double s) { return mock_scale(a, s); }
static inline Mat operator*(double s, const Mat& a) { return mock_scale(a, s); }
static inline Mat operator/(const Mat& a, double s) { return mock_scale(a, 1.0 / s); }
static inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
static inline double norm(const Mat& a, const Mat& b, int type) {      // CV_32F, NORM_L1: sum |a - b| accumulated in double (OpenCV's normDiffL1_32f)
    assert(type == NORM_L1);
    double s = 0;
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) s += std::fabs((double)(a.at<float>(y, x) - b.at<float>(y, x)));
    return s;
}

template <class T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : (sizeof(T) == 8 ? CV_64F : CV_8U)) {}
    T& operator()(int y, int x) { return at<T>(y, x); }
    const T& operator()(int y, int x) const { return at<T>(y, x); }
};

// InputArray / OutputArray: thin proxies around a Mat (the only array kind the hot path passes)
class _InputArray {
public:
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
protected:
    Mat* m;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& mm) : _InputArray(mm) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void create(Size s, int t) const { m->create(s.height, s.width, t); }
    void release() const { if (m) m->release(); }
    Mat& getMatRef() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
static inline InputArray noArray() { static _InputArray a; return a; }

static inline float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

// cv::resize, INTER_LINEAR on CV_8UC1 with an explicit dsize (the only form the path uses, src/ORBextractor.cc:1139)
static inline void resize(InputArray src_, OutputArray dst_, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR) {
    Mat src = src_.getMat();
    if (interp != INTER_LINEAR || src.type() != CV_8UC1) std::abort();
    dst_.create(dsize.height, dsize.width, src.type());
    Mat dst = dst_.getMat();
    orc::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.buf[0], dst.data, dst.cols, dst.rows, (int)dst.step.buf[0]);
}

// cv::copyMakeBorder, BORDER_REFLECT_101 (+BORDER_ISOLATED: never read outside the ROI — this mock never does).
// Safe when src is the centre ROI of dst (src/ORBextractor.cc:1141): interior pixels map to themselves.
static inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType,
                                  const Scalar& = Scalar()) {
    Mat src = src_.getMat();
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || src.elemSize() != 1) std::abort();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    auto refl = [](int p, int n) {
        if (n == 1) return 0;
        while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
        return p;
    };
    const bool inplace = src.data == dst.data + (size_t)top * dst.step.buf[0] + left;
    for (int y = 0; y < dst.rows; y++) {
        int sy = refl(y - top, src.rows);
        const uchar* s = src.data + (size_t)sy * src.step.buf[0];
        uchar* d = dst.data + (size_t)y * dst.step.buf[0];
        for (int x = 0; x < dst.cols; x++) {
            bool interior = y >= top && y < top + src.rows && x >= left && x < left + src.cols;
            if (inplace && interior) continue;
            d[x] = s[refl(x - left, src.cols)];
        }
    }
}

// cv::GaussianBlur, 7x7 sigma 2 REFLECT_101 on CV_8UC1 (src/ORBextractor.cc:1010, 1083), in place allowed
static inline void GaussianBlur(InputArray src_, OutputArray dst_, Size k, double sx, double sy = 0, int borderType = BORDER_DEFAULT) {
    Mat src = src_.getMat();
    if (k.width != 7 || k.height != 7 || sx != 2.0 || (sy != 2.0 && sy != 0.0) || borderType != BORDER_REFLECT_101 ||
        src.type() != CV_8UC1)
        std::abort();
    Mat tmp = src.clone();
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    orc::gauss7_sigma2_u8(tmp.data, tmp.cols, tmp.rows, (int)tmp.step.buf[0], dst.data, (int)dst.step.buf[0]);
}

// cv::FAST (TYPE_9_16).  Key points as OpenCV fills them: pt local to the (sub)image, size 7, angle -1, response = score.
static inline void FAST(InputArray img_, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
    Mat img = img_.getMat();
    std::vector<orc::KeyPoint> out;
    orc::fast9_16(img.data, img.cols, img.rows, (int)img.step.buf[0], threshold, nms, out);
    kps.clear();
    kps.reserve(out.size());
    for (const orc::KeyPoint& k : out) kps.push_back(KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id));
}

struct KeyPointsFilter {
    // keep the n strongest (ties at the cut kept, as OpenCV) — only reached from ComputeKeyPointsOld (dead code upstream)
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n < 0 || (int)k.size() <= n) return;
        if (n == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n - 1, k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        float amb = k[n - 1].response;
        auto e = std::partition(k.begin() + n, k.end(), [amb](const KeyPoint& a) { return a.response >= amb; });
        k.resize(e - k.begin());
    }
};

// cv::FileStorage / cv::FileNode: the YAML persistence of DBoW2's TemplatedVocabulary (save / load are virtual, so they are
// instantiated with the class although the hot path only uses loadFromBinaryFile / loadFromTextFile).  Inert stand-ins: a
// storage that never opens, nodes that read as empty.
struct FileNode {
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

// highgui: the display branch of SparseImgAlign (display_ == false on this path)
#define CV_WINDOW_AUTOSIZE 1
static inline void namedWindow(const std::string&, int = 1) {}
static inline void imshow(const std::string&, const Mat&) {}
static inline int waitKey(int = 0) { return -1; }

}  // namespace cv
