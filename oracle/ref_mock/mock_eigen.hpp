// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// Stand-in for the slice of Eigen / Sophus that the reference's hot-path sources touch, so they compile unmodified
// into oracle/_ref/*.so (see mock_cv.hpp for the rationale).  Eigen and Sophus' Eigen dependency are not in this image
// (SURVEY.md §0.5).  Fixed-size dense matrices with the plain textbook evaluation order (row times column, left to
// right, float products summed in index order) — which is what Eigen's unrolled fixed-size 2x2 / 3x3 products compute
// without FMA contraction.  Only what the compiled sources need is provided.
#pragma once
#include <cmath>
#include <cstddef>
#include <algorithm>
#include <initializer_list>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// Third-party arithmetic the aligner needs and this image does not have (Eigen's LDLT, Sophus' quaternion SE3): forwarded to the
// oracle's restatements of the published algorithms (oracle/sparse_align.cpp), exactly as the image primitives forward to cvprims.cpp.
extern "C" {
void orc_ldlt_solve6(const float* A_rowmajor36, const float* b6, float* x6);
void orc_se3_exp(const float* x6, float* q, float* t);
void orc_se3_mul(const float* qa, const float* ta, const float* qb, const float* tb, float* q, float* t);
void orc_se3_inv(const float* q, const float* t, float* qo, float* to);
void orc_se3_apply(const float* q, const float* t, const float* p3, float* o3);
void orc_quat_to_matrix(const float* q, float* R9);
}

namespace Eigen {

enum { Dynamic = -1, ColMajor = 0, RowMajor = 1, NoChange = -2 };

template <class T, int R, int C, int Opt = 0> struct Matrix {
    T d[R * C];                                             // row-major storage (layout is not observable here)
    Matrix() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    Matrix(T a, T b, T c, T e) { static_assert(R * C == 4, "size"); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    template <class U> explicit Matrix(const Matrix<U, R, C>& o) { for (int i = 0; i < R * C; i++) d[i] = (T)o.d[i]; }

    T& operator()(int i, int j) { return d[i * C + j]; }
    const T& operator()(int i, int j) const { return d[i * C + j]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    T& x() { return d[0]; } const T& x() const { return d[0]; }
    T& y() { return d[1]; } const T& y() const { return d[1]; }
    T& z() { return d[2]; } const T& z() const { return d[2]; }
    int rows() const { return R; }
    int cols() const { return C; }
    T* data() { return d; }
    const T* data() const { return d; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
    void setZero() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
    void setIdentity() { *this = Identity(); }

    Matrix<T, C, R> transpose() const { Matrix<T, C, R> m; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) m(j, i) = (*this)(i, j); return m; }
    T dot(const Matrix& o) const { T s = d[0] * o.d[0]; for (int i = 1; i < R * C; i++) s += d[i] * o.d[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { T n = norm(); Matrix m(*this); for (int i = 0; i < R * C; i++) m.d[i] /= n; return m; }
    Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> m; for (int j = 0; j < C; j++) m.d[j] = (*this)(i, j); return m; }
    Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> m; for (int i = 0; i < R; i++) m.d[i] = (*this)(i, j); return m; }
    struct ColRef {
        Matrix& m; int j;
        ColRef& operator=(const Matrix<T, R, 1>& v) { for (int i = 0; i < R; i++) m(i, j) = v.d[i]; return *this; }
        operator Matrix<T, R, 1>() const { return const_cast<const Matrix&>(m).col(j); }
    };
    ColRef col(int j) { return ColRef{*this, j}; }
    struct CommaInit {
        Matrix& m; int n;
        CommaInit& operator,(T v) { m.d[n++] = v; return *this; }
    };
    CommaInit operator<<(T v) { d[0] = v; return CommaInit{*this, 1}; }
    template <class U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = (U)d[i]; return m; }
    T determinant() const {
        static_assert(R == C && (R == 2 || R == 3), "determinant: 2x2 / 3x3 only");
        const Matrix& a = *this;
        if (R == 2) return a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
        return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
               a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    }
    Matrix inverse() const {
        static_assert(R == C && (R == 2 || R == 3), "inverse: 2x2 / 3x3 only");
        const Matrix& a = *this;
        Matrix m;
        if (R == 2) {
            T id = T(1) / determinant();
            m(0, 0) = a(1, 1) * id; m(0, 1) = -a(0, 1) * id; m(1, 0) = -a(1, 0) * id; m(1, 1) = a(0, 0) * id;
            return m;
        }
        // Eigen's compute_inverse_size3: det = sum_k cofactor<k,0> * m(k,0), result(i,j) = cofactor<j,i> * (1/det)
        T c0 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1), c1 = a(2, 1) * a(0, 2) - a(2, 2) * a(0, 1), c2 = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
        T id = T(1) / ((c0 * a(0, 0) + c1 * a(1, 0)) + c2 * a(2, 0));
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                int i1 = (j + 1) % 3, i2 = (j + 2) % 3, j1 = (i + 1) % 3, j2 = (i + 2) % 3;
                m(i, j) = (a(i1, j1) * a(i2, j2) - a(i1, j2) * a(i2, j1)) * id;
            }
        return m;
    }
    // ---- the pieces src/SparseImageAlign.cc / include/NLSSolver_impl.hpp use
    Matrix& noalias() { return *this; }
    struct LDLTProxy {
        const Matrix& m;
        Matrix<T, R, 1> solve(const Matrix<T, R, 1>& b) const {
            static_assert(R == 6 && C == 6, "ldlt().solve: 6x6 only");
            Matrix<T, R, 1> x;
            orc_ldlt_solve6(m.d, b.d, x.d);                     // Eigen LDLT<Matrix6f, Lower>: pivoted, pseudo-inverse of D
            return x;
        }
    };
    LDLTProxy ldlt() const { return LDLTProxy{*this}; }
    Matrix<T, R, 1> diagonal() const { Matrix<T, R, 1> v; for (int i = 0; i < R; i++) v.d[i] = (*this)(i, i); return v; }
    Matrix<T, R * C, R * C> asDiagonal() const { Matrix<T, R * C, R * C> m; for (int i = 0; i < R * C; i++) m(i, i) = d[i]; return m; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; i++) d[i] += o.d[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; i++) d[i] -= o.d[i]; return *this; }
    Matrix& operator*=(T s) { for (int i = 0; i < R * C; i++) d[i] *= s; return *this; }
    Matrix& operator/=(T s) { for (int i = 0; i < R * C; i++) d[i] /= s; return *this; }
};

template <class T, int R, int C> Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] + b.d[i]; return m; }
template <class T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] - b.d[i]; return m; }
template <class T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = -a.d[i]; return m; }
template <class T, int R, int K, int C> Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
    Matrix<T, R, C> m;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) {
            T s = a(i, 0) * b(0, j);
            for (int k = 1; k < K; k++) s += a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}
template <class T, int R, int C, class S> Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S s) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] * (T)s; return m; }
template <class T, int R, int C, class S> Matrix<T, R, C> operator*(S s, const Matrix<T, R, C>& a) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = (T)s * a.d[i]; return m; }
template <class T, int R, int C, class S> Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S s) { Matrix<T, R, C> m; for (int i = 0; i < R * C; i++) m.d[i] = a.d[i] / (T)s; return m; }

// R x Dynamic (jacobian_cache_ of SparseImgAlign: Matrix<float, 6, Dynamic, ColMajor>)
template <class T, int R, int Opt> struct Matrix<T, R, Dynamic, Opt> {
    std::vector<T> d;            // column major
    int c = 0;
    void resize(int, int cols_) { c = cols_; d.assign((size_t)R * cols_, T(0)); }
    void setZero() { std::fill(d.begin(), d.end(), T(0)); }
    int rows() const { return R; }
    int cols() const { return c; }
    struct ColRef {
        Matrix& m; int j;
        ColRef& operator=(const Matrix<T, R, 1>& v) { for (int i = 0; i < R; i++) m.d[(size_t)j * R + i] = v.d[i]; return *this; }
        ColRef& operator=(const Matrix<T, 1, R>& v) { for (int i = 0; i < R; i++) m.d[(size_t)j * R + i] = v.d[i]; return *this; }   // row assigned to a column
        operator Matrix<T, R, 1>() const { Matrix<T, R, 1> v; for (int i = 0; i < R; i++) v.d[i] = m.d[(size_t)j * R + i]; return v; }
    };
    ColRef col(int j) { return ColRef{*this, j}; }
};

struct VectorXf {
    std::vector<float> v;
    VectorXf() {}
    template <int N> VectorXf(const Matrix<float, N, 1>& m) : v(m.d, m.d + N) {}
    int size() const { return (int)v.size(); }
    float operator[](int i) const { return v[i]; }
};

template <class T, int R, int C> std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
    for (int i = 0; i < R; i++) { for (int j = 0; j < C; j++) os << m(i, j) << (j + 1 < C ? " " : ""); if (i + 1 < R) os << "\n"; }
    return os;
}

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 6, 1> Vector6f;       // Thirdparty/sophus/sophus/se3.hpp:40 puts these into namespace Eigen
typedef Matrix<double, 6, 1> Vector6d;

template <class T> struct Quaternion {
    T w_, x_, y_, z_;
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
    T w() const { return w_; } T x() const { return x_; } T y() const { return y_; } T z() const { return z_; }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;

template <class T> struct aligned_allocator : std::allocator<T> {
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
};

}  // namespace Eigen

namespace Sophus {

// Rigid transforms stored as rotation matrix + translation: the hot-path sources only read rotationMatrix() /
// translation() and compose / invert / apply them (src/ORBmatcher.cc:1228-1236, 1356-1358, 1531-1532, 1582).
template <class T> struct SO3 {
    Eigen::Matrix<T, 3, 3> R;
    SO3() : R(Eigen::Matrix<T, 3, 3>::Identity()) {}
    explicit SO3(const Eigen::Matrix<T, 3, 3>& r) : R(r) {}
    const Eigen::Matrix<T, 3, 3>& matrix() const { return R; }
    SO3 inverse() const { return SO3(R.transpose()); }
    SO3 operator*(const SO3& o) const { return SO3(R * o.R); }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return R * p; }
};
#ifndef MOCK_SE3_QUATERNION
template <class T> struct SE3 {
    Eigen::Matrix<T, 3, 3> R;
    Eigen::Matrix<T, 3, 1> t;
    SE3() : R(Eigen::Matrix<T, 3, 3>::Identity()) {}
    SE3(const Eigen::Matrix<T, 3, 3>& r, const Eigen::Matrix<T, 3, 1>& tt) : R(r), t(tt) {}
    SE3(const SO3<T>& r, const Eigen::Matrix<T, 3, 1>& tt) : R(r.R), t(tt) {}
    const Eigen::Matrix<T, 3, 3>& rotationMatrix() const { return R; }
    const Eigen::Matrix<T, 3, 3>& rotation_matrix() const { return R; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t; }
    Eigen::Matrix<T, 3, 1>& translation() { return t; }
    SO3<T> so3() const { return SO3<T>(R); }
    SE3 inverse() const { Eigen::Matrix<T, 3, 3> Rt = R.transpose(); return SE3(Rt, Rt * (t * T(-1))); }
    SE3 operator*(const SE3& o) const { return SE3(R * o.R, t + R * o.t); }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return R * p + t; }
};
#else
// Quaternion form for the aligner (oracle/Makefile: -DMOCK_SE3_QUATERNION): unit quaternion (x, y, z, w) + translation with the
// arithmetic of Sophus' SE3Group / SO3Group (compose + renormalise, rotate by Eigen's _transformVector, exp with the small-angle
// branch), forwarded to the oracle's restatement of Thirdparty/sophus/sophus/{se3,so3}.hpp.
template <class T> struct SE3 {
    float q[4], t[3];
    SE3() { q[0] = q[1] = q[2] = 0.f; q[3] = 1.f; t[0] = t[1] = t[2] = 0.f; }
    static SE3 fromQT(const float* qq, const float* tt) { SE3 r; for (int i = 0; i < 4; i++) r.q[i] = qq[i]; for (int i = 0; i < 3; i++) r.t[i] = tt[i]; return r; }
    static SE3 exp(const Eigen::Matrix<T, 6, 1>& a) { SE3 r; orc_se3_exp(a.d, r.q, r.t); return r; }
    SE3 inverse() const { SE3 r; orc_se3_inv(q, t, r.q, r.t); return r; }
    SE3 operator*(const SE3& o) const { SE3 r; orc_se3_mul(q, t, o.q, o.t, r.q, r.t); return r; }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { Eigen::Matrix<T, 3, 1> o; orc_se3_apply(q, t, p.d, o.d); return o; }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { Eigen::Matrix<T, 3, 3> R; orc_quat_to_matrix(q, R.d); return R; }
    Eigen::Matrix<T, 3, 1> translation() const { return Eigen::Matrix<T, 3, 1>(t[0], t[1], t[2]); }
};
#endif
typedef SO3<float> SO3f;
typedef SE3<float> SE3f;
typedef SO3<double> SO3d;
typedef SE3<double> SE3d;

}  // namespace Sophus
