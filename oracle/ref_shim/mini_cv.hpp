// ORACLE — TEST INFRASTRUCTURE ONLY (see ora_primitives.hpp). Nothing under mcvslam_b200/ may include this.
//
// mini_cv.hpp — the slice of the OpenCV 4 C++ API that the reference's hot-path translation units use, so that
// /root/reference's OWN sources (ORBextractor.cc, ORBExtractor.cpp, src/Matcher.cpp, src/Frame.cpp, src/Object.cpp,
// src/MapPoint.cpp, modules/camera/Pinhole.cpp, ...) compile UNMODIFIED into oracle/_ref/ (recipe: oracle/build_ref.py).
// OpenCV C++ is not installed in this image, so:
//   * the value types (Point_, Size_, Rect_, KeyPoint, DMatch, Mat with refcounted ROI views, Mat_, MatExpr for A*B+C,
//     InputArray / OutputArray) are re-implemented here with OpenCV's semantics;
//   * the image-processing primitives (cv::resize INTER_LINEAR u8, cv::GaussianBlur 7x7, cv::FAST 9/16 + NMS,
//     cv::fastAtan2, cvRound, cv::copyMakeBorder, cv::BFMatcher(NORM_HAMMING), cv::calcOpticalFlowPyrLK) are routed to
//     the cv2-4.13-pinned models in ora_primitives.hpp / ora_lk.hpp (tests/test_oracle_golden.py, tests/golden/).
// What this buys: every line of FIRST-PARTY logic on the path (cell loop, quadtree + std::priority_queue, IC_Angle, steered
// rBRIEF, streaming top-2, the filters, ComputeStereoMatch, the 30x30 grid, ProjectBunchMapPoints, the Frame constructor
// with its ThreadPool(3)) runs from the reference's own source text; tests/test_ref_parity.py checks oracle == _ref.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../ora_lk.hpp"
#include "../ora_primitives.hpp"

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(#expr); } while (0)

static inline int cvRound(double v) { return ora::cv_round(v); }
static inline int cvRound(float v) { return ora::cv_round(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {
typedef std::string String;

class Exception : public std::runtime_error {
   public:
    explicit Exception(const std::string& s) : std::runtime_error(s) {}
};

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4,
       BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };

// ---- small value types (modules/core/include/opencv2/core/types.hpp) ---------------------------------------------
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename T2> operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
};
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline std::ostream& operator<<(std::ostream& o, const Point_<T>& p) { return o << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size2i;
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

struct TermCriteria {
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount; double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int m, double e) : type(t), maxCount(m), epsilon(e) {}
};

class KeyPoint {
   public:
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class DMatch {
   public:
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");

// ---- Mat ---------------------------------------------------------------------------------------------------------
static inline size_t cv_elem_size1(int type) { static const size_t s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return s[CV_MAT_DEPTH(type)]; }
static inline size_t cv_elem_size(int type) { return cv_elem_size1(type) * CV_MAT_CN(type); }

class Mat;
class MatExpr;
template <typename T> class Mat_;
class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

class Mat {
   public:
    int flags = 0;  // = type
    int dims = 2;
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    const uchar* datastart = nullptr;  // whole allocation (locateROI)
    int whole_rows = 0, whole_cols = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar& s) { create(sz.height, sz.width, type); setTo(s); }
    // header over caller memory
    Mat(int r, int c, int type, void* d, size_t _step = 0)
        : flags(type), rows(r), cols(c), step(_step ? _step : (size_t)c * cv_elem_size(type)), data((uchar*)d), datastart((uchar*)d), whole_rows(r), whole_cols(c) {}
    Mat(const Mat& m, const Rect& roi) : Mat(m) {
        data += (size_t)roi.y * step + (size_t)roi.x * elemSize();
        rows = roi.height; cols = roi.width;
    }
    // column vector header over a std::vector (cv::Mat(const std::vector<T>&, copyData=false))
    template <typename T> explicit Mat(const std::vector<T>& v, bool copy = false);
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags) return;
        flags = type; rows = r; cols = c; step = (size_t)c * cv_elem_size(type);
        size_t bytes = step * (size_t)r;
        own_.reset(new uchar[bytes + 16], std::default_delete<uchar[]>());
        data = own_.get(); datastart = data; whole_rows = r; whole_cols = c;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { *this = Mat(); }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return cv_elem_size(flags); }
    size_t elemSize1() const { return cv_elem_size1(flags); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    bool isSubmatrix() const { return rows != whole_rows || cols != whole_cols; }
    void locateROI(Size& whole, Point& ofs) const {
        size_t delta = (size_t)(data - datastart);
        ofs.y = step ? (int)(delta / step) : 0;
        ofs.x = (int)((delta - (size_t)ofs.y * step) / elemSize());
        whole = Size(whole_cols, whole_rows);
    }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat operator()(Range rr, Range cr) const {
        Mat m(*this);
        if (rr.start != INT_MIN) m = m.rowRange(rr.start, rr.end);
        if (cr.start != INT_MIN) m = m.colRange(cr.start, cr.end);
        return m;
    }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data += (size_t)a * step; m.rows = b - a; return m; }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(int a, int b) const { Mat m(*this); m.data += (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }

    template <typename T = uchar> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <typename T = uchar> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
    // single index: element i of a 1 x N or N x 1 matrix (Mat::at(int i0))
    template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : (cols == 1 ? ptr<T>(i)[0] : ptr<T>(i / cols)[i % cols]); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : (cols == 1 ? ptr<T>(i)[0] : ptr<T>(i / cols)[i % cols]); }

    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, rb);
        return m;
    }
    void copyTo(OutputArray dst) const;
    void convertTo(OutputArray dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s);
    Mat reshape(int cn, int new_rows = 0) const {
        // only what the reference uses: continuous single-channel data, cn 0/1, new row count
        assert((cn == 0 || cn == channels()) && isContinuous());
        Mat m(*this);
        if (new_rows > 0 && new_rows != rows) {
            size_t tot = total();
            m.rows = new_rows; m.cols = (int)(tot / new_rows); m.step = (size_t)m.cols * elemSize();
            m.whole_rows = m.rows; m.whole_cols = m.cols; m.datastart = m.data;
        }
        return m;
    }
    Mat t() const;
    Mat inv() const;
    double dot(const Mat& m) const;
    Mat mul(const Mat& m, double scale = 1) const;

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); memset(m.data, 0, m.step * r); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat eye(int r, int c, int type) {
        Mat m = zeros(r, c, type);
        for (int i = 0; i < std::min(r, c); ++i) m.set_double(i, i, 1.0);
        return m;
    }
    double get_double(int y, int x) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(y, x);
            case CV_8S: return at<schar>(y, x);
            case CV_16U: return at<ushort>(y, x);
            case CV_16S: return at<short>(y, x);
            case CV_32S: return at<int>(y, x);
            case CV_32F: return at<float>(y, x);
            default: return at<double>(y, x);
        }
    }
    void set_double(int y, int x, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(y, x) = saturate_cast<uchar>(v); break;
            case CV_8S: at<schar>(y, x) = (schar)cvRound(v); break;
            case CV_16U: at<ushort>(y, x) = (ushort)cvRound(v); break;
            case CV_16S: at<short>(y, x) = saturate_cast<short>(v); break;
            case CV_32S: at<int>(y, x) = cvRound(v); break;
            case CV_32F: at<float>(y, x) = (float)v; break;
            default: at<double>(y, x) = v; break;
        }
    }

   private:
    std::shared_ptr<uchar> own_;
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<schar> { enum { type = CV_MAKETYPE(CV_8S, 1) }; };
template <> struct DataType<ushort> { enum { type = CV_MAKETYPE(CV_16U, 1) }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<Point2f> { enum { type = CV_32FC2 }; };

template <typename T> Mat::Mat(const std::vector<T>& v, bool copy)
    : flags(DataType<T>::type), rows((int)v.size()), cols(1), step(sizeof(T)), data((uchar*)v.data()), datastart((uchar*)v.data()),
      whole_rows((int)v.size()), whole_cols(1) {
    if (copy) *this = clone();
}

// Mat_<T> + comma initialiser: (cv::Mat_<float>(3, 3) << a, b, ...)
template <typename T> class MatCommaInitializer_;
template <typename T> class Mat_ : public Mat {
   public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) { assert(m.empty() || m.type() == DataType<T>::type); }
    T& operator()(int y, int x) { return this->template at<T>(y, x); }
    const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};
template <typename T> class MatCommaInitializer_ {
   public:
    MatCommaInitializer_(Mat_<T>* m) : m_(*m), i_(0) {}
    template <typename T2> MatCommaInitializer_<T>& operator,(T2 v) {
        m_.template at<T>(i_ / m_.cols, i_ % m_.cols) = T(v);
        ++i_;
        return *this;
    }
    operator Mat_<T>() const { return m_; }
    operator Mat() const { return m_; }
    Mat_<T> m_;
    int i_;
};
template <typename T, typename T2> static inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 val) {
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    return (ci, val);
}

// ---- gemm / arithmetic (modules/core/src/matmul.dispatch.cpp, arithm.cpp) -------------------------------------------
// d = alpha * a * b + beta * c. OpenCV's gemm has a dedicated small-matrix path (2 <= len <= 4 and len equal to one of
// the result's dimensions): float products summed LEFT TO RIGHT in float, then d = (float)(t * alpha + c * beta) with the
// scalars in double. Everything else goes through GEMMSingleMul<float, double>: double accumulation. (Pinned against
// cv2.gemm by the round-1 oracle work: DESIGN.md §5.)
static inline Mat gemm_impl(const Mat& a, const Mat& b, double alpha, const Mat* c, double beta) {
    assert(a.cols == b.rows && a.type() == b.type() && (a.type() == CV_32FC1 || a.type() == CV_64FC1));
    const int M = a.rows, N = b.cols, len = a.cols;
    Mat d(M, N, a.type());
    const bool small = 2 <= len && len <= 4 && (len == N || len == M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            if (a.type() == CV_32FC1) {
                double cv = c ? (double)c->at<float>(i, j) * beta : 0.0;
                if (small) {
                    float t = a.at<float>(i, 0) * b.at<float>(0, j);
                    for (int k = 1; k < len; ++k) t = t + a.at<float>(i, k) * b.at<float>(k, j);
                    d.at<float>(i, j) = (float)(t * alpha + cv);
                } else {
                    double t = 0;
                    for (int k = 0; k < len; ++k) t += (double)a.at<float>(i, k) * b.at<float>(k, j);
                    d.at<float>(i, j) = (float)(t * alpha + cv);
                }
            } else {
                double t = 0;
                for (int k = 0; k < len; ++k) t += a.at<double>(i, k) * b.at<double>(k, j);
                d.at<double>(i, j) = t * alpha + (c ? c->at<double>(i, j) * beta : 0.0);
            }
        }
    return d;
}

// Only the lazy form the reference relies on is modelled: (A * B) and (A * B) + C fold into ONE gemm call, like
// MatOp_GEMM. Every other operator is evaluated eagerly (identical values: negation and * 1.0 are exact).
class MatExpr {
   public:
    Mat a, b, c;
    double alpha = 1, beta = 0;
    bool is_gemm = false;
    MatExpr() {}
    MatExpr(const Mat& m) : a(m) {}
    Mat eval() const { return is_gemm ? gemm_impl(a, b, alpha, beta != 0 ? &c : nullptr, beta) : a; }
    operator Mat() const { return eval(); }
    template <typename T> operator Mat_<T>() const { return Mat_<T>(eval()); }
    Mat t() const { return eval().t(); }
    Mat inv() const { return eval().inv(); }
    template <typename T> T& at(int y, int x) { tmp_ = eval(); return tmp_.at<T>(y, x); }
    Mat tmp_;
};
inline Mat::Mat(const MatExpr& e) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { Mat m = e.eval(); *this = m; return *this; }

static inline MatExpr operator*(const Mat& a, const Mat& b) { MatExpr e; e.a = a; e.b = b; e.is_gemm = true; return e; }
static inline MatExpr operator*(const MatExpr& a, const Mat& b) { return a.eval() * b; }
static inline MatExpr operator*(const Mat& a, const MatExpr& b) { return a * b.eval(); }
static inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return a.eval() * b.eval(); }

template <typename F> static inline Mat cv_binary(const Mat& a, const Mat& b, F f) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat d(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols * a.channels(); ++x) {
            switch (a.depth()) {
                case CV_32F: d.ptr<float>(y)[x] = (float)f(a.ptr<float>(y)[x], b.ptr<float>(y)[x]); break;   // float op (arithm add32f)
                case CV_64F: d.ptr<double>(y)[x] = f(a.ptr<double>(y)[x], b.ptr<double>(y)[x]); break;
                case CV_16S: d.ptr<short>(y)[x] = saturate_cast<short>((double)f((int)a.ptr<short>(y)[x], (int)b.ptr<short>(y)[x])); break;
                case CV_32S: d.ptr<int>(y)[x] = f(a.ptr<int>(y)[x], b.ptr<int>(y)[x]); break;
                case CV_8U: d.ptr<uchar>(y)[x] = saturate_cast<uchar>((double)f((int)a.ptr<uchar>(y)[x], (int)b.ptr<uchar>(y)[x])); break;
                default: assert(!"depth"); break;
            }
        }
    return d;
}
struct cv_add_f { template <typename T> T operator()(T a, T b) const { return a + b; } };
struct cv_sub_f { template <typename T> T operator()(T a, T b) const { return a - b; } };
static inline Mat operator+(const Mat& a, const Mat& b) { return cv_binary(a, b, cv_add_f()); }
static inline Mat operator-(const Mat& a, const Mat& b) { return cv_binary(a, b, cv_sub_f()); }
static inline MatExpr operator+(const MatExpr& e, const Mat& m) {
    if (e.is_gemm && e.beta == 0) { MatExpr r = e; r.c = m; r.beta = 1; return r; }
    return MatExpr(e.eval() + m);
}
static inline MatExpr operator+(const Mat& m, const MatExpr& e) { return e + m; }
static inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return a + b.eval(); }
static inline Mat operator-(const MatExpr& e, const Mat& m) { return e.eval() - m; }
static inline Mat operator-(const Mat& m, const MatExpr& e) { return m - e.eval(); }
static inline Mat operator-(const MatExpr& a, const MatExpr& b) { return a.eval() - b.eval(); }

// Mat (op) scalar: result type = matrix type, computed in double (cv::subtract / convertScaleAbs family), rounded/saturated
static inline Mat cv_scale(const Mat& a, double s, double shift) {
    Mat d(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            if (a.depth() == CV_32F) d.at<float>(y, x) = (float)(a.at<float>(y, x) * s + shift);  // cvt32f: (float)(src*alpha + beta) in double? -> see note
            else d.set_double(y, x, a.get_double(y, x) * s + shift);
        }
    return d;
}
static inline Mat operator-(const Mat& a) { return cv_scale(a, -1, 0); }
static inline Mat operator-(const MatExpr& a) { return cv_scale(a.eval(), -1, 0); }
static inline Mat operator-(const Mat& a, const Scalar& s) { return cv_scale(a, 1, -s.val[0]); }
static inline Mat operator-(const Mat& a, double s) { return cv_scale(a, 1, -s); }
static inline Mat operator+(const Mat& a, double s) { return cv_scale(a, 1, s); }
static inline Mat operator*(const Mat& a, double s) { return cv_scale(a, s, 0); }
static inline Mat operator*(double s, const Mat& a) { return cv_scale(a, s, 0); }
static inline Mat operator*(const MatExpr& a, double s) { return cv_scale(a.eval(), s, 0); }
static inline Mat operator/(const Mat& a, double s) { return cv_scale(a, 1. / s, 0); }   // MatOp_AddEx with alpha = 1./s
static inline Mat operator/(const MatExpr& a, double s) { return cv_scale(a.eval(), 1. / s, 0); }

static inline Mat& operator+=(Mat& a, const Mat& b) { a = a + b; return a; }
static inline Mat& operator-=(Mat& a, const Mat& b) { a = a - b; return a; }
static inline Mat& operator*=(Mat& a, double s) { a = a * s; return a; }
static inline Mat& operator/=(Mat& a, double s) { a = a / s; return a; }

inline Mat Mat::t() const {
    Mat d(cols, rows, flags);
    const size_t es = elemSize();
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) memcpy(d.data + (size_t)x * d.step + (size_t)y * es, data + (size_t)y * step + (size_t)x * es, es);
    return d;
}
inline double Mat::dot(const Mat& m) const {
    // cv::Mat::dot for CV_32F: dotProd_<float> accumulates the float products in double
    assert(type() == m.type() && total() == m.total());
    double r = 0;
    Mat a = isContinuous() ? *this : clone(), b = m.isContinuous() ? m : m.clone();
    const size_t n = total();
    if (depth() == CV_32F) { const float* p = a.ptr<float>(); const float* q = b.ptr<float>(); for (size_t i = 0; i < n; ++i) r += (double)p[i] * q[i]; }
    else if (depth() == CV_64F) { const double* p = a.ptr<double>(); const double* q = b.ptr<double>(); for (size_t i = 0; i < n; ++i) r += p[i] * q[i]; }
    else assert(!"dot depth");
    return r;
}
inline Mat Mat::inv() const {
    // Gauss-Jordan in double; NOT bit-pinned against cv::invert (no hot-path caller uses it)
    assert(rows == cols);
    const int n = rows;
    std::vector<double> A((size_t)n * 2 * n, 0.0);
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) A[(size_t)i * 2 * n + j] = get_double(i, j); A[(size_t)i * 2 * n + n + i] = 1; }
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(A[(size_t)r * 2 * n + c]) > std::fabs(A[(size_t)p * 2 * n + c])) p = r;
        for (int j = 0; j < 2 * n; ++j) std::swap(A[(size_t)c * 2 * n + j], A[(size_t)p * 2 * n + j]);
        double d = A[(size_t)c * 2 * n + c];
        if (d == 0) return Mat::zeros(n, n, flags);
        for (int j = 0; j < 2 * n; ++j) A[(size_t)c * 2 * n + j] /= d;
        for (int r = 0; r < n; ++r) if (r != c) { double f = A[(size_t)r * 2 * n + c]; for (int j = 0; j < 2 * n; ++j) A[(size_t)r * 2 * n + j] -= f * A[(size_t)c * 2 * n + j]; }
    }
    Mat d(n, n, flags);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) d.set_double(i, j, A[(size_t)i * 2 * n + n + j]);
    return d;
}
inline Mat& Mat::setTo(const Scalar& s) {
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) set_double(y, x, s.val[0]);
    return *this;
}

static inline std::ostream& operator<<(std::ostream& o, const Mat& m) {
    o << "[";
    for (int y = 0; y < m.rows; ++y) { for (int x = 0; x < m.cols; ++x) o << (x ? ", " : "") << m.get_double(y, x); o << (y + 1 < m.rows ? ";\n " : ""); }
    return o << "]";
}

// ---- InputArray / OutputArray --------------------------------------------------------------------------------------
class _InputArray {
   public:
    enum Kind { NONE, MAT, VEC_MAT, VEC_POINT2F, VEC_UCHAR, VEC_FLOAT, VEC_KP };
    _InputArray() {}
    _InputArray(const Mat& m) : kind(MAT), mat(const_cast<Mat*>(&m)) {}
    _InputArray(const MatExpr& e) : kind(MAT), held(e.eval()) { mat = &held; }
    _InputArray(const std::vector<Mat>& v) : kind(VEC_MAT), obj((void*)&v) {}
    _InputArray(const std::vector<Point2f>& v) : kind(VEC_POINT2F), obj((void*)&v) {}
    _InputArray(const std::vector<uchar>& v) : kind(VEC_UCHAR), obj((void*)&v) {}
    _InputArray(const std::vector<float>& v) : kind(VEC_FLOAT), obj((void*)&v) {}
    bool empty() const {
        switch (kind) {
            case MAT: return mat->empty();
            case VEC_MAT: return ((std::vector<Mat>*)obj)->empty();
            case VEC_POINT2F: return ((std::vector<Point2f>*)obj)->empty();
            case VEC_UCHAR: return ((std::vector<uchar>*)obj)->empty();
            case VEC_FLOAT: return ((std::vector<float>*)obj)->empty();
            default: return true;
        }
    }
    Mat getMat(int = -1) const {
        switch (kind) {
            case MAT: return *mat;
            case VEC_POINT2F: { auto* v = (std::vector<Point2f>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_32FC2, v->data()); }
            case VEC_UCHAR: { auto* v = (std::vector<uchar>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_8UC1, v->data()); }
            case VEC_FLOAT: { auto* v = (std::vector<float>*)obj; return v->empty() ? Mat() : Mat((int)v->size(), 1, CV_32FC1, v->data()); }
            default: return Mat();
        }
    }
    Kind kind = NONE;
    Mat* mat = nullptr;
    void* obj = nullptr;
    Mat held;
};
class _OutputArray : public _InputArray {
   public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(const Mat& m) : _InputArray(m) {}   // OpenCV allows binding a temporary header (e.g. an ROI) for in-place writes
    _OutputArray(std::vector<Point2f>& v) : _InputArray(v) {}
    _OutputArray(std::vector<uchar>& v) : _InputArray(v) {}
    _OutputArray(std::vector<float>& v) : _InputArray(v) {}
    void create(int r, int c, int type) const {
        switch (kind) {
            case MAT: mat->create(r, c, type); break;
            case VEC_POINT2F: ((std::vector<Point2f>*)obj)->resize((size_t)r * c); break;
            case VEC_UCHAR: ((std::vector<uchar>*)obj)->resize((size_t)r * c); break;
            case VEC_FLOAT: ((std::vector<float>*)obj)->resize((size_t)r * c); break;
            default: assert(!"OutputArray::create"); break;
        }
    }
    void create(Size sz, int type) const { create(sz.height, sz.width, type); }
    void release() const { if (kind == MAT) mat->release(); }
    bool needed() const { return kind != NONE; }
};
static inline InputArray noArray() { static _OutputArray none; return none; }

inline void Mat::copyTo(OutputArray dst) const {
    if (empty()) { dst.release(); return; }
    dst.create(rows, cols, flags);
    Mat d = dst.getMat();
    if (d.data == data && d.step == step) return;
    const size_t rb = (size_t)cols * elemSize();
    for (int y = 0; y < rows; ++y) memmove(d.data + (size_t)y * d.step, data + (size_t)y * step, rb);
}
inline void Mat::convertTo(OutputArray dst, int rtype, double alpha, double beta) const {
    const int dtype = CV_MAKETYPE(CV_MAT_DEPTH(rtype < 0 ? flags : rtype), channels());
    Mat src = *this;                       // keeps the source alive when dst is this very header (in-place convert)
    Mat d(rows, cols, dtype);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) d.set_double(y, x, src.get_double(y, x) * alpha + beta);
    if (dst.kind == _InputArray::MAT) *dst.mat = d; else d.copyTo(dst);
}

// ---- core functions ------------------------------------------------------------------------------------------------
static inline float fastAtan2(float y, float x) { return ora::fast_atan2(y, x); }
static inline int borderInterpolate(int p, int len, int borderType) {
    borderType &= ~BORDER_ISOLATED;
    if ((unsigned)p < (unsigned)len) return p;
    if (borderType == BORDER_REPLICATE) return p < 0 ? 0 : len - 1;
    if (borderType == BORDER_REFLECT_101) return ora::reflect101(p, len);
    if (borderType == BORDER_REFLECT) { if (len == 1) return 0; do { if (p < 0) p = -p - 1; else p = len - 1 - (p - len); } while ((unsigned)p >= (unsigned)len); return p; }
    if (borderType == BORDER_WRAP) { if (p < 0) p -= ((p - len + 1) / len) * len; if (p >= len) p %= len; return p; }
    return -1;
}
// cv::copyMakeBorder (modules/core/src/copy.cpp): without BORDER_ISOLATED a submatrix source first takes what real pixels
// its parent has around it; dst may be the very buffer src is an ROI of (ORBextractor.cc:913-916) — rows are moved with
// memmove semantics and the border is written afterwards from the interior.
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar()) {
    Mat src = _src.getMat();
    if (src.isSubmatrix() && (borderType & BORDER_ISOLATED) == 0) {
        Size wholeSize; Point ofs;
        src.locateROI(wholeSize, ofs);
        int dtop = std::min(ofs.y, top), dbottom = std::min(wholeSize.height - src.rows - ofs.y, bottom);
        int dleft = std::min(ofs.x, left), dright = std::min(wholeSize.width - src.cols - ofs.x, right);
        src.data -= (size_t)dtop * src.step + (size_t)dleft * src.elemSize();
        src.rows += dtop + dbottom; src.cols += dleft + dright;
        top -= dtop; left -= dleft; bottom -= dbottom; right -= dright;
    }
    borderType &= ~BORDER_ISOLATED;
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    if (top == 0 && left == 0 && bottom == 0 && right == 0) { if (src.data != dst.data || src.step != dst.step) src.copyTo(dst); return; }
    assert(borderType != BORDER_CONSTANT || true);
    const size_t es = src.elemSize();
    // interior
    for (int y = 0; y < src.rows; ++y) {
        uchar* d = dst.data + (size_t)(y + top) * dst.step + (size_t)left * es;
        const uchar* s = src.data + (size_t)y * src.step;
        if (d != s) memmove(d, s, (size_t)src.cols * es);
    }
    // left / right of the interior rows
    for (int y = 0; y < src.rows; ++y) {
        uchar* row = dst.data + (size_t)(y + top) * dst.step;
        for (int x = 0; x < left; ++x) {
            if (borderType == BORDER_CONSTANT) { for (size_t b = 0; b < es; ++b) row[(size_t)x * es + b] = saturate_cast<uchar>(value.val[0]); continue; }
            int sx = borderInterpolate(x - left, src.cols, borderType);
            memcpy(row + (size_t)x * es, row + (size_t)(sx + left) * es, es);
        }
        for (int x = 0; x < right; ++x) {
            if (borderType == BORDER_CONSTANT) { for (size_t b = 0; b < es; ++b) row[(size_t)(left + src.cols + x) * es + b] = saturate_cast<uchar>(value.val[0]); continue; }
            int sx = borderInterpolate(src.cols + x, src.cols, borderType);
            memcpy(row + (size_t)(left + src.cols + x) * es, row + (size_t)(sx + left) * es, es);
        }
    }
    // top / bottom rows copy whole bordered rows
    const size_t rb = (size_t)dst.cols * es;
    for (int y = 0; y < top; ++y) {
        uchar* d = dst.data + (size_t)y * dst.step;
        if (borderType == BORDER_CONSTANT) { memset(d, saturate_cast<uchar>(value.val[0]), rb); continue; }
        int sy = borderInterpolate(y - top, src.rows, borderType);
        memcpy(d, dst.data + (size_t)(sy + top) * dst.step, rb);
    }
    for (int y = 0; y < bottom; ++y) {
        uchar* d = dst.data + (size_t)(top + src.rows + y) * dst.step;
        if (borderType == BORDER_CONSTANT) { memset(d, saturate_cast<uchar>(value.val[0]), rb); continue; }
        int sy = borderInterpolate(src.rows + y, src.rows, borderType);
        memcpy(d, dst.data + (size_t)(sy + top) * dst.step, rb);
    }
}

// cv::norm(a, b, NORM_L1) — src/Frame.cpp:267 on CV_16S patches: integer |a - b| sum, returned as double
static inline double norm(InputArray _a, InputArray _b, int normType = NORM_L2) {
    Mat a = _a.getMat(), b = _b.getMat();
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    double s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            double d = a.get_double(y, x) - b.get_double(y, x);
            if (normType == NORM_L1) s += std::fabs(d); else if (normType == NORM_INF) s = std::max(s, std::fabs(d)); else s += d * d;
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
// cv::norm(a) (NORM_L2) for CV_32F: normL2_32f accumulates float squares in double, sqrt in double
static inline double norm(InputArray _a, int normType = NORM_L2) {
    Mat a = _a.getMat();
    double s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            double d = a.get_double(y, x);
            if (normType == NORM_L1) s += std::fabs(d); else if (normType == NORM_INF) s = std::max(s, std::fabs(d)); else s += d * d;
        }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

// cv::SVD::compute — src/Map.cpp:375 (linear triangulation inside TrangularizationTwoObject, after the matching front-end).
// Not modelled: the parity scope ends at the match list; reaching it throws instead of returning numbers
// that no pin stands behind.
struct SVD {
    enum Flags { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    static void compute(InputArray, OutputArray, OutputArray, OutputArray, int = 0) { throw Exception("mini_cv: cv::SVD is not modelled"); }
};

// ---- parallel_for_ -------------------------------------------------------------------------------------------------
class ParallelLoopBody {
   public:
    virtual ~ParallelLoopBody() {}
    virtual void operator()(const Range& range) const = 0;
};
// MINI_CV_THREADS = worker count for cv::parallel_for_ (the reference's KnnMatch(vector<Mat>, ...) runs through it,
// src/Matcher.cpp:299). Default 1: inside the per-keypoint stereo loop a range is a single query.
static inline int& mini_cv_threads() { static int n = 1; return n; }
static inline void setNumThreads(int n) { mini_cv_threads() = n < 1 ? 1 : n; }
static inline void parallel_for_(const Range& r, const ParallelLoopBody& body, double = -1.) {
    const int n = r.end - r.start, T = std::min(mini_cv_threads(), std::max(1, n / 64));
    if (T <= 1) { if (n > 0) body(r); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
        int a = r.start + (int)((long long)n * t / T), b = r.start + (int)((long long)n * (t + 1) / T);
        th.emplace_back([&body, a, b] { body(Range(a, b)); });
    }
    for (auto& x : th) x.join();
}

// ---- imgproc -------------------------------------------------------------------------------------------------------
static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && "mini_cv: only the 8-bit bilinear resize of ORBextractor.cc:911");
    if (dsize.width == 0) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
    _dst.create(dsize.height, dsize.width, src.type());
    Mat dst = _dst.getMat();
    ora::Img s{src.data, src.cols, src.rows, src.step};
    ora::resize_linear_u8(s, dst.data, dst.cols, dst.rows, dst.step);
}
static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && (sigmaY == 2 || sigmaY == 0) &&
           (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && "mini_cv: only GaussianBlur(7x7, 2, 2, REFLECT_101) of ORBextractor.cc:875");
    Mat in = src.clone();  // in-place call
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    ora::Img s{in.data, in.cols, in.rows, in.step};
    ora::gauss7_u8(s, dst.data, dst.step);
}
enum { COLOR_BGR2GRAY = 6 };
static inline void cvtColor(InputArray _src, OutputArray _dst, int code) {
    Mat src = _src.getMat();
    assert(code == COLOR_BGR2GRAY && src.type() == CV_8UC3);
    _dst.create(src.rows, src.cols, CV_8UC1);
    Mat dst = _dst.getMat();
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            const uchar* p = src.ptr<uchar>(y) + 3 * x;
            dst.at<uchar>(y, x) = (uchar)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
        }
}

// ---- features2d ----------------------------------------------------------------------------------------------------
static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    Mat img = _img.getMat();
    assert(img.type() == CV_8UC1 && nonmaxSuppression);
    ora::Img s{img.data, img.cols, img.rows, img.step};
    std::vector<ora::FastKp> v;
    ora::fast9_16_nms(s, threshold, v);
    keypoints.clear();
    for (const auto& k : v) keypoints.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}
struct KeyPointsFilter {
    // only referenced by the dead ComputeKeyPointsOld (ORBextractor.cc:656-798)
    static void retainBest(std::vector<KeyPoint>& kps, int n) {
        if (n >= 0 && (int)kps.size() > n) {
            std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            kps.resize(n);
        }
    }
};
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> static inline Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

// cv::BFMatcher(NORM_HAMMING): knnMatch = per query the k best train rows ordered by (distance, trainIdx); imgIdx 0
// (pinned against cv2.BFMatcher on tie-heavy data: tests/golden g4).
class DescriptorMatcher {
   public:
    virtual ~DescriptorMatcher() {}
};
class BFMatcher : public DescriptorMatcher {
   public:
    explicit BFMatcher(int normType = NORM_L2, bool crossCheck = false) : norm_(normType) { (void)crossCheck; }
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false) { return std::make_shared<BFMatcher>(normType, crossCheck); }
    void knnMatch(InputArray _q, InputArray _t, std::vector<std::vector<DMatch>>& matches, int k, InputArray = noArray(), bool = false) const {
        assert(norm_ == NORM_HAMMING);
        Mat q = _q.getMat(), t = _t.getMat();
        matches.assign((size_t)q.rows, std::vector<DMatch>());
        if (q.rows == 0 || t.rows == 0) return;
        assert(q.cols == 32 && t.cols == 32 && q.type() == CV_8UC1);
        struct Body : ParallelLoopBody {
            const Mat &q, &t; std::vector<std::vector<DMatch>>& out; int k;
            Body(const Mat& q_, const Mat& t_, std::vector<std::vector<DMatch>>& o, int k_) : q(q_), t(t_), out(o), k(k_) {}
            void operator()(const Range& r) const override {
                std::vector<std::pair<int, int>> best;
                for (int i = r.start; i < r.end; ++i) {
                    best.clear();
                    const uchar* a = q.ptr<uchar>(i);
                    for (int j = 0; j < t.rows; ++j) {
                        std::pair<int, int> c((int)ora::hamming256(a, t.ptr<uchar>(j)), j);
                        if ((int)best.size() < k) { best.push_back(c); std::sort(best.begin(), best.end()); }
                        else if (c < best.back()) { best.back() = c; for (size_t p = best.size() - 1; p > 0 && best[p] < best[p - 1]; --p) std::swap(best[p], best[p - 1]); }
                    }
                    for (auto& b : best) out[i].push_back(DMatch(i, b.second, 0, (float)b.first));
                }
            }
        } body(q, t, matches, k);
        parallel_for_(Range(0, q.rows), body);
    }
    void match(InputArray q, InputArray t, std::vector<DMatch>& matches, InputArray = noArray()) const {
        std::vector<std::vector<DMatch>> knn;
        knnMatch(q, t, knn, 1);
        matches.clear();
        for (auto& v : knn) for (auto& m : v) matches.push_back(m);
    }
    int norm_;
};

static inline void drawMatches(InputArray, const std::vector<KeyPoint>&, InputArray, const std::vector<KeyPoint>&, const std::vector<DMatch>&, Mat&) {}
static inline void imshow(const std::string&, InputArray) {}
static inline int waitKey(int = 0) { return -1; }
namespace xfeatures2d {
static inline void matchGMS(const Size&, const Size&, const std::vector<KeyPoint>&, const std::vector<KeyPoint>&, const std::vector<DMatch>&,
                            std::vector<DMatch>&, bool = false, bool = false, double = 6.0) {
    throw Exception("mini_cv: matchGMS (opencv_contrib) is not modelled; Filter_GMS is unused by the reference's src/");
}
}  // namespace xfeatures2d

// ---- video ---------------------------------------------------------------------------------------------------------
// cv::calcOpticalFlowPyrLK with KL_Track's fixed arguments (src/Frame.cpp:52-54) -> ora_lk.hpp (pinned bit-exactly against cv2)
static inline void calcOpticalFlowPyrLK(InputArray _prev, InputArray _next, InputArray _pts, InputOutputArray _next_pts, OutputArray _status,
                                        OutputArray _err, Size winSize = Size(21, 21), int maxLevel = 3,
                                        TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                        double minEigThreshold = 1e-4) {
    Mat prev = _prev.getMat(), next = _next.getMat();
    assert(winSize.width == 10 && winSize.height == 10 && maxLevel == 1 && criteria.maxCount == 10 && criteria.epsilon == 0.01 && flags == 0 &&
           minEigThreshold == 0.001 && "mini_cv: only KL_Track's calcOpticalFlowPyrLK arguments");
    assert(prev.type() == CV_8UC1 && prev.rows == next.rows && prev.cols == next.cols && prev.step == next.step);
    const auto& pts = *(const std::vector<Point2f>*)_pts.obj;
    auto& npts = *(std::vector<Point2f>*)_next_pts.obj;
    auto& st = *(std::vector<uchar>*)_status.obj;
    auto& er = *(std::vector<float>*)_err.obj;
    const int n = (int)pts.size();
    npts.resize(n); st.resize(n); er.resize(n);
    if (n == 0) return;
    ora_lk::calc_lk(prev.data, next.data, prev.cols, prev.rows, (int)prev.step, (const float*)pts.data(), n, (float*)npts.data(), st.data(), er.data());
}

// ---- persistence (compile-only) ------------------------------------------------------------------------------------
// DBoW3's headers/bodies mention cv::FileStorage for its YAML vocabulary format (modules/DBow3/src/Vocabulary.cpp:839-857,
// 873-935,1144-1190). The tests load vocabularies through DBoW3's own BINARY stream format instead, so these only need to compile.
class FileNode {
   public:
    enum { NONE = 0, INT = 1, REAL = 2, STR = 3, SEQ = 4, MAP = 5 };
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    int type() const { return NONE; }
    size_t size() const { return 0; }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator float() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
   public:
    enum Mode { READ = 0, WRITE = 1, APPEND = 2 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { throw Exception("mini_cv: cv::FileStorage is not modelled (use the binary vocabulary format)"); }
    FileNode operator[](const std::string&) const { throw Exception("mini_cv: cv::FileStorage is not modelled (use the binary vocabulary format)"); }
};
template <typename T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
