// Minimal layout-compatible stand-ins for the OpenCV value types that cross the extractor / Matcher API of the reference
// (cv::KeyPoint 28 B, cv::DMatch 16 B, a cv::Mat-like refcounted 8-bit matrix). OpenCV C++ is not installed in this image;
// a tree that has it compiles the host mirror with -DMCV_WITH_OPENCV and gets the real types instead (the C ABI only ever
// sees plain pointers, and the static_asserts in mcvslam_b200.hpp pin the layouts either way).
#pragma once
#ifdef MCV_WITH_OPENCV
#include <opencv2/core/mat.hpp>
#include <opencv2/core/types.hpp>
#else
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <memory>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

struct Point2f {
    float x = 0, y = 0;
    Point2f() = default;
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};

struct KeyPoint {  // modules/core/include/opencv2/core/types.hpp: pt, size, angle, response, octave, class_id
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() = default;
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 3.402823466e+38f;
    DMatch() = default;
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};

// 8-bit single-channel matrix with shared ownership; row()/rowRange-style headers alias the parent (like cv::Mat).
class Mat {
   public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    // non-owning header over caller memory (cv::Mat(rows, cols, type, data, step))
    Mat(int r, int c, int /*type*/, void* d, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (size_t)c), data((uint8_t*)d) {}

    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && own_ && step == (size_t)c) return;
        rows = r; cols = c; step = (size_t)c;
        own_.reset(new uint8_t[(size_t)r * c + 1], std::default_delete<uint8_t[]>());
        data = own_.get();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols || rows <= 1; }
    int type() const { return CV_8UC1; }
    Size size() const { return Size(cols, rows); }
    Mat row(int i) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)i * step; m.own_ = own_; return m; }
    Mat clone() const {
        Mat m(rows, cols, CV_8U);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
        return m;
    }
    template <class T = uint8_t> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T = uint8_t> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <class T = uint8_t> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <class T = uint8_t> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

   private:
    std::shared_ptr<uint8_t> own_;
};

}  // namespace cv
#endif
