// cvmini.hpp -- the small part of cv:: that the stereo:: facade needs when the real
// OpenCV headers are not installed (they are not in this image).  With OpenCV present,
// stereo.h includes <opencv2/core/mat.hpp> instead and this file is unused.
// Written from scratch; only the members the ADCensus / EpipolarRectify path touches.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>

#ifndef CV_8U
#define CV_8U 0
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#endif

namespace cv {

typedef unsigned char uchar;

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

template <typename T, int N> struct Vec {
    T val[N];
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;

// Reference-counted 2-D matrix with row stride; ROI views share the buffer.
class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    // wrap user memory (not owned)
    Mat(int r, int c, int type, void* ptr, size_t stp = 0) : rows(r), cols(c), data((uchar*)ptr), type_(type)
    {
        step = stp ? stp : (size_t)c * elemSize();
    }

    static int depthBytes(int type)
    {
        switch (type & 7) {
            case CV_8U: case 1: return 1;
            case CV_16U: case CV_16S: return 2;
            case CV_32S: case CV_32F: return 4;
            default: return 8;
        }
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)depthBytes(type_) * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && type_ == type && owner_) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        const size_t bytes = step * (size_t)r;
        owner_ = std::shared_ptr<uchar>((uchar*)std::malloc(bytes ? bytes : 1), std::free);
        data = owner_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const
    {
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, type_);
        const size_t row = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, row);
    }
    Mat operator()(const Rect& r) const
    {
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw std::out_of_range("cv::Mat ROI outside the matrix");
        Mat m(*this);
        m.rows = r.height; m.cols = r.width;
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        return m;
    }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

private:
    int type_ = 0;
    std::shared_ptr<uchar> owner_;
};

inline void hconcat(const Mat& a, const Mat& b, Mat& dst)
{
    if (a.rows != b.rows || a.type() != b.type()) throw std::invalid_argument("cv::hconcat: size/type mismatch");
    Mat out(a.rows, a.cols + b.cols, a.type());
    const size_t ra = (size_t)a.cols * a.elemSize(), rb = (size_t)b.cols * b.elemSize();
    for (int y = 0; y < a.rows; ++y) {
        std::memcpy(out.data + (size_t)y * out.step, a.data + (size_t)y * a.step, ra);
        std::memcpy(out.data + (size_t)y * out.step + ra, b.data + (size_t)y * b.step, rb);
    }
    dst = out;
}

}  // namespace cv
