/*
 * oracle/ref_shim/cvshim.hpp -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Just enough of the cv:: namespace for the UNMODIFIED reference translation
 * unit /root/reference/source/ADCensus.cpp (and the headers it pulls in) to
 * compile with plain g++: OpenCV C++ headers are not installed in this image
 * and the reference does not vendor them.  Storage types are written from
 * scratch here; the five imgproc functions forward to oracle/cvport.c, whose
 * arithmetic is pinned against Python cv2 4.13.0.  HSI-only helpers
 * getGaussianKernel / filter2D / Mat::t / Mat*Mat cover exactly what computeGaussMedian (HSI path,
 * ADCensus.cpp:1475-1499) needs: a 3-tap fixed Gaussian, its outer product, and a float-kernel filter2D on
 * CV_8UC3 with BORDER_CONSTANT; pinned against cv2 4.13 in tests/test_cvport.py.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include "../cvport.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { INTER_LINEAR = 1 };

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Point {
    int x = 0, y = 0;
    Point() = default;
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3-ch only"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    bool operator==(const Vec& o) const { for (int i = 0; i < N; ++i) if (val[i] != o.val[i]) return false; return true; }
    bool operator!=(const Vec& o) const { return !(*this == o); }
};
typedef Vec<uchar, 3> Vec3b;

template <typename T> static inline T min(const T& a, const T& b) { return b < a ? b : a; }
template <typename T> static inline T max(const T& a, const T& b) { return a < b ? b : a; }

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() = default;
    Mat(Size sz, int type) { create(sz, type); }
    Mat(int r, int c, int type) { create(Size(c, r), type); }
    Mat(Size sz, int type, const Scalar& s) { create(sz, type); setTo(s); }

    static int depthBytes(int type) {
        switch (type & 7) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2;
                            case CV_32S: case CV_32F: return 4; default: return 8; }
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)depthBytes(type_) * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }

    void create(Size sz, int type) {
        if (data && rows == sz.height && cols == sz.width && type_ == type) return;
        rows = sz.height; cols = sz.width; type_ = type;
        step = (size_t)cols * elemSize();
        size_t bytes = step * (size_t)rows;
        buf_ = std::shared_ptr<uchar>((uchar*)std::malloc(bytes ? bytes : 1), std::free);
        data = buf_.get();
    }
    void create(int r, int c, int type) { create(Size(c, r), type); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        if (empty()) { dst = Mat(); return; }
        dst.create(size(), type_);
        std::memcpy(dst.data, data, step * (size_t)rows);
    }
    Mat& setTo(const Scalar& s) {
        const int cn = channels();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                for (int c = 0; c < cn; ++c) {
                    uchar* p = data + y * step + ((size_t)x * cn + c) * depthBytes(type_);
                    double v = s.val[c];
                    switch (type_ & 7) {
                        case CV_8U: *p = (uchar)v; break;
                        case CV_16S: *(short*)p = (short)v; break;
                        case CV_16U: *(ushort*)p = (ushort)v; break;
                        case CV_32S: *(int*)p = (int)v; break;
                        case CV_32F: *(float*)p = (float)v; break;
                        default: *(double*)p = v; break;
                    }
                }
        return *this;
    }
    static Mat zeros(Size sz, int type) { return Mat(sz, type, Scalar(0, 0, 0, 0)); }
    static Mat ones(Size sz, int type) { return Mat(sz, type, Scalar(1, 0, 0, 0)); }

    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T* begin() { return (T*)data; }
    template <typename T> T* end() { return (T*)(data + step * (size_t)rows); }
    template <typename T> const T* begin() const { return (const T*)data; }
    template <typename T> const T* end() const { return (const T*)(data + step * (size_t)rows); }

    Mat t() const {  // CV_32FC1 only (computeGaussMedian, ADCensus.cpp:1478)
        if (type_ != CV_32F) throw std::runtime_error("cvshim: Mat::t only for CV_32FC1");
        Mat m(Size(rows, cols), CV_32F);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) m.at<float>(x, y) = at<float>(y, x);
        return m;
    }

private:
    int type_ = 0;
    std::shared_ptr<uchar> buf_;
};
template <typename T> using MatIterator_ = T*;

inline Mat operator*(const Mat& a, const Mat& b) {  // CV_32FC1 matrix product, fp32 products and sums (cv::gemm, 4.13)
    if (a.type() != CV_32F || b.type() != CV_32F || a.cols != b.rows) throw std::runtime_error("cvshim: Mat*Mat only CV_32FC1");
    Mat m(Size(b.cols, a.rows), CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < b.cols; ++x) {
            float s = 0.f;
            for (int k = 0; k < a.cols; ++k) s += a.at<float>(y, k) * b.at<float>(k, x);
            m.at<float>(y, x) = s;
        }
    return m;
}
// cv::getGaussianKernel(ksize, sigma <= 0, CV_32F) for ksize 3: OpenCV's fixed small kernel {0.25, 0.5, 0.25}
inline Mat getGaussianKernel(int ksize, double sigma, int ktype = CV_64F) {
    if (ksize != 3 || sigma > 0 || ktype != CV_32F) throw std::runtime_error("cvshim: getGaussianKernel only (3, sigma <= 0, CV_32F)");
    Mat k(Size(1, 3), CV_32F);
    k.at<float>(0, 0) = 0.25f; k.at<float>(1, 0) = 0.5f; k.at<float>(2, 0) = 0.25f;
    return k;
}
// cv::filter2D(src CV_8UC3, dst, -1, float 3x3 kernel, anchor centre, delta 0, BORDER_CONSTANT): fp32 sum of
// kernel * pixel over the window (zeros outside), saturate_cast<uchar> = round half to even.  With the dyadic
// Gaussian weights every partial sum is exact, so the accumulation order cannot matter.
inline void filter2D(const Mat& src, Mat& dst, int, const Mat& kernel, Point = Point(-1, -1), double = 0, int border = BORDER_DEFAULT) {
    if (src.type() != CV_8UC3 || kernel.type() != CV_32F || kernel.rows != 3 || kernel.cols != 3 || border != BORDER_CONSTANT)
        throw std::runtime_error("cvshim: filter2D only CV_8UC3 / 3x3 CV_32F / BORDER_CONSTANT");
    Mat out(src.size(), CV_8UC3);
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x)
            for (int c = 0; c < 3; ++c) {
                float s = 0.f;
                for (int i = -1; i <= 1; ++i)
                    for (int j = -1; j <= 1; ++j) {
                        const int yy = y + i, xx = x + j;
                        if (yy < 0 || yy >= src.rows || xx < 0 || xx >= src.cols) continue;
                        s += kernel.at<float>(i + 1, j + 1) * (float)src.ptr<uchar>(yy)[3 * xx + c];
                    }
                const long r = std::lrintf(s);
                out.ptr<uchar>(y)[3 * x + c] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r);
            }
    dst = out;
}
inline bool imwrite(const std::string&, const Mat&) { return false; }

inline void equalizeHist(const Mat& src, Mat& dst) {
    Mat out(src.size(), CV_8UC1);
    cvp_equalize_hist(src.data, out.data, src.rows, src.cols);
    dst = out;
}
inline void blur(const Mat& src, Mat& dst, Size ksize) {
    if (ksize.width != 3 || ksize.height != 3) throw std::runtime_error("cvshim: blur only 3x3");
    Mat out(src.size(), CV_8UC1);
    cvp_blur3x3(src.data, out.data, src.rows, src.cols);
    dst = out;
}
inline void Canny(const Mat& src, Mat& dst, double t1, double t2, int aperture = 3, bool L2 = false) {
    if (aperture != 3 || L2) throw std::runtime_error("cvshim: Canny only aperture 3 / L1");
    Mat out(src.size(), CV_8UC1);
    cvp_canny3(src.data, out.data, src.rows, src.cols, (int)std::floor(t1), (int)std::floor(t2));
    dst = out;
}
inline void medianBlur(const Mat& src, Mat& dst, int ksize) {
    if (ksize != 3 || src.type() != CV_32FC1) throw std::runtime_error("cvshim: medianBlur only 3x3 CV_32F");
    Mat out(src.size(), CV_32FC1);
    cvp_median3x3_f32((const float*)src.data, (float*)out.data, src.rows, src.cols);
    dst = out;
}

}  // namespace cv
