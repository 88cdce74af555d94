// Minimal stand-in for <opencv2/core.hpp>, for TYPE-CHECKING the cv::Mat branch of include/pixflow_b200.hpp in an image that
// has no OpenCV C++ (tests/test_cpp_mirror.py).  It declares exactly the members of cv::Mat that the header and the test
// drivers touch, with OpenCV's signatures (opencv2/core/mat.hpp): rows, cols, data, step (a MatStep convertible to size_t),
// type(), empty(), clone(), ptr<T>(int), the (rows, cols, type) and (rows, cols, type, data, step) constructors, and the
// CV_8UC1 / CV_8UC4 / CV_32FC1 / CV_32FC2 type codes.  Not a product file; never installed.
#ifndef PIXFLOW_B200_TEST_STUB_OPENCV_CORE_HPP
#define PIXFLOW_B200_TEST_STUB_OPENCV_CORE_HPP
#include <cstddef>
#include <cstring>
#include <memory>

#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

namespace cv {

typedef unsigned char uchar;

struct MatStep {
    size_t p0 = 0;
    MatStep() {}
    MatStep(size_t s) : p0(s) {}
    operator size_t() const { return p0; }
    MatStep& operator=(size_t s) { p0 = s; return *this; }
};

class Mat {
public:
    enum { AUTO_STEP = 0 };
    int flags = 0, dims = 2, rows = 0, cols = 0;
    uchar* data = nullptr;
    MatStep step;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* d, size_t s = AUTO_STEP) : flags(type), rows(r), cols(c), data((uchar*)d) {
        step = s ? s : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        own_.reset(new uchar[(size_t)r * (size_t)step], std::default_delete<uchar[]>());
        data = own_.get();
    }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_32F ? 4 : 1); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }

private:
    std::shared_ptr<uchar> own_;
};

}  // namespace cv
#endif
