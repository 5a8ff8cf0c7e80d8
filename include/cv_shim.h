// cv_shim.h -- the handful of cv::Mat members the facade touches, for build environments without the OpenCV C++ SDK
// (this image ships only the Python cv2 wheel). With real OpenCV available the facade includes <opencv2/core.hpp>
// instead and this header is not used.
#ifndef B2J_CV_SHIM_H_
#define B2J_CV_SHIM_H_
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8UC3
#define CV_8UC1 0
#define CV_8UC3 16
#endif

namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext, size_t ext_step = 0) : rows(r), cols(c), type_(type) {
        step = ext_step ? ext_step : (size_t)c * channels();
        data = (unsigned char *)ext;
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step = (size_t)c * channels();
        // like cv::Mat::create (fastMalloc): uninitialised memory -- the pages are first touched by whoever fills them
        store_ = std::shared_ptr<unsigned char>((unsigned char *)malloc(step * (size_t)r + 64), free);
        data = store_.get();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    size_t elemSize() const { return (size_t)channels(); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * channels(); }
    template <typename T = unsigned char> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T = unsigned char> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * channels()); return m; }
private:
    int type_ = CV_8UC3;
    std::shared_ptr<unsigned char> store_;
};
}  // namespace cv
#endif
