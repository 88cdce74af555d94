// pixflow_b200.hpp -- header-only C++ host side over the C-ABI (pixflow_b200.h), mirroring the reference's own
// operator interface for the flow / novel-view path so that its call sites compile unchanged:
//
//   optical_flow::OpticalFlowInterface, DirectionHint          CPU/PixFlow.hpp:15-26
//   optical_flow::makeOpticalFlowByName                        CPU/PixFlow.hpp:459-500
//   optical_flow::NovelViewUtil::combineNovelViews             CPU/OpticalFlow.hpp:19-32, CPU/OpticalFlow.cpp:30-92
//   optical_flow::NovelViewGenerator(AsymmetricFlow)           CPU/OpticalFlow.hpp:34-70, CPU/OpticalFlow.cpp:94-145
//   stitch_tools::Stitchtools                                  CPU/StitchTool.hpp:21-61, CPU/StitchTool.cpp:7-191
//
// Matrix type: when OpenCV's headers are available (they are not in the build image) cv::Mat is used directly
// and this header is a drop-in for the reference's PixFlow.hpp + OpticalFlow.hpp.  Otherwise a minimal ref-counted
// pf::Mat with the same few members (rows, cols, data, step, type(), clone(), empty()) stands in, so that host
// code and tests can be written -- and compiled here -- exactly like the reference's.
// Errors: the reference throws util::VrCamException; so does this header (same name, same what()).
#ifndef PIXFLOW_B200_HPP
#define PIXFLOW_B200_HPP

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>

#include "pixflow_b200.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(PIXFLOW_B200_NO_OPENCV)
#include <opencv2/core.hpp>
#define PIXFLOW_B200_HAVE_OPENCV 1
#endif
#endif

namespace util {
#ifndef PIXFLOW_B200_HAVE_VRCAM_EXCEPTION
#define PIXFLOW_B200_HAVE_VRCAM_EXCEPTION
// CPU/util.hpp:38-43
struct VrCamException : public std::exception {
    std::string msg;
    VrCamException() {}
    explicit VrCamException(const std::string& m) : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
#endif
}  // namespace util

namespace pf {

enum { PF_8UC1 = 0, PF_8UC4 = 24, PF_32FC1 = 5, PF_32FC2 = 13 };   // same numeric values as CV_8UC1 / CV_8UC4 / CV_32FC1 / CV_32FC2

#ifdef PIXFLOW_B200_HAVE_OPENCV
using Mat = cv::Mat;
inline Mat make_mat(int rows, int cols, int type) { return Mat(rows, cols, type); }
#else
// The subset of cv::Mat the flow path touches.
class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;   // bytes per row

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    // wraps user memory (no ownership), like cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), data((uint8_t*)d), type_(type) {
        step = s ? s : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        own_.reset(new uint8_t[(size_t)r * step], std::default_delete<uint8_t[]>());
        data = own_.get();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return type_ == PF_8UC1 ? 1 : (type_ == PF_32FC2 ? 8 : 4); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    template <class T> T* ptr(int y) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }

private:
    int type_ = 0;
    std::shared_ptr<uint8_t> own_;
};
inline Mat make_mat(int rows, int cols, int type) { return Mat(rows, cols, type); }
#endif

inline void check(int rc) {
    if (rc != PF_OK) throw util::VrCamException(pf_last_error());
}

}  // namespace pf

namespace optical_flow {

using pf::Mat;

// CPU/PixFlow.hpp:15-26
class OpticalFlowInterface {
public:
    virtual ~OpticalFlowInterface() {}
    enum class DirectionHint { UNKNOWN, RIGHT, DOWN, LEFT, UP };
    virtual void computeOpticalFlow(const Mat& I0BGRA, const Mat& I1BGRA, Mat& flow, DirectionHint hint) = 0;
};

// PixFlow<MaxPercentage> (CPU/PixFlow.hpp:28-457) on a B200; the twin of the reference's PixFlow_GPU
// (GPU/PixFlow_GPU.hpp:15-28) which this library replaces wholesale.
class PixFlowB200 : public OpticalFlowInterface {
public:
    explicit PixFlowB200(const std::string& flowAlgName, int device = -1) {
        // unknown name -> PF_ERR_UNKNOWN_ALGORITHM -> VrCamException("unrecognized flow algorithm name: ...")
        pf::check(pf_engine_create(flowAlgName.c_str(), device, &engine_));
    }
    ~PixFlowB200() override { pf_engine_destroy(engine_); }
    PixFlowB200(const PixFlowB200&) = delete;
    PixFlowB200& operator=(const PixFlowB200&) = delete;

    void computeOpticalFlow(const Mat& I0BGRA, const Mat& I1BGRA, Mat& flow, DirectionHint hint) override {
        flow = pf::make_mat(I0BGRA.rows, I0BGRA.cols, pf::PF_32FC2);    // `flow = Mat()` then reassigned, :112
        pf::check(pf_compute_flow(engine_, I0BGRA.data, I0BGRA.step, I1BGRA.data, I1BGRA.step, I0BGRA.rows, I0BGRA.cols,
                                  (int)hint, flow.data, flow.step));
    }
    pf_engine* handle() const { return engine_; }

    // Asynchronous batch of n independent pairs (pf_prepare_bidirectional_batch_async): returns at once, the flows are complete
    // after wait(slot).  slot is 0 or 1; alternate the slots to overlap the copies of one batch with the compute of the next.
    // The Mats must stay alive and untouched until wait(slot) returns.  flowsLtoR / flowsRtoL are allocated when empty.
    void prepareBidirectionalBatchAsync(int slot, const Mat* imagesL, const Mat* imagesR, Mat* flowsLtoR, Mat* flowsRtoL, int n) {
        if (n <= 0) throw util::VrCamException("prepareBidirectionalBatchAsync: n must be positive");
        std::unique_ptr<const void*[]> pl(new const void*[n]), pr(new const void*[n]);
        std::unique_ptr<void*[]> pa(new void*[n]), pb(new void*[n]);
        for (int i = 0; i < n; ++i) {
            if (imagesL[i].rows != imagesL[0].rows || imagesL[i].cols != imagesL[0].cols || (size_t)imagesL[i].step != (size_t)imagesL[0].step ||
                imagesR[i].rows != imagesL[0].rows || imagesR[i].cols != imagesL[0].cols || (size_t)imagesR[i].step != (size_t)imagesR[0].step)
                throw util::VrCamException("prepareBidirectionalBatchAsync: all pairs of a batch must share size and stride");
            if (flowsLtoR[i].empty()) flowsLtoR[i] = pf::make_mat(imagesL[0].rows, imagesL[0].cols, pf::PF_32FC2);
            if (flowsRtoL[i].empty()) flowsRtoL[i] = pf::make_mat(imagesL[0].rows, imagesL[0].cols, pf::PF_32FC2);
            pl[i] = imagesL[i].data; pr[i] = imagesR[i].data; pa[i] = flowsLtoR[i].data; pb[i] = flowsRtoL[i].data;
        }
        pf::check(pf_prepare_bidirectional_batch_async(engine_, slot, n, pl.get(), imagesL[0].step, pr.get(), imagesR[0].step,
                                                       imagesL[0].rows, imagesL[0].cols, pa.get(), flowsLtoR[0].step, pb.get(),
                                                       flowsRtoL[0].step));
    }
    void wait(int slot) { pf::check(pf_wait(engine_, slot)); }

private:
    pf_engine* engine_ = nullptr;
};

// CPU/PixFlow.hpp:459-500: caller owns the returned object (`delete flowAlg`, CPU/OpticalFlow.cpp:141)
static inline OpticalFlowInterface* makeOpticalFlowByName(const std::string flowAlgName) {
    return new PixFlowB200(flowAlgName);
}

// CPU/OpticalFlow.hpp:19-32
struct NovelViewUtil {
    static Mat combineNovelViews(const Mat& imageL, const Mat& imageR, const Mat& flowLtoR, const Mat& flowRtoL,
                                 const Mat& blend, pf_engine* engine = nullptr) {
        std::unique_ptr<PixFlowB200> own;
        if (!engine) { own.reset(new PixFlowB200("pixflow_low")); engine = own->handle(); }
        Mat out = pf::make_mat(imageL.rows, imageL.cols, pf::PF_8UC4);
        pf::check(pf_combine_novel_views(engine, imageL.data, imageL.step, imageR.data, imageR.step, flowLtoR.data, flowLtoR.step,
                                         flowRtoL.data, flowRtoL.step, blend.data, blend.step, imageL.rows, imageL.cols,
                                         out.data, out.step));
        return out;
    }
};

// CPU/OpticalFlow.hpp:34-48
class NovelViewGenerator {
public:
    virtual ~NovelViewGenerator() {}
    virtual void prepare(const Mat& colorImageL, const Mat& colorImageR) = 0;
    virtual void generateNovelView(Mat& outNovelViewMerged) = 0;
    virtual Mat getFlowLtoR() { return Mat(); }
    virtual Mat getFlowRtoL() { return Mat(); }
    virtual void setBlend(const Mat& blend) = 0;
};

// CPU/OpticalFlow.hpp:50-70, CPU/OpticalFlow.cpp:94-145
class NovelViewGeneratorAsymmetricFlow : public NovelViewGenerator {
public:
    std::string flowAlgName;
    Mat imageL, imageR;
    Mat flowLtoR, flowRtoL;
    Mat Blend;

    explicit NovelViewGeneratorAsymmetricFlow(const std::string flowAlgName_) : flowAlgName(flowAlgName_) {}
    ~NovelViewGeneratorAsymmetricFlow() override {}

    void prepare(const Mat& colorImageL, const Mat& colorImageR) override {
        imageL = colorImageL.clone();
        imageR = colorImageR.clone();
        alg_.reset(new PixFlowB200(flowAlgName));       // makeOpticalFlowByName(flowAlgName), CPU/OpticalFlow.cpp:128
        flowLtoR = pf::make_mat(imageL.rows, imageL.cols, pf::PF_32FC2);
        flowRtoL = pf::make_mat(imageL.rows, imageL.cols, pf::PF_32FC2);
        // circular pad by cols/20, flow(L,R,LEFT), flow(R,L,RIGHT), crop -- one call, both directions concurrent
        pf::check(pf_prepare_bidirectional(alg_->handle(), imageL.data, imageL.step, imageR.data, imageR.step, imageL.rows,
                                           imageL.cols, flowLtoR.data, flowLtoR.step, flowRtoL.data, flowRtoL.step));
    }
    void generateNovelView(Mat& outNovelViewMerged) override {
        outNovelViewMerged = NovelViewUtil::combineNovelViews(imageL, imageR, flowLtoR, flowRtoL, Blend, alg_ ? alg_->handle() : nullptr);
    }
    Mat getFlowLtoR() override { return flowLtoR; }
    Mat getFlowRtoL() override { return flowRtoL; }
    void setBlend(const Mat& blend) override { Blend = blend.clone(); }

private:
    std::unique_ptr<PixFlowB200> alg_;
};

}  // namespace optical_flow

namespace stitch_tools {

using pf::Mat;

// CPU/StitchTool.hpp:21-61.  Same public members and methods as the reference class; prepare() and Gather() run on the B200
// (pf_stitch_prepare / pf_stitch_gather).  MatchImages / GenerateBlend / countblend are not separately callable: prepare()
// is their only caller in the reference (CPU/StitchTool.cpp:15, :35) and runs all three in one device pass.
class Stitchtools {
public:
    Mat ImageL, ImageR;
    Mat Blend;
    Mat OverlappedL, OverlappedR;
    Mat Mergedmiddle;
    Mat Map;
    Mat FinalResult;
    Mat MergedDis;

    Stitchtools() {}
    ~Stitchtools() {}

    void prepare(const Mat& colorImageL, const Mat& colorImageR) {
        ImageL = colorImageL.clone();
        ImageR = colorImageR.clone();
        const int rows = ImageL.rows, cols = ImageL.cols;
        Map = pf::make_mat(rows, cols, pf::PF_8UC1);
        OverlappedL = pf::make_mat(rows, cols, pf::PF_8UC4);
        OverlappedR = pf::make_mat(rows, cols, pf::PF_8UC4);
        MergedDis = pf::make_mat(rows, cols, pf::PF_32FC1);
        Blend = pf::make_mat(rows, cols, pf::PF_32FC1);
        pf::check(pf_stitch_prepare(engine(), ImageL.data, ImageL.step, ImageR.data, ImageR.step, rows, cols, Map.data, Map.step,
                                    OverlappedL.data, OverlappedL.step, OverlappedR.data, OverlappedR.step, nullptr, 0,
                                    MergedDis.data, MergedDis.step, Blend.data, Blend.step));
    }

    void Gather() {
        FinalResult = pf::make_mat(ImageL.rows, ImageL.cols, pf::PF_8UC4);
        pf::check(pf_stitch_gather(engine(), ImageL.data, ImageL.step, ImageR.data, ImageR.step, Mergedmiddle.data, Mergedmiddle.step,
                                   Map.data, Map.step, ImageL.rows, ImageL.cols, FinalResult.data, FinalResult.step));
    }

    Mat getImageL() { return ImageL; }
    Mat getImageR() { return ImageR; }
    Mat getBlend() { return Blend; }
    Mat getMap() { return Map; }
    Mat getOverlappedL() { return OverlappedL; }
    Mat getOverlappedR() { return OverlappedR; }
    Mat getFinalResult() { return FinalResult; }

    void setMergedmiddle(const Mat& image) { Mergedmiddle = image.clone(); }

private:
    pf_engine* engine() {
        if (!alg_) alg_.reset(new optical_flow::PixFlowB200("pixflow_low"));
        return alg_->handle();
    }
    std::shared_ptr<optical_flow::PixFlowB200> alg_;
};

// The loop body of the reference driver (CPU/main.cpp:72-95) as one device-resident call; colorImageR / FinalResult may
// wrap device memory (Mat(rows, cols, type, devptr, step)) to keep the canvas on the GPU across the five iterations.
inline void stitchIteration(optical_flow::PixFlowB200& flowAlg, const Mat& colorImageL, const Mat& colorImageR, Mat& FinalResult) {
    if (FinalResult.empty()) FinalResult = pf::make_mat(colorImageL.rows, colorImageL.cols, pf::PF_8UC4);
    pf::check(pf_stitch_iteration(flowAlg.handle(), colorImageL.data, colorImageL.step, colorImageR.data, colorImageR.step,
                                  colorImageL.rows, colorImageL.cols, FinalResult.data, FinalResult.step, nullptr, 0, nullptr, 0,
                                  nullptr, 0));
}

}  // namespace stitch_tools

#endif  // PIXFLOW_B200_HPP
