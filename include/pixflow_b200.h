/*
 * pixflow_b200.h -- C-ABI of libpixflow_b200.so: the B200-native (sm_100a) replacement for the flow /
 * novel-view hot path of MungoMeng/Panorama-OpticalFlow.
 *
 * Every entry point names the reference interface it replaces (file:line under the reference tree).
 * Conventions: plain pointers and sizes, int return codes (0 = PF_OK), no exceptions cross the ABI,
 * caller-owned buffers.  Image pointers may be HOST or DEVICE pointers (detected with
 * cudaPointerGetAttributes); strides are in BYTES.  Images are 8-bit BGRA (CV_8UC4), flows are
 * interleaved (dx, dy) fp32 (CV_32FC2) in input-resolution pixels, blend is fp32 (CV_32FC1).
 * All calls are synchronous (results are complete on return) unless stated otherwise.
 * A CUDA device is mandatory: there is no CPU fallback -- engine creation fails with PF_ERR_NO_DEVICE.
 */
#ifndef PIXFLOW_B200_H
#define PIXFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_API __attribute__((visibility("default")))

enum pf_status {
    PF_OK = 0,
    PF_ERR_INVALID_ARGUMENT = 1,
    PF_ERR_UNKNOWN_ALGORITHM = 2, /* reference: throw VrCamException("unrecognized flow algorithm name"), CPU/PixFlow.hpp:499 */
    PF_ERR_CUDA = 3,
    PF_ERR_NO_DEVICE = 4
};

/* OpticalFlowInterface::DirectionHint, CPU/PixFlow.hpp:19 (same numeric values as the enum class) */
enum pf_direction_hint { PF_HINT_UNKNOWN = 0, PF_HINT_RIGHT = 1, PF_HINT_DOWN = 2, PF_HINT_LEFT = 3, PF_HINT_UP = 4 };

typedef struct pf_engine pf_engine;

/* Replaces makeOpticalFlowByName (CPU/PixFlow.hpp:459-500; GPU twin makeOpticalFlowByName_GPU,
 * GPU/PixFlow_GPU.hpp) plus the device probe of GPU/OpticalFlow.cpp:132-171.
 * flow_alg_name: "pixflow_low" (PixFlow<0>) or "pixflow_search_20" (PixFlow<20>); anything else ->
 * PF_ERR_UNKNOWN_ALGORITHM.  device < 0 selects the current CUDA device. */
PF_API int pf_engine_create(const char* flow_alg_name, int device, pf_engine** out_engine);
/* Replaces `delete flowAlg` (CPU/OpticalFlow.cpp:141). */
PF_API void pf_engine_destroy(pf_engine* engine);

/* Replaces OpticalFlowInterface::computeOpticalFlow(I0BGRA, I1BGRA, flow, hint), CPU/PixFlow.hpp:21-25 and
 * :72-135.  flow_out: rows x cols x (dx,dy) fp32. */
PF_API int pf_compute_flow(pf_engine* engine,
                           const void* i0_bgra, size_t i0_stride,
                           const void* i1_bgra, size_t i1_stride,
                           int rows, int cols, int hint,
                           void* flow_out, size_t flow_stride);

/* Replaces NovelViewGeneratorAsymmetricFlow::prepare(colorImageL, colorImageR), CPU/OpticalFlow.cpp:102-145:
 * circular pad by cols/20, flow(L,R,LEFT) and flow(R,L,RIGHT) (both directions run concurrently), crop.
 * Outputs the fields returned by getFlowLtoR()/getFlowRtoL() (CPU/OpticalFlow.hpp:67-68). */
PF_API int pf_prepare_bidirectional(pf_engine* engine,
                                    const void* image_l, size_t stride_l,
                                    const void* image_r, size_t stride_r,
                                    int rows, int cols,
                                    void* flow_l_to_r, size_t stride_lr,
                                    void* flow_r_to_l, size_t stride_rl);

/* Batched form of pf_prepare_bidirectional for n independent overlap pairs of identical size (the replica /
 * farm mode of SURVEY.md section 8e): all 2n flow computations are in flight concurrently on one device.
 * Each array holds n pointers; all pairs share rows, cols and the stride of their kind. */
PF_API int pf_prepare_bidirectional_batch(pf_engine* engine, int n,
                                          const void* const* images_l, size_t stride_l,
                                          const void* const* images_r, size_t stride_r,
                                          int rows, int cols,
                                          void* const* flows_l_to_r, size_t stride_lr,
                                          void* const* flows_r_to_l, size_t stride_rl);

/* Asynchronous form of the batch call ("optional async/stream + batch variants for replicas";
 * the reference itself is synchronous, CPU/OpticalFlow.cpp:128-141).  Enqueues the uploads (one H2D stream, pair order), the
 * 2n flow computations and the downloads (one D2H stream) and returns at once; the outputs are complete after
 * pf_wait(engine, slot).  slot is 0 or 1: the two slots own disjoint sets of workspaces, so a caller that alternates slots
 *     async(slot 0, batch k);  async(slot 1, batch k+1);  wait(slot 0);  async(slot 0, batch k+2);  wait(slot 1); ...
 * overlaps the download of batch k and the upload of batch k+2 with the compute of batch k+1.  Re-using a slot whose batch
 * has not been waited for waits for it first.  Input and output buffers must stay valid and untouched until pf_wait returns;
 * host buffers should be pinned (pf_host_alloc) or the copies degrade to synchronous ones.  The synchronous calls above use
 * slot 0 and wait for whatever is pending on it. */
PF_API int pf_prepare_bidirectional_batch_async(pf_engine* engine, int slot, int n,
                                                const void* const* images_l, size_t stride_l,
                                                const void* const* images_r, size_t stride_r,
                                                int rows, int cols,
                                                void* const* flows_l_to_r, size_t stride_lr,
                                                void* const* flows_r_to_l, size_t stride_rl);
PF_API int pf_wait(pf_engine* engine, int slot);

/* Replaces NovelViewUtil::combineNovelViews(imageL, imageR, flowLtoR, flowRtoL, blend),
 * CPU/OpticalFlow.cpp:30-92 (generateNovelViewPoint :9-28 inlined).  out_bgra: rows x cols BGRA8. */
PF_API int pf_combine_novel_views(pf_engine* engine,
                                  const void* image_l, size_t stride_l,
                                  const void* image_r, size_t stride_r,
                                  const void* flow_l_to_r, size_t stride_lr,
                                  const void* flow_r_to_l, size_t stride_rl,
                                  const void* blend, size_t stride_blend,
                                  int rows, int cols,
                                  void* out_bgra, size_t stride_out);

/* Fused prepare + setBlend + generateNovelView (CPU/main.cpp:82-89): the flows stay in HBM, only the merged
 * BGRA image is produced.  flow outputs may be NULL. */
PF_API int pf_novel_view(pf_engine* engine,
                         const void* image_l, size_t stride_l,
                         const void* image_r, size_t stride_r,
                         const void* blend, size_t stride_blend,
                         int rows, int cols,
                         void* out_bgra, size_t stride_out,
                         void* flow_l_to_r, size_t stride_lr,
                         void* flow_r_to_l, size_t stride_rl);

/* The stitching step around the flow path.
 *
 * pf_stitch_prepare replaces Stitchtools::prepare(colorImageL, colorImageR), CPU/StitchTool.cpp:7-36: MatchImages (:38-50),
 * the overlap masking (:16-33) and GenerateBlend (:98-146) with countblend (:148-191; replaces the reference's own
 * countblend_Kernel, GPU/StitchTool_GPU.cu:10-66, following the CPU path's sqrt(2) in double rather than that kernel's
 * literal 1.4142).  Outputs: getMap() (CV_8UC1: 100 = L only, 50 = R only, 150 = overlap), getOverlappedL()/getOverlappedR()
 * (CV_8UC4, the inputs of pf_prepare_bidirectional), the un-smoothed blend (CV_32FC1, GenerateBlend up to :131), MergedDis,
 * and getBlend() = the blend after the block-wise in-place cv::blur on ROIs (:133-142, order-dependent: every block sees the
 * blocks smoothed before it) and the final cv::blur (:143).  Any output pointer may be NULL.
 * Sizes the reference itself cannot run are rejected with PF_ERR_INVALID_ARGUMENT: shorter side < 200 (its search step
 * cols/200 would be 0 and countblend's loop would not terminate) and, when `blend` is requested, rows < 400 (cv::blur with
 * the empty kernel rows/400).  HOST or DEVICE pointers, strides in bytes. */
PF_API int pf_stitch_prepare(pf_engine* engine,
                             const void* image_l, size_t stride_l,
                             const void* image_r, size_t stride_r,
                             int rows, int cols,
                             void* map_u8, size_t stride_map,
                             void* overlapped_l, size_t stride_ol,
                             void* overlapped_r, size_t stride_or,
                             void* blend_raw, size_t stride_blend_raw,
                             void* merged_dis, size_t stride_dis,
                             void* blend, size_t stride_blend);

/* Replaces Stitchtools::setMergedmiddle + Stitchtools::Gather(), CPU/StitchTool.cpp:52-96: composes FinalResult (CV_8UC4)
 * from ImageL, ImageR, Mergedmiddle and Map.  The reference's unchecked map.at<uchar>(y +- i, x +- i) is reproduced as a
 * flat index into the continuous rows x cols map; indices outside the allocation (undefined behaviour in the reference)
 * match neither image. */
PF_API int pf_stitch_gather(pf_engine* engine,
                            const void* image_l, size_t stride_l,
                            const void* image_r, size_t stride_r,
                            const void* merged_middle, size_t stride_merged,
                            const void* map_u8, size_t stride_map,
                            int rows, int cols,
                            void* final_result, size_t stride_out);

/* One iteration of the reference driver's loop body, CPU/main.cpp:72-95, with every intermediate resident in HBM:
 * Stitchtools::prepare -> NovelViewGeneratorAsymmetricFlow::prepare(OverlappedL, OverlappedR) -> setBlend(getBlend()) ->
 * generateNovelView -> setMergedmiddle -> Gather -> getFinalResult().  final_result may be a DEVICE buffer that is passed
 * back as image_r of the next iteration (the reference's colorImageR = FinalResult).  blend / merged_middle / map_u8 are
 * optional outputs (NULL to skip). */
PF_API int pf_stitch_iteration(pf_engine* engine,
                               const void* image_l, size_t stride_l,
                               const void* image_r, size_t stride_r,
                               int rows, int cols,
                               void* final_result, size_t stride_out,
                               void* blend, size_t stride_blend,
                               void* merged_middle, size_t stride_merged,
                               void* map_u8, size_t stride_map);

/* Replaces the input preparation of the 4-input driver, CPU_4Input/main.cpp:64-79: a column of input k is blanked when
 * that input's alpha on the canvas' middle row (rows/2) is 0 in that column; colorImageL = image1 + image3 and
 * colorImageR = image2 + image4 with cv::Mat's saturating 8-bit addition.  images: 4 pointers (1.tif .. 4.tif order), one
 * common stride.  The two outputs feed pf_stitch_iteration (the rest of that driver is the same single pass). */
PF_API int pf_four_input_frontend(pf_engine* engine, const void* const images[4], size_t stride_in, int rows, int cols,
                                  void* image_l, size_t stride_l, void* image_r, size_t stride_r);

/* Pinned host memory for zero-staging transfers (optional; any host pointer is accepted by the calls above). */
PF_API int pf_host_alloc(void** ptr, size_t bytes);
PF_API int pf_host_free(void* ptr);

/* Number of kernel launches issued by this library so far in the process (bench.py's gpu_launches). */
PF_API uint64_t pf_kernel_launch_count(void);
/* Device time, in milliseconds, spent inside the wavefront sweep kernels during the last call on `engine`
 * (CUDA events on the launching streams; 0 if timing is disabled).  pf_set_sweep_timing(engine, 1) enables. */
PF_API int pf_set_sweep_timing(pf_engine* engine, int enabled);
PF_API double pf_last_sweep_ms(pf_engine* engine);
PF_API uint64_t pf_last_sweep_launches(pf_engine* engine);

/* Device-side stopwatch for benchmarks: CUDA events recorded on a stream of the engine.  start() stamps the
 * device clock now; stop() waits for every stream of the engine, stamps again and returns the elapsed ms. */
PF_API int pf_timer_start(pf_engine* engine);
PF_API int pf_timer_stop(pf_engine* engine, double* elapsed_ms);

/* Message of the last error on the calling thread ("" if none).  Replaces exception::what(). */
PF_API const char* pf_last_error(void);
PF_API const char* pf_version(void);

/* ---- diagnostic single-stage entry points (HOST pointers, contiguous arrays) --------------------------------
 * One kernel each, used by tests/ to localise a divergence to a stage.  Not a product API. */
/* the smoothing of GenerateBlend (CPU/StitchTool.cpp:133-145) alone, in place on a dense HOST blend (rows x cols fp32) */
PF_API int pf_stage_blend_smooth(pf_engine* engine, float* blend, const float* merged_dis, int rows, int cols);
PF_API int pf_stage_frontend(const void* bgra, int rows, int cols, int pad, float* grey, float* alpha, int dh, int dw);
PF_API int pf_stage_gauss5(const float* src, float* dst, int h, int w);
PF_API int pf_stage_pyr_down(const float* src, int sh, int sw, float* dst, int dh, int dw);
PF_API int pf_stage_gradient(const float* I, float* G_interleaved, int h, int w);
PF_API int pf_stage_blur15(const float* flow, float* dst, int h, int w, const float* alpha0, const float* alpha1);
PF_API int pf_stage_median5(const float* flow, float* dst, int h, int w);
PF_API int pf_stage_sweep(const float* alpha0, const float* alpha1, const float* G0, const float* G1,
                          const float* blurred, float* flow_inout, int h, int w, int dir);
PF_API int pf_stage_upsample_cubic(const float* src, int sh, int sw, float* dst, int dh, int dw);
PF_API int pf_stage_tail(const float* flow0, int sh, int sw, int rows, int pcols, int pad, int cols, float* out);
PF_API int pf_stage_initial_flow(const float* I0, const float* I1, const float* alpha0, const float* alpha1,
                                 float* flow, int h, int w, int hint, int dist);

/* Exhaustive on-device check of the branch-free exactly-rounded sqrt and division-by-constant used by the sweep
 * kernels, over every float in their validity range and every integer divisor in [wmin, wmax] (image widths).
 * out[0] = sqrt mismatches, out[1] = x/0.001f mismatches, out[2] = x/float(w) mismatches, out[3] = first bad w. */
PF_API int pf_selftest_exact_math(int wmin, int wmax, uint64_t* out4);

#ifdef __cplusplus
}
#endif
#endif /* PIXFLOW_B200_H */
