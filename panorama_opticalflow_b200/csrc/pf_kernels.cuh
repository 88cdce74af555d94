// pf_kernels.cuh -- launchers of the hand-written sm_100a kernels of the PixFlow hot path.
// Every launcher is asynchronous on the given stream; all pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pf {

// ---- front end (CPU/PixFlow.hpp:78-103) -----------------------------------------------------------
// Cubic 1/2 downscale of a BGRA8 image whose columns are read through the circular pad of
// NovelViewGeneratorAsymmetricFlow::prepare (CPU/OpticalFlow.cpp:113-126): padded column c maps to source
// column (c - pad) mod cols.  pad = 0 gives plain computeOpticalFlow.  Writes grey/255 and alpha/255.
void launch_frontend_resize(const uint8_t* bgra, size_t stride, int rows, int cols, int pad,
                            float* grey, float* alpha, int dh, int dw, cudaStream_t st);
// 5x5 sigma 0.25 pre-blur of the grey plane (reflect-101)
void launch_gauss5(const float* src, float* dst, int h, int w, cudaStream_t st);

// ---- pyramid (CPU/PixFlow.hpp:137-151): INTER_LINEAR resize of nplanes planes at once ----------------
struct PlaneSet { const float* src[4]; float* dst[4]; };
void launch_pyr_down(const PlaneSet& ps, int nplanes, int sh, int sw, int dh, int dw, cudaStream_t st);

// ---- gradients (CPU/PixFlow.hpp:284-294): Sobel k=1 (replicate) + 3x3 sigma 0.5 blur, interleaved out --
void launch_gradient(const float* I, float2* G, int h, int w, cudaStream_t st);

// ---- 15x15 sigma 8 blur of the 2-channel flow (CPU/PixFlow.hpp:307, :390) ----------------------------
void launch_blur15_rows(const float2* src, float2* tmp, int h, int w, cudaStream_t st);
// column pass; if alpha0 != nullptr fuses lowAlphaFlowDiffusion (CPU/PixFlow.hpp:395-404) with `flow`
void launch_blur15_cols(const float2* tmp, float2* dst, int h, int w,
                        const float* alpha0, const float* alpha1, const float2* flow, cudaStream_t st);

// ---- medianBlur(32FC2, 5) (CPU/PixFlow.hpp:325, :338) ------------------------------------------------
void launch_median5(const float2* src, float2* dst, int h, int w, cudaStream_t st);

// ---- Gauss-Seidel sweep as an exact anti-diagonal wavefront (CPU/PixFlow.hpp:315-337) ----------------
struct SweepArgs {
    const float* alpha0; const float* alpha1;
    const float2* G0; const float2* G1; const float2* blurred;
    float2* flow;
    int h, w;
    uint4* boundary;      // (nblocks-1) x w LL lines {fx, flag, fy, flag}, zero-initialised
    int* ticket;          // zero-initialised block ticket counter
};
size_t sweep_boundary_lines(int h, int w);   // number of uint4 lines a sweep of this size needs
void launch_sweep(const SweepArgs& a, int dir, cudaStream_t st);

// ---- inter-level upsample (CPU/PixFlow.hpp:123-124): INTER_CUBIC 32FC2 + "*= 1/0.9" --------------------
void launch_upsample_cubic(const float2* src, int sh, int sw, float2* dst, int dh, int dw, cudaStream_t st);

// ---- tail (CPU/PixFlow.hpp:128-134): INTER_LINEAR to (rows x pcols), *2, 3x3 sigma 1 blur; only columns
// [pad, pad+cols) are written (the crop of CPU/OpticalFlow.cpp:143-144), out_stride in bytes -------------
void launch_tail(const float2* flow0, int sh, int sw, int rows, int pcols, int pad, int cols,
                 float2* out, size_t out_stride, cudaStream_t st);

// ---- coarsest-level search (CPU/PixFlow.hpp:190-270) ---------------------------------------------------
// ratio[0] <- computeIntensityRatio; flow <- zeros + adjustInitialFlow (hint 1..4, dist > 0) or zeros
void launch_initial_flow(const float* I0, const float* I1, const float* alpha0, const float* alpha1,
                         float2* flow, float* ratio, int h, int w, int hint, int dist, cudaStream_t st);

// ---- combineNovelViews (CPU/OpticalFlow.cpp:9-92); strides in bytes -------------------------------------
void launch_combine(const uint8_t* imageL, size_t strideL, const uint8_t* imageR, size_t strideR,
                    const float2* flowLR, size_t strideLR, const float2* flowRL, size_t strideRL,
                    const float* blend, size_t strideB, int rows, int cols,
                    uint8_t* out, size_t strideOut, cudaStream_t st);

}  // namespace pf
