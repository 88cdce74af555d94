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

// ---- Gauss-Seidel sweeps as an exact anti-diagonal wavefront (CPU/PixFlow.hpp:315-337), pf_sweep.cu ------
// Skewed (anti-diagonal-major) layout: element (x,y) at (x+y)*pitch + (posx ? x : y)
struct Skew { int w, h, pitch, posx; };
Skew make_skew(int w, int h);
size_t skew_elems(const Skew& s);
// row-major -> skewed copy of an interleaved gradient image
void launch_skew_copy_f2(const float2* src, float2* dst, const Skew& s, cudaStream_t st);
// per-pixel records for one sweep, in the wavefront-packed layout the sweep streams through shared memory:
// a = {E(f0), r0.x, r0.y, -} (own-flow terms; {-inf, f0} where alpha <= 0.9), b = {I0x, I0y, blurred.x, blurred.y}.
struct __align__(16) SweepRec { float4 a, b; };
constexpr int SWEEP_GROUP_ROWS = 16;      // logical rows per sweep warp (two lanes per row)
int sweep_nsteps_pad(int w);              // wavefront steps of a row group, rounded up to whole TMA stages
size_t sweep_rec_count(int h, int w);     // SweepRec elements a (h x w) level needs
void launch_sweep_prep(const float* alpha0, const float* alpha1, const float2* G0, const float2* G1,
                       const float2* blurred, const float2* flow, int fp, SweepRec* rec, int h, int w, int dir, cudaStream_t st);

// ---- fused, shared-memory-tiled stencils of one level (pf_fused.cu) -----------------------------------------------
// 15x15 sigma 8 blur of the 2-channel flow (CPU/PixFlow.hpp:307): plain, with the forward sweep's records fused in,
// or with lowAlphaFlowDiffusion (CPU/PixFlow.hpp:388-405) fused in
// The flow buffers (flow ping / pong, blurred flow) are row-major float2 with a row pitch of flow_pitch(w) elements -- a multiple
// of 16 bytes, which is what a TMA tensor map needs.  FlowTileMap: a CUtensorMap (opaque here, so that this header needs no
// <cuda.h>) describing one such buffer at one level for the tile + halo box of the blur (46 x 46) or of the median (36 x 12);
// valid == 0 (layout not describable, or maps unavailable) makes the kernels use per-thread loads for every tile.
int flow_pitch(int w);
struct alignas(64) FlowTileMap { unsigned char opaque[128]; int valid; };
bool make_blur_tile_map(FlowTileMap* out, const float2* base, int h, int w, int fp);
bool make_median_tile_map(FlowTileMap* out, const float2* base, int h, int w, int fp);
// fp: row pitch (elements) of the flow buffers involved; tm: tile map of the INPUT buffer (may be NULL)
void launch_blur15(const float2* flow, float2* blurred, int h, int w, int fp, const FlowTileMap* tm, cudaStream_t st);
void launch_blur15_prep(const float2* flow, float2* blurred, int h, int w, int fp, const float* alpha0, const float* alpha1,
                        const float2* G0, const float2* G1, SweepRec* rec, int dir, const FlowTileMap* tm, cudaStream_t st);
void launch_blur15_diffuse(const float2* flow, float2* out, int h, int w, int fp, const float* alpha0, const float* alpha1,
                           const FlowTileMap* tm, cudaStream_t st);
// medianBlur(32FC2, 5) (CPU/PixFlow.hpp:325, :338): plain, or with the backward sweep's records fused in
void launch_median5(const float2* src, float2* dst, int h, int w, int fp, const FlowTileMap* tm, cudaStream_t st);
void launch_median5_prep(const float2* src, float2* dst, const float2* blurred, int h, int w, int fp, const float* alpha0,
                         const float* alpha1, const float2* G0, const float2* G1, SweepRec* rec, int dir, const FlowTileMap* tm,
                         cudaStream_t st);
struct Sweep2Args {
    const SweepRec* rec;                      // wavefront-packed records from launch_sweep_prep (same dir)
    const float2* G1s;                        // skewed gradients of image 1
    long long g1s_last;                       // index of the last element of G1s (prefetch clamp)
    float2* flow;                             // row-major with row pitch fp, updated in place where alpha > 0.9
    int fp;                                   // row pitch of flow, in elements
    Skew s;
    uint4* boundary;      // LL lines {fx, flag, fy, flag}, zero-initialised (sweep2_boundary_lines of them)
    int* ticket;          // zero-initialised block ticket counter
    int cta_divisor;      // 1: as many CTAs as row blocks overlap in time (shortest sweep); k > 1: 1/k of that (see launch_sweep2)
};
size_t sweep2_boundary_lines(int h, int w);
void launch_sweep2(const Sweep2Args& a, int dir, cudaStream_t st);

// ---- inter-level upsample (CPU/PixFlow.hpp:123-124): INTER_CUBIC 32FC2 + "*= 1/0.9" --------------------
// sp / dp: row pitch (elements) of the source / destination flow buffer
void launch_upsample_cubic(const float2* src, int sh, int sw, int sp, float2* dst, int dh, int dw, int dp, cudaStream_t st);

// ---- tail (CPU/PixFlow.hpp:128-134): INTER_LINEAR to (rows x pcols), *2, 3x3 sigma 1 blur; only columns
// [pad, pad+cols) are written (the crop of CPU/OpticalFlow.cpp:143-144), out_stride in bytes -------------
void launch_tail(const float2* flow0, int sh, int sw, int sp, int rows, int pcols, int pad, int cols,
                 float2* out, size_t out_stride, cudaStream_t st);

// ---- coarsest-level search (CPU/PixFlow.hpp:190-270) ---------------------------------------------------
// ratio[0] <- computeIntensityRatio; flow <- zeros + adjustInitialFlow (hint 1..4, dist > 0) or zeros
void launch_initial_flow(const float* I0, const float* I1, const float* alpha0, const float* alpha1,
                         float2* flow, int fp, float* ratio, int h, int w, int hint, int dist, cudaStream_t st);

// ---- combineNovelViews (CPU/OpticalFlow.cpp:9-92); strides in bytes -------------------------------------
void launch_combine(const uint8_t* imageL, size_t strideL, const uint8_t* imageR, size_t strideR,
                    const float2* flowLR, size_t strideLR, const float2* flowRL, size_t strideRL,
                    const float* blend, size_t strideB, int rows, int cols,
                    uint8_t* out, size_t strideOut, cudaStream_t st);

// ---- Stitchtools::prepare without the blend smoothing (CPU/StitchTool.cpp:7-50, :98-131, :148-191), pf_stitch.cu ------
void launch_stitch_match_mask(const uint8_t* L, size_t strideL, const uint8_t* R, size_t strideR, int rows, int cols,
                              uint8_t* map, size_t strideM, uint8_t* oL, size_t strideOL, uint8_t* oR, size_t strideOR, cudaStream_t st);
void launch_stitch_blend_raw(const uint8_t* map, size_t strideM, int rows, int cols, float* blend, size_t strideB,
                             float* mdis, size_t strideD, cudaStream_t st);

// the smoothing of GenerateBlend (:133-145), in place on blend; scratch = stitch_smooth_scratch_bytes(rows, cols) device bytes.
// stitch_smooth_geometry: 0 ok, 1 the reference itself cannot run this size (rows < 400 or shorter side < 200), 2 unsupported.
int stitch_smooth_geometry(int rows, int cols, int* step, int* k1, int* k2, size_t* smem_bytes);
size_t stitch_smooth_scratch_bytes(int rows, int cols);
int launch_stitch_blend_smooth(float* blend, size_t strideB, const float* mdis, size_t strideD, int rows, int cols,
                               void* scratch, cudaStream_t st);
// ---- Stitchtools::Gather (CPU/StitchTool.cpp:52-96); gmap = rows*cols bytes of device scratch ---------------------------
void launch_stitch_gather(const uint8_t* L, size_t strideL, const uint8_t* R, size_t strideR, const uint8_t* merged, size_t strideG,
                          const uint8_t* map, size_t strideM, int rows, int cols, uint8_t* gmap, uint8_t* out, size_t strideO,
                          cudaStream_t st);

// ---- CPU_4Input front end (CPU_4Input/main.cpp:64-79): column pre-crop by the middle row's alpha + saturating adds -------
void launch_four_input(const uint8_t* const img[4], const size_t stride[4], int rows, int cols, uint8_t* outL, size_t strideL,
                       uint8_t* outR, size_t strideR, cudaStream_t st);

// ---- exhaustive self-test of the branch-free exact division / square root (pf_selftest.cu) ---------------
int selftest_exact_math(int wmin, int wmax, unsigned long long* out_mismatch_sqrt, unsigned long long* out_mismatch_eps,
                        unsigned long long* out_mismatch_w, int* out_first_bad_w);

}  // namespace pf
