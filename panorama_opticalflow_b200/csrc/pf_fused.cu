// pf_fused.cu -- shared-memory-tiled, fused stencil kernels of one pyramid level (CPU/PixFlow.hpp:272-340).
//
// Per level and direction the reference runs: 15x15 blur of the flow (:307), forward sweep, median5 (:325), backward
// sweep, median5 (:338), 15x15 blur + alpha blend (:339, :388-405).  Around the two sweep kernels (pf_sweep.cu) that is
// done here in four launches instead of eight, each reading its input tile once:
//   k_blur15<PREP>     flow -> [rows pass -> cols pass in shared memory] -> blurred  (+ the forward sweep's records)
//   k_median5<PREP>    flow -> 5x5 median from a shared-memory tile                  (+ the backward sweep's records)
//   k_median5<NONE>    plain median
//   k_blur15<DIFFUSE>  flow -> blur -> lowAlphaFlowDiffusion blend
// Arithmetic is unchanged (same taps, same order of the separately rounded fp32 operations as the two-pass versions):
// the row pass of a reflected row equals the row pass computed at that row, so building the halo with reflect-101 /
// replicate indices at load time reproduces OpenCV's border handling exactly.
//
// Tile staging.  The flow buffers have a row pitch that is a multiple of 16 bytes (flow_pitch), so a tile + halo whose
// footprint lies inside the image -- 91 % of the tiles at level 0 -- is ONE 2-D TMA tensor copy (cp.async.bulk.tensor,
// completion on an mbarrier, issued by one thread) into a dense shared-memory tile: no per-thread index arithmetic, no
// per-element load instructions.  Tiles that touch the image border keep the per-thread loads with reflect-101 / replicate
// indices (TMA's out-of-bounds fill is zero, which is neither).
#include <cuda.h>
#include <cstring>

#include "pf_kernels.cuh"
#include "pf_math.cuh"
#include "pf_prep.cuh"

namespace pf {

namespace {

constexpr int TILE = 32;                 // tile width
constexpr int BTH = 32;                  // blur tile height: 32 x 32 outputs per CTA, two per thread (rows ty and ty + 16)
constexpr int BTY = 16;                  // blur thread rows (512 threads)
constexpr int BR = 7;                    // radius of the 15-tap Gaussian
constexpr int BW = TILE + 2 * BR;        // 46
constexpr int BH = BTH + 2 * BR;         // 46
// A TMA box must start on a 16-byte boundary in global memory (an odd float2 column is an "illegal instruction" on the B200;
// tools/tma_box_probe.cu), so the blur's box starts one column further left, at x0 - 8, and is 48 columns wide: column c of the
// shared-memory tile holds image column x0 - 8 + c (c = 0 is never read).  The median's box starts at x0 - 2, which is even.
constexpr int BX0 = BR + 1;              // 8
constexpr int BWP = 48;                  // columns (and row pitch) of the blur tile in shared memory
constexpr int MTH = 8;                   // median tile height (256 threads)
constexpr int MR = 2;                    // radius of the 5x5 median
constexpr int MW = TILE + 2 * MR;        // 36
constexpr int MH = MTH + 2 * MR;         // 12

enum { MODE_PLAIN = 0, MODE_PREP = 1, MODE_DIFFUSE = 2 };

// BORDER_REFLECT_101: one branch-free reflection covers -n < p < 2n-1 (every halo position when n > 7); images narrower
// than the blur radius (half-resolution sides 4..7) need repeated reflections and take the looping form.
__device__ __forceinline__ int reflect1_clamped(int p, int n) {
    p = p < 0 ? -p : p;
    p = p >= n ? 2 * n - 2 - p : p;
    if ((unsigned)p >= (unsigned)n) p = reflect101(p, n);
    return p;
}

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One thread: 2-D TMA copy of the box at (x, y) of the tensor described by tm into dense shared memory, completion (byte count)
// on the mbarrier; every thread of the CTA then waits for phase 0 of it.  The barrier is used once per CTA.
__device__ __forceinline__ void tma_tile_load(const CUtensorMap* tm, void* dst, unsigned long long* bar, int x, int y, unsigned bytes,
                                              bool leader) {
    const unsigned b = smem_addr(bar);
    if (leader) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_addr(dst)), "l"(tm), "r"(x), "r"(y), "r"(b) : "memory");
    }
    __syncthreads();              // the initialised barrier is visible to the waiters
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
}

// 15x15 sigma 8 Gaussian of the 2-channel flow: row pass left-to-right over the 15 taps, column pass in the symmetric
// form, both from shared memory.  The two channels of a flow vector go through identical arithmetic, so
// every tap is ONE packed fp32x2 multiply and ONE packed add (pf_math.cuh) -- bit-identical to the scalar form at half
// the instructions; 46 x 46 input tile -> 46 x 32 row pass -> 32 x 32 outputs (halo overhead 1.44x instead of 1.9x).
// fp: row pitch (elements) of flow and out.
template <int MODE>
__global__ void __launch_bounds__(TILE * BTY)
k_blur15(const float2* __restrict__ flow, float2* __restrict__ out, int h, int w, int fp, PrepArgs pa,
         const __grid_constant__ CUtensorMap tm, int use_tma) {
    PF_GAUSS_TABLES
    __shared__ __align__(128) f2p s_in[BH][BWP];  // flow tile + halo (columns 1..46 used), reflect-101 at the image border
    __shared__ f2p s_row[BH][TILE];               // row pass
    __shared__ __align__(8) unsigned long long s_bar;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * BTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const f2p* fl = reinterpret_cast<const f2p*>(flow);
    if (use_tma && x0 >= BX0 && y0 >= BR && x0 - BX0 + BWP <= w && y0 - BR + BH <= h) {
        // interior tile: the whole 48 x 46 box lies inside the image -> one TMA tensor copy
        tma_tile_load(&tm, &s_in[0][0], &s_bar, x0 - BX0, y0 - BR, BH * BWP * (unsigned)sizeof(f2p), tx == 0 && ty == 0);
    } else {
        // border tile: each thread fetches (up to) 2 columns x 3 rows; reflect-101 indices computed once per thread
        const int gx0 = reflect1_clamped(x0 - BR + tx, w), gx1 = reflect1_clamped(x0 - BR + tx + TILE, w);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int ly = ty + r * BTY;
            if (ly < BH) {
                const f2p* row = fl + reflect1_clamped(y0 - BR + ly, h) * fp;
                s_in[ly][tx + 1] = row[gx0];
                if (tx < BW - TILE) s_in[ly][tx + 1 + TILE] = row[gx1];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int ly = ty + r * BTY;
        if (ly < BH) {
            f2p acc = pmuls(s_in[ly][tx + 1], kG15[7]);
#pragma unroll
            for (int i = 1; i < 15; ++i) acc = padd(acc, pmuls(s_in[ly][tx + 1 + i], kG15[i < 7 ? 7 - i : i - 7]));
            s_row[ly][tx] = acc;
        }
    }
    __syncthreads();
    const int x = x0 + tx;
    if (x >= w) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int ly = ty + r * BTY, y = y0 + ly;
        if (y >= h) break;
        f2p acc = pmuls(s_row[ly + BR][tx], kG15[0]);
#pragma unroll
        for (int i = 1; i <= 7; ++i) acc = padd(acc, pmuls(padd(s_row[ly + BR + i][tx], s_row[ly + BR - i][tx]), kG15[i]));
        const size_t p = (size_t)y * w + x;          // dense planes (alpha)
        const size_t q = (size_t)y * fp + x;         // pitched flow buffers
        const f2p fv = s_in[ly + BR][tx + BX0];
        if (MODE == MODE_DIFFUSE) {          // lowAlphaFlowDiffusion, CPU/PixFlow.hpp:395-404
            const float d = fsub(1.0f, fmul(pa.alpha0[p], pa.alpha1[p]));
            const float e = fsub(1.0f, d);
            out[q] = upk(padd(pmuls(acc, d), pmuls(fv, e)));
        } else {
            const float2 sv = upk(acc);
            out[q] = sv;
            if (MODE == MODE_PREP) emit_record(pa, make_err_ctx(pa.G1, w, h), x, y, w, h, upk(fv), sv);
        }
    }
    (void)kG5; (void)kG3O; (void)kG3H;
}

// medianBlur(32FC2, 5), replicate border, from a shared-memory tile (+ the records of the coming sweep)
template <int MODE>
__global__ void __launch_bounds__(TILE * MTH)
k_median5(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ blurred, int h, int w, int fp, PrepArgs pa,
          const __grid_constant__ CUtensorMap tm, int use_tma) {
    __shared__ __align__(128) float2 s_in[MH][MW];
    __shared__ __align__(8) unsigned long long s_bar;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * MTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (use_tma && x0 >= MR && y0 >= MR && x0 - MR + MW <= w && y0 - MR + MH <= h) {
        tma_tile_load(&tm, &s_in[0][0], &s_bar, x0 - MR, y0 - MR, MH * MW * (unsigned)sizeof(float2), tx == 0 && ty == 0);
    } else {
        // border tile: (up to) 2 columns x 2 rows per thread, replicate indices computed once per thread
        const int gx0 = clampi(x0 - MR + tx, 0, w - 1), gx1 = clampi(x0 - MR + tx + TILE, 0, w - 1);
        const int gy0 = clampi(y0 - MR + ty, 0, h - 1), gy1 = clampi(y0 - MR + ty + MTH, 0, h - 1);
        const float2* r0 = src + gy0 * fp;
        const float2* r1 = src + gy1 * fp;
        s_in[ty][tx] = r0[gx0];
        if (tx < MW - TILE) s_in[ty][tx + TILE] = r0[gx1];
        if (ty < MH - MTH) {
            s_in[ty + MTH][tx] = r1[gx0];
            if (tx < MW - TILE) s_in[ty + MTH][tx + TILE] = r1[gx1];
        }
        __syncthreads();
    }
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) return;
    float vx[25], vy[25];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const float2 v = s_in[ty + dy][tx + dx];
            vx[dy * 5 + dx] = v.x;
            vy[dy * 5 + dx] = v.y;
        }
#ifndef PF_MEDIAN_S3
#define PF_MEDIAN_S3 1
#endif
#if PF_MEDIAN_S3
    const float2 m = make_float2(median25_s3(vx), median25_s3(vy));
#else
    const float2 m = make_float2(median25(vx), median25(vy));
#endif
    const size_t q = (size_t)y * fp + x;
    dst[q] = m;
    if (MODE == MODE_PREP) emit_record(pa, make_err_ctx(pa.G1, w, h), x, y, w, h, m, blurred[q]);
}

inline dim3 blur_grid(int w, int h) { return dim3((w + TILE - 1) / TILE, (h + BTH - 1) / BTH); }
const dim3 kBlurBlock(TILE, BTY);
inline dim3 median_grid(int w, int h) { return dim3((w + TILE - 1) / TILE, (h + MTH - 1) / MTH); }

const CUtensorMap kNoMap = {};

// cuTensorMapEncodeTiled through the runtime's driver entry point table (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// Tensor map of a (h x w) float2 image with row pitch fp (elements) for boxes of bw x bh elements; false when the layout
// cannot be described (pitch not a multiple of 16 bytes, box larger than the image) -- the kernels then use per-thread loads.
bool make_flow_map(FlowTileMap* out, const float2* base, int h, int w, int fp, int bw, int bh) {
    out->valid = 0;
    static const bool disabled = getenv("PF_NO_TMA_TILES") != nullptr;       // diagnostics: per-thread loads everywhere
    EncodeTiledFn enc = encode_fn();
    if (disabled || !enc || (fp % 2) != 0 || w < bw || h < bh || ((uintptr_t)base % 16) != 0) return false;
    static_assert(sizeof(out->opaque) == sizeof(CUtensorMap), "FlowTileMap must hold a CUtensorMap");
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)fp * sizeof(float2)};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    const CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<float2*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
    memcpy(out->opaque, &tm, sizeof(tm));
    out->valid = 1;
    return true;
}

inline const CUtensorMap& as_map(const FlowTileMap* m) {
    return (m && m->valid) ? *reinterpret_cast<const CUtensorMap*>(m->opaque) : kNoMap;
}
inline int has_map(const FlowTileMap* m) { return (m && m->valid) ? 1 : 0; }

}  // namespace

int flow_pitch(int w) { return (w + 1) & ~1; }
bool make_blur_tile_map(FlowTileMap* out, const float2* base, int h, int w, int fp) { return make_flow_map(out, base, h, w, fp, BWP, BH); }
bool make_median_tile_map(FlowTileMap* out, const float2* base, int h, int w, int fp) { return make_flow_map(out, base, h, w, fp, MW, MH); }

void launch_blur15(const float2* flow, float2* blurred, int h, int w, int fp, const FlowTileMap* tm, cudaStream_t st) {
    PrepArgs pa = {};
    k_blur15<MODE_PLAIN><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, blurred, h, w, fp, pa, as_map(tm), has_map(tm));
}

void launch_blur15_prep(const float2* flow, float2* blurred, int h, int w, int fp, const float* alpha0, const float* alpha1,
                        const float2* G0, const float2* G1, SweepRec* rec, int dir, const FlowTileMap* tm, cudaStream_t st) {
    const PrepArgs pa = make_prep_args(alpha0, alpha1, G0, G1, rec, w, dir);
    k_blur15<MODE_PREP><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, blurred, h, w, fp, pa, as_map(tm), has_map(tm));
}

void launch_blur15_diffuse(const float2* flow, float2* out, int h, int w, int fp, const float* alpha0, const float* alpha1,
                           const FlowTileMap* tm, cudaStream_t st) {
    PrepArgs pa = {};
    pa.alpha0 = alpha0; pa.alpha1 = alpha1;
    k_blur15<MODE_DIFFUSE><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, out, h, w, fp, pa, as_map(tm), has_map(tm));
}

void launch_median5(const float2* src, float2* dst, int h, int w, int fp, const FlowTileMap* tm, cudaStream_t st) {
    PrepArgs pa = {};
    k_median5<MODE_PLAIN><<<median_grid(w, h), dim3(TILE, MTH), 0, st>>>(src, dst, nullptr, h, w, fp, pa, as_map(tm), has_map(tm));
}

void launch_median5_prep(const float2* src, float2* dst, const float2* blurred, int h, int w, int fp, const float* alpha0,
                         const float* alpha1, const float2* G0, const float2* G1, SweepRec* rec, int dir, const FlowTileMap* tm,
                         cudaStream_t st) {
    const PrepArgs pa = make_prep_args(alpha0, alpha1, G0, G1, rec, w, dir);
    k_median5<MODE_PREP><<<median_grid(w, h), dim3(TILE, MTH), 0, st>>>(src, dst, blurred, h, w, fp, pa, as_map(tm), has_map(tm));
}

}  // namespace pf
