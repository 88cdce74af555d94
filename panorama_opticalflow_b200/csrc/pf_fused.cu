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

// 15x15 sigma 8 Gaussian of the 2-channel flow: row pass left-to-right over the 15 taps, column pass in the symmetric
// form, both from shared memory.  The two channels of a flow vector go through identical arithmetic, so
// every tap is ONE packed fp32x2 multiply and ONE packed add (pf_math.cuh) -- bit-identical to the scalar form at half
// the instructions; 46 x 46 input tile -> 46 x 32 row pass -> 32 x 32 outputs (halo overhead 1.44x instead of 1.9x).
template <int MODE>
__global__ void __launch_bounds__(TILE * BTY)
k_blur15(const float2* __restrict__ flow, float2* __restrict__ out, int h, int w, PrepArgs pa) {
    PF_GAUSS_TABLES
    __shared__ f2p s_in[BH][BW];             // flow tile + halo, reflect-101 at the image border
    __shared__ f2p s_row[BH][TILE];          // row pass
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * BTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const f2p* fl = reinterpret_cast<const f2p*>(flow);
    {   // tile + halo: each thread fetches (up to) 2 columns x 3 rows; reflect-101 indices computed once per thread.
        const int gx0 = reflect1_clamped(x0 - BR + tx, w), gx1 = reflect1_clamped(x0 - BR + tx + TILE, w);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int ly = ty + r * BTY;
            if (ly < BH) {
                const f2p* row = fl + reflect1_clamped(y0 - BR + ly, h) * w;
                s_in[ly][tx] = row[gx0];
                if (tx < BW - TILE) s_in[ly][tx + TILE] = row[gx1];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int ly = ty + r * BTY;
        if (ly < BH) {
            f2p acc = pmuls(s_in[ly][tx], kG15[7]);
#pragma unroll
            for (int i = 1; i < 15; ++i) acc = padd(acc, pmuls(s_in[ly][tx + i], kG15[i < 7 ? 7 - i : i - 7]));
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
        const size_t p = (size_t)y * w + x;
        const f2p fp = s_in[ly + BR][tx + BR];
        if (MODE == MODE_DIFFUSE) {          // lowAlphaFlowDiffusion, CPU/PixFlow.hpp:395-404
            const float d = fsub(1.0f, fmul(pa.alpha0[p], pa.alpha1[p]));
            const float e = fsub(1.0f, d);
            out[p] = upk(padd(pmuls(acc, d), pmuls(fp, e)));
        } else {
            const float2 sv = upk(acc);
            out[p] = sv;
            if (MODE == MODE_PREP) emit_record(pa, make_err_ctx(pa.G1, w, h), x, y, w, h, upk(fp), sv);
        }
    }
    (void)kG5; (void)kG3O; (void)kG3H;
}

// medianBlur(32FC2, 5), replicate border, from a shared-memory tile (+ the records of the coming sweep)
template <int MODE>
__global__ void __launch_bounds__(TILE * MTH)
k_median5(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ blurred, int h, int w, PrepArgs pa) {
    __shared__ float2 s_in[MH][MW + 1];
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * MTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    {   // tile + halo: (up to) 2 columns x 2 rows per thread, replicate indices computed once per thread
        const int gx0 = clampi(x0 - MR + tx, 0, w - 1), gx1 = clampi(x0 - MR + tx + TILE, 0, w - 1);
        const int gy0 = clampi(y0 - MR + ty, 0, h - 1), gy1 = clampi(y0 - MR + ty + MTH, 0, h - 1);
        const float2* r0 = src + gy0 * w;
        const float2* r1 = src + gy1 * w;
        s_in[ty][tx] = r0[gx0];
        if (tx < MW - TILE) s_in[ty][tx + TILE] = r0[gx1];
        if (ty < MH - MTH) {
            s_in[ty + MTH][tx] = r1[gx0];
            if (tx < MW - TILE) s_in[ty + MTH][tx + TILE] = r1[gx1];
        }
    }
    __syncthreads();
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
    const size_t p = (size_t)y * w + x;
    dst[p] = m;
    if (MODE == MODE_PREP) emit_record(pa, make_err_ctx(pa.G1, w, h), x, y, w, h, m, blurred[p]);
}

inline dim3 blur_grid(int w, int h) { return dim3((w + TILE - 1) / TILE, (h + BTH - 1) / BTH); }
const dim3 kBlurBlock(TILE, BTY);
inline dim3 median_grid(int w, int h) { return dim3((w + TILE - 1) / TILE, (h + MTH - 1) / MTH); }

}  // namespace

void launch_blur15(const float2* flow, float2* blurred, int h, int w, cudaStream_t st) {
    PrepArgs pa = {};
    k_blur15<MODE_PLAIN><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, blurred, h, w, pa);
}

void launch_blur15_prep(const float2* flow, float2* blurred, int h, int w, const float* alpha0, const float* alpha1,
                        const float2* G0, const float2* G1, SweepRec* rec, int dir, cudaStream_t st) {
    const PrepArgs pa = make_prep_args(alpha0, alpha1, G0, G1, rec, w, dir);
    k_blur15<MODE_PREP><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, blurred, h, w, pa);
}

void launch_blur15_diffuse(const float2* flow, float2* out, int h, int w, const float* alpha0, const float* alpha1, cudaStream_t st) {
    PrepArgs pa = {};
    pa.alpha0 = alpha0; pa.alpha1 = alpha1;
    k_blur15<MODE_DIFFUSE><<<blur_grid(w, h), kBlurBlock, 0, st>>>(flow, out, h, w, pa);
}

void launch_median5(const float2* src, float2* dst, int h, int w, cudaStream_t st) {
    PrepArgs pa = {};
    k_median5<MODE_PLAIN><<<median_grid(w, h), dim3(TILE, MTH), 0, st>>>(src, dst, nullptr, h, w, pa);
}

void launch_median5_prep(const float2* src, float2* dst, const float2* blurred, int h, int w, const float* alpha0,
                         const float* alpha1, const float2* G0, const float2* G1, SweepRec* rec, int dir, cudaStream_t st) {
    const PrepArgs pa = make_prep_args(alpha0, alpha1, G0, G1, rec, w, dir);
    k_median5<MODE_PREP><<<median_grid(w, h), dim3(TILE, MTH), 0, st>>>(src, dst, blurred, h, w, pa);
}

}  // namespace pf
