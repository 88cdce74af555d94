// pf_stitch.cu -- first "next" row of the hot-path contract: the canvas map, the overlap masking
// and the 8-direction blend-weight search of Stitchtools::prepare / GenerateBlend / countblend
// (CPU/StitchTool.cpp:7-50, :98-131, :148-191; the reference's own CUDA twin: countblend_Kernel,
// GPU/StitchTool_GPU.cu:10-66 -- which uses the literal 1.4142 where the CPU path uses sqrt(2); this follows the CPU path).
// The block-wise in-place blur that follows in GenerateBlend (:133-145) is order-dependent OpenCV ROI filtering and is
// NOT part of this unit.
#include "pf_kernels.cuh"
#include "pf_math.cuh"

namespace pf {

// MatchImages (:38-50) + overlap masking (:16-33): Map = 100*[alphaL>0] + 50*[alphaR>0]; images kept where Map > 140
__global__ void __launch_bounds__(256)
k_stitch_match_mask(const uint8_t* __restrict__ L, size_t strideL, const uint8_t* __restrict__ R, size_t strideR, int rows, int cols,
                    uint8_t* __restrict__ map, size_t strideM, uint8_t* __restrict__ oL, size_t strideOL,
                    uint8_t* __restrict__ oR, size_t strideOR) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const uchar4 l = *reinterpret_cast<const uchar4*>(L + (size_t)y * strideL + (size_t)x * 4);
    const uchar4 r = *reinterpret_cast<const uchar4*>(R + (size_t)y * strideR + (size_t)x * 4);
    const int m = (l.w > 0 ? 100 : 0) + (r.w > 0 ? 50 : 0);
    map[(size_t)y * strideM + x] = (uint8_t)m;
    const uchar4 z = make_uchar4(0, 0, 0, 0);
    *reinterpret_cast<uchar4*>(oL + (size_t)y * strideOL + (size_t)x * 4) = m > 140 ? l : z;
    *reinterpret_cast<uchar4*>(oR + (size_t)y * strideOR + (size_t)x * 4) = m > 140 ? r : z;
}

// GenerateBlend (:113-124) + countblend (:148-191).  The map is read through its circular extension by len = cols/5
// (:101-111): extended column xe maps to source column (xe - len) mod cols.
__global__ void __launch_bounds__(256)
k_stitch_blend_raw(const uint8_t* __restrict__ map, size_t strideM, int rows, int cols, int len, int step,
                   float* __restrict__ blend, size_t strideB, float* __restrict__ mdis, size_t strideD) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const int ecols = cols + 2 * len;
    auto EM = [&](int yy, int xe) -> int {
        int sc = xe - len;
        if (sc < 0) sc += cols; else if (sc >= cols) sc -= cols;
        return map[(size_t)yy * strideM + sc];
    };
    const int xe = x + len;
    const int m = EM(y, xe);
    float b, md = 0.0f;
    if (m == 100) b = 0.0f;
    else if (m == 50) b = 1.0f;
    else if (m == 150) {
        float minL = (float)(10 * cols), minR = (float)(10 * cols);
        const double sqrt2 = 1.4142135623730951;                    // sqrt(2) in double, as the CPU path computes it
        for (int i = 0; i < cols / 2; i += step) {
            const float fi = (float)i;
            const double di = __dmul_rn((double)i, sqrt2);
            const float fd = __double2float_rn(di);
            const bool xp = xe + i < ecols, xm = xe - i > 0, yp = y + i < rows, ym = y - i > 0;
            int v;
            if (xp) { v = EM(y, xe + i); if (v == 100 && fi < minL) minL = fi; if (v == 50 && fi < minR) minR = fi; }
            if (xm) { v = EM(y, xe - i); if (v == 100 && fi < minL) minL = fi; if (v == 50 && fi < minR) minR = fi; }
            if (yp) { v = EM(y + i, xe); if (v == 100 && fi < minL) minL = fi; if (v == 50 && fi < minR) minR = fi; }
            if (ym) { v = EM(y - i, xe); if (v == 100 && fi < minL) minL = fi; if (v == 50 && fi < minR) minR = fi; }
            if (xp && yp) { v = EM(y + i, xe + i); if (v == 100 && di < (double)minL) minL = fd; if (v == 50 && di < (double)minR) minR = fd; }
            if (xm && ym) { v = EM(y - i, xe - i); if (v == 100 && di < (double)minL) minL = fd; if (v == 50 && di < (double)minR) minR = fd; }
            if (xp && ym) { v = EM(y - i, xe + i); if (v == 100 && di < (double)minL) minL = fd; if (v == 50 && di < (double)minR) minR = fd; }
            if (xm && yp) { v = EM(y + i, xe - i); if (v == 100 && di < (double)minL) minL = fd; if (v == 50 && di < (double)minR) minR = fd; }
        }
        b = __fdiv_rn(minL, fadd(minR, minL));
        md = (minL < minR) ? minL : minR;
    } else b = 0.5f;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(blend) + (size_t)y * strideB + (size_t)x * 4) = b;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(mdis) + (size_t)y * strideD + (size_t)x * 4) = md;
}

void launch_stitch_match_mask(const uint8_t* L, size_t strideL, const uint8_t* R, size_t strideR, int rows, int cols,
                              uint8_t* map, size_t strideM, uint8_t* oL, size_t strideOL, uint8_t* oR, size_t strideOR, cudaStream_t st) {
    dim3 b(32, 8), g((cols + 31) / 32, (rows + 7) / 8);
    k_stitch_match_mask<<<g, b, 0, st>>>(L, strideL, R, strideR, rows, cols, map, strideM, oL, strideOL, oR, strideOR);
}

void launch_stitch_blend_raw(const uint8_t* map, size_t strideM, int rows, int cols, float* blend, size_t strideB,
                             float* mdis, size_t strideD, cudaStream_t st) {
    const int step = (cols <= rows) ? cols / 200 : rows / 200;
    dim3 b(32, 8), g((cols + 31) / 32, (rows + 7) / 8);
    k_stitch_blend_raw<<<g, b, 0, st>>>(map, strideM, rows, cols, cols / 5, step, blend, strideB, mdis, strideD);
}

}  // namespace pf
