// pf_kernels.cu -- hand-written sm_100a kernels of the PixFlow hot path (round-1 layout: row-major).
//
// Arithmetic contract: every kernel is bit-compatible with the reference CPU path (CPU/PixFlow.hpp,
// CPU/OpticalFlow.cpp and the OpenCV primitives they call).  Compile with
// -fmad=false; never --use_fast_math.
#include "pf_kernels.cuh"
#include "pf_math.cuh"

namespace pf {

static inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

// ====================================================================================================
// front end
// ====================================================================================================
__device__ __forceinline__ int sat_short_rint(float v) {
    int r = __float2int_rn(v);
    return r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
}

__global__ void __launch_bounds__(256)
k_frontend_resize(const uint8_t* __restrict__ bgra, size_t stride, int rows, int cols, int pad,
                  float* __restrict__ grey, float* __restrict__ alpha, int dh, int dw,
                  double scale_x, double scale_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int pcols = cols + 2 * pad;
    int sx, sy, ia[4], ib[4];
    {
        float f, c[4];
        resize_coord(x, scale_x, sx, f);
        cubic_coeffs(f, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) ia[k] = sat_short_rint(fmul(c[k], 2048.0f));
        resize_coord(y, scale_y, sy, f);
        cubic_coeffs(f, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) ib[k] = sat_short_rint(fmul(c[k], 2048.0f));
    }
    int scol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int sc = clampi(sx - 1 + j, 0, pcols - 1) - pad;   // circular pad of prepare()
        if (sc < 0) sc += cols; else if (sc >= cols) sc -= cols;
        scol[j] = sc;
    }
    int hs[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint8_t* S = bgra + (size_t)clampi(sy - 1 + k, 0, rows - 1) * stride;
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uchar4 p = *reinterpret_cast<const uchar4*>(S + (size_t)scol[j] * 4);
            a0 += p.x * ia[j]; a1 += p.y * ia[j]; a2 += p.z * ia[j]; a3 += p.w * ia[j];
        }
        hs[k][0] = a0; hs[k][1] = a1; hs[k][2] = a2; hs[k][3] = a3;
    }
    // vertical pass: float form for the first n - n%8 bytes of the row, integer form for the tail
    const int n = dw * 4, nv = n - n % 8;
    const float sc = 1.0f / 4194304.0f;
    const float b0 = fmul((float)ib[0], sc), b1 = fmul((float)ib[1], sc), b2 = fmul((float)ib[2], sc), b3 = fmul((float)ib[3], sc);
    int px[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int v;
        if (x * 4 + c < nv) {
            const float t = fadd(fadd(fadd(fmul((float)hs[3][c], b3), fmul((float)hs[2][c], b2)), fmul((float)hs[1][c], b1)), fmul((float)hs[0][c], b0));
            v = __float2int_rn(t);
        } else {
            v = (hs[0][c] * ib[0] + hs[1][c] * ib[1] + hs[2][c] * ib[2] + hs[3][c] * ib[3] + (1 << 21)) >> 22;
        }
        px[c] = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
    const int g = (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + 16384) >> 15;   // BGRA2GRAY
    grey[(size_t)y * dw + x] = fmul((float)g, PF_INV255);
    alpha[(size_t)y * dw + x] = fmul((float)px[3], PF_INV255);
}

void launch_frontend_resize(const uint8_t* bgra, size_t stride, int rows, int cols, int pad,
                            float* grey, float* alpha, int dh, int dw, cudaStream_t st) {
    const int pcols = cols + 2 * pad;
    const double scale_x = 1.0 / ((double)dw / (double)pcols);
    const double scale_y = 1.0 / ((double)dh / (double)rows);
    dim3 b(32, 8);
    k_frontend_resize<<<grid2d(dw, dh, b), b, 0, st>>>(bgra, stride, rows, cols, pad, grey, alpha, dh, dw, scale_x, scale_y);
}

__global__ void __launch_bounds__(256)
k_gauss5(const float* __restrict__ src, float* __restrict__ dst, int h, int w) {
    PF_GAUSS_TABLES
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int xm1 = reflect101(x - 1, w), xp1 = reflect101(x + 1, w), xm2 = reflect101(x - 2, w), xp2 = reflect101(x + 2, w);
    float r[5];
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
        const float* S = src + (size_t)reflect101(y + dy, h) * w;
        float s0 = fmul(S[x], kG5[0]);
        s0 = fadd(s0, fmul(fadd(S[xm1], S[xp1]), kG5[1]));
        s0 = fadd(s0, fmul(fadd(S[xm2], S[xp2]), kG5[2]));
        r[dy + 2] = s0;
    }
    float o = fmul(kG5[0], r[2]);
    o = fadd(o, fmul(kG5[1], fadd(r[3], r[1])));
    o = fadd(o, fmul(kG5[2], fadd(r[4], r[0])));
    dst[(size_t)y * w + x] = o;
    (void)kG3H; (void)kG3O; (void)kG15;
}

void launch_gauss5(const float* src, float* dst, int h, int w, cudaStream_t st) {
    dim3 b(32, 8);
    k_gauss5<<<grid2d(w, h, b), b, 0, st>>>(src, dst, h, w);
}

// ====================================================================================================
// pyramid: INTER_LINEAR
// ====================================================================================================
__device__ __forceinline__ void linear_coord_x(int d, double scale, int sw, int& s, float& f) {
    resize_coord(d, scale, s, f);
    if (s < 0) { s = 0; f = 0.0f; }
    if (s >= sw - 1) { s = sw - 1; f = 0.0f; }
}

// Resize coordinates are computed in double precision exactly like OpenCV does -- once per CTA column / row into
// shared memory instead of once per pixel (the FP64 sequence was the bulk of this kernel's instructions).
__global__ void __launch_bounds__(256)
k_pyr_down(PlaneSet ps, int sh, int sw, int dh, int dw, double scale_x, double scale_y) {
    __shared__ int s_xs[32];
    __shared__ float s_xf[32];
    __shared__ int s_y0[8], s_y1[8];
    __shared__ float s_fy[8];
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (threadIdx.y == 0) {
        int s_; float f_;
        linear_coord_x(x < dw ? x : dw - 1, scale_x, sw, s_, f_);
        s_xs[threadIdx.x] = s_; s_xf[threadIdx.x] = f_;
    } else if (threadIdx.y == 1 && threadIdx.x < 8) {
        const int yy = blockIdx.y * 8 + threadIdx.x;
        int sy; float fy;
        resize_coord(yy < dh ? yy : dh - 1, scale_y, sy, fy);
        s_y0[threadIdx.x] = clampi(sy, 0, sh - 1);
        s_y1[threadIdx.x] = clampi(sy + 1, 0, sh - 1);
        s_fy[threadIdx.x] = fy;
    }
    __syncthreads();
    if (x >= dw || y >= dh) return;
    const float* __restrict__ src = ps.src[blockIdx.z];
    float* __restrict__ dst = ps.dst[blockIdx.z];
    const int s = s_xs[threadIdx.x];
    const float f = s_xf[threadIdx.x], fy = s_fy[threadIdx.y];
    const float* S0 = src + s_y0[threadIdx.y] * sw;
    const float* S1 = src + s_y1[threadIdx.y] * sw;
    float r0, r1;
    if (s >= sw - 1) { r0 = S0[s]; r1 = S1[s]; }
    else {
        const float g = fsub(1.0f, f);
        r0 = fadd(fmul(S0[s], g), fmul(S0[s + 1], f));
        r1 = fadd(fmul(S1[s], g), fmul(S1[s + 1], f));
    }
    dst[y * dw + x] = fadd(fmul(r0, fsub(1.0f, fy)), fmul(r1, fy));
}

void launch_pyr_down(const PlaneSet& ps, int nplanes, int sh, int sw, int dh, int dw, cudaStream_t st) {
    const double scale_x = 1.0 / ((double)dw / (double)sw);
    const double scale_y = 1.0 / ((double)dh / (double)sh);
    dim3 b(32, 8);
    dim3 g = grid2d(dw, dh, b);
    g.z = nplanes;
    k_pyr_down<<<g, b, 0, st>>>(ps, sh, sw, dh, dw, scale_x, scale_y);
}

// ====================================================================================================
// gradients: Sobel k=1 (replicate) then 3x3 sigma 0.5 blur (reflect-101), output interleaved (Ix, Iy)
// ====================================================================================================
// Tile of 32 x 8 outputs.  (1) I with a 2-pixel halo into shared memory (replicated borders, which is what the clamped Sobel
// taps read); (2) the Sobel pair (Ix, Iy) ONCE per position of the tile + 1-pixel halo, at the reflect-101 coordinates the
// 3x3 blur will ask for; (3) + (4) the separable blur, row pass then column pass, both channels of a pair in one packed
// fp32x2 instruction.  Same taps, same order, same roundings as computing every Sobel value at every use.
__global__ void __launch_bounds__(256)
k_gradient(const float* __restrict__ I, float2* __restrict__ G, int h, int w) {
    PF_GAUSS_TABLES
    __shared__ float s_I[12][36 + 1];
    __shared__ f2p s_S[10][34];
    __shared__ f2p s_R[10][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    {
        const int gx0 = clampi(x0 - 2 + tx, 0, w - 1), gx1 = clampi(x0 - 2 + tx + 32, 0, w - 1);
        const int gy0 = clampi(y0 - 2 + ty, 0, h - 1), gy1 = clampi(y0 - 2 + ty + 8, 0, h - 1);
        s_I[ty][tx] = I[gy0 * w + gx0];
        if (tx < 4) s_I[ty][tx + 32] = I[gy0 * w + gx1];
        if (ty < 4) {
            s_I[ty + 8][tx] = I[gy1 * w + gx0];
            if (tx < 4) s_I[ty + 8][tx + 32] = I[gy1 * w + gx1];
        }
    }
    __syncthreads();
    // local (shared-memory) coordinates of global column c / row r: c - (x0 - 2), r - (y0 - 2)
    const int ox = x0 - 2, oy = y0 - 2;
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
        const int ly = ty + 8 * ry;
        if (ly < 10) {
            const int Y = reflect101(min(y0 - 1 + ly, h), h) - oy;       // rows past h are only used by outputs past the image
#pragma unroll
            for (int rx = 0; rx < 2; ++rx) {
                const int lx = tx + 32 * rx;
                if (lx < 34) {
                    const int X = reflect101(min(x0 - 1 + lx, w), w) - ox;
                    // Sobel k = 1 with BORDER_REPLICATE: the tile holds replicated values outside the image
                    s_S[ly][lx] = pk(fsub(s_I[Y][X + 1], s_I[Y][X - 1]), fsub(s_I[Y + 1][X], s_I[Y - 1][X]));
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
        const int ly = ty + 8 * ry;
        if (ly < 10)
            s_R[ly][tx] = padd(pmuls(s_S[ly][tx + 1], kG3H[0]), pmuls(padd(s_S[ly][tx], s_S[ly][tx + 2]), kG3H[1]));
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) return;
    G[y * w + x] = upk(padd(pmuls(s_R[ty + 1][tx], kG3H[0]), pmuls(padd(s_R[ty + 2][tx], s_R[ty][tx]), kG3H[1])));
    (void)kG5; (void)kG3O; (void)kG15;
}

void launch_gradient(const float* I, float2* G, int h, int w, cudaStream_t st) {
    dim3 b(32, 8);
    k_gradient<<<grid2d(w, h, b), b, 0, st>>>(I, G, h, w);
}

// ====================================================================================================
// inter-level upsample: INTER_CUBIC on float2, then * (1/0.9)
// ====================================================================================================
__global__ void __launch_bounds__(256)
k_upsample_cubic(const float2* __restrict__ src, int sh, int sw, int sp, float2* __restrict__ dst, int dh, int dw, int dp,
                 double scale_x, double scale_y) {
    __shared__ int s_sx[32], s_sy[8];
    __shared__ float s_ca[32][4], s_cb[8][4];
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (threadIdx.y == 0) {
        int s_; float f, c[4];
        resize_coord(x < dw ? x : dw - 1, scale_x, s_, f); cubic_coeffs(f, c);
        s_sx[threadIdx.x] = s_;
#pragma unroll
        for (int k = 0; k < 4; ++k) s_ca[threadIdx.x][k] = c[k];
    } else if (threadIdx.y == 1 && threadIdx.x < 8) {
        const int yy = blockIdx.y * 8 + threadIdx.x;
        int s_; float f, c[4];
        resize_coord(yy < dh ? yy : dh - 1, scale_y, s_, f); cubic_coeffs(f, c);
        s_sy[threadIdx.x] = s_;
#pragma unroll
        for (int k = 0; k < 4; ++k) s_cb[threadIdx.x][k] = c[k];
    }
    __syncthreads();
    if (x >= dw || y >= dh) return;
    const int sx = s_sx[threadIdx.x], sy = s_sy[threadIdx.y];
    float ca[4], cb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ca[k] = s_ca[threadIdx.x][k]; cb[k] = s_cb[threadIdx.y][k]; }
    int xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = clampi(sx - 1 + k, 0, sw - 1);
    // both channels of a flow vector in one packed fp32x2 instruction per tap (pf_math.cuh)
    f2p r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f2p* S = reinterpret_cast<const f2p*>(src + clampi(sy - 1 + k, 0, sh - 1) * sp);
        r[k] = padd(padd(padd(pmuls(S[xs[0]], ca[0]), pmuls(S[xs[1]], ca[1])), pmuls(S[xs[2]], ca[2])), pmuls(S[xs[3]], ca[3]));
    }
    // vertical pass: right-to-left for the first n - n%4 floats of the row, left-to-right for the tail (n = 2*dw is even and
    // n - n%4 a multiple of 4, so both floats of a pixel fall on the same side)
    const int n = dw * 2, nv = n - n % 4;
    f2p o;
    if (2 * x < nv) o = padd(padd(padd(pmuls(r[3], cb[3]), pmuls(r[2], cb[2])), pmuls(r[1], cb[1])), pmuls(r[0], cb[0]));
    else            o = padd(padd(padd(pmuls(r[0], cb[0]), pmuls(r[1], cb[1])), pmuls(r[2], cb[2])), pmuls(r[3], cb[3]));
    dst[y * dp + x] = upk(pmuls(o, PF_INV_PYR));
}

void launch_upsample_cubic(const float2* src, int sh, int sw, int sp, float2* dst, int dh, int dw, int dp, cudaStream_t st) {
    const double scale_x = 1.0 / ((double)dw / (double)sw);
    const double scale_y = 1.0 / ((double)dh / (double)sh);
    dim3 b(32, 8);
    k_upsample_cubic<<<grid2d(dw, dh, b), b, 0, st>>>(src, sh, sw, sp, dst, dh, dw, dp, scale_x, scale_y);
}

// ====================================================================================================
// tail: INTER_LINEAR to (rows x pcols), *2, 3x3 sigma 1 blur; writes the cropped columns only
// ====================================================================================================
// Tile of 32 x 8 outputs: the INTER_LINEAR-upsampled (x2-scaled) flow is computed ONCE per position of the tile + 1 px
// halo into shared memory (reflect-101 of the 3x3 blur applied to the coordinates of the padded full-size image), then
// the 3x3 sigma-1 row and column passes run from shared memory.
__global__ void __launch_bounds__(256)
k_tail(const float2* __restrict__ src, int sh, int sw, int sp, int rows, int pcols, int pad, int cols,
       float2* __restrict__ out, size_t out_stride, double scale_x, double scale_y) {
    PF_GAUSS_TABLES
    __shared__ int s_xs[34];
    __shared__ float s_xf[34];
    __shared__ int s_y0[10], s_y1[10];
    __shared__ float s_fy[10];
    __shared__ f2p s_up[10][34];
    __shared__ f2p s_rb[10][32];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int xo0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    if (tid < 34) {                       // columns xo0-1 .. xo0+32 of the crop = padded columns (+pad), reflected
        const int xx = reflect101(xo0 + pad - 1 + tid, pcols);
        int s_; float f_;
        linear_coord_x(xx, scale_x, sw, s_, f_);
        s_xs[tid] = s_; s_xf[tid] = f_;
    } else if (tid >= 64 && tid < 74) {   // rows y0-1 .. y0+8, reflected
        const int k = tid - 64;
        const int yy = reflect101(y0 - 1 + k, rows);
        int sy; float fy;
        resize_coord(yy, scale_y, sy, fy);
        s_y0[k] = clampi(sy, 0, sh - 1);
        s_y1[k] = clampi(sy + 1, 0, sh - 1);
        s_fy[k] = fy;
    }
    __syncthreads();
    const f2p* srcp = reinterpret_cast<const f2p*>(src);
    for (int e = tid; e < 10 * 34; e += 256) {
        const int ly = e / 34, lx = e - ly * 34;
        const f2p* S0 = srcp + s_y0[ly] * sp;
        const f2p* S1 = srcp + s_y1[ly] * sp;
        const int s = s_xs[lx];
        const float b1 = s_fy[ly], b0 = fsub(1.0f, b1);
        f2p r0, r1;
        if (s >= sw - 1) { r0 = S0[s]; r1 = S1[s]; }
        else {
            const float f = s_xf[lx], g = fsub(1.0f, f);
            r0 = padd(pmuls(S0[s], g), pmuls(S0[s + 1], f));
            r1 = padd(pmuls(S1[s], g), pmuls(S1[s + 1], f));
        }
        s_up[ly][lx] = pmuls(padd(pmuls(r0, b0), pmuls(r1, b1)), 2.0f);      // flow *= 1/downscaleFactor
    }
    __syncthreads();
    for (int e = tid; e < 10 * 32; e += 256) {      // row pass of the 3x3 blur
        const int ly = e >> 5, lx = e & 31;
        s_rb[ly][lx] = padd(pmuls(s_up[ly][lx + 1], kG3O[0]), pmuls(padd(s_up[ly][lx], s_up[ly][lx + 2]), kG3O[1]));
    }
    __syncthreads();
    const int xo = xo0 + tx, y = y0 + ty;
    if (xo >= cols || y >= rows) return;
    const float2 o = upk(padd(pmuls(s_rb[ty + 1][tx], kG3O[0]), pmuls(padd(s_rb[ty + 2][tx], s_rb[ty][tx]), kG3O[1])));
    *reinterpret_cast<float2*>(reinterpret_cast<char*>(out) + (size_t)y * out_stride + (size_t)xo * sizeof(float2)) = o;
    (void)kG5; (void)kG3H; (void)kG15;
}

void launch_tail(const float2* flow0, int sh, int sw, int sp, int rows, int pcols, int pad, int cols,
                 float2* out, size_t out_stride, cudaStream_t st) {
    const double scale_x = 1.0 / ((double)pcols / (double)sw);
    const double scale_y = 1.0 / ((double)rows / (double)sh);
    dim3 b(32, 8);
    k_tail<<<grid2d(cols, rows, b), b, 0, st>>>(flow0, sh, sw, sp, rows, pcols, pad, cols, out, out_stride, scale_x, scale_y);
}

// ====================================================================================================
// coarsest-level search
// ====================================================================================================
// computeIntensityRatio: sequential fp32 sums in raster order.  The sums are order-dependent, so ONE thread adds; the products
// (independent, same operations) are formed by the whole CTA, a chunk at a time, into shared memory, so that the adding thread
// walks a 4-cycle fadd chain instead of a chain of dependent global loads (100 -> ~10 us on a 25 x 45 level, which sits on the
// critical path of each direction).
constexpr int RATIO_CHUNK = 2048;
__global__ void __launch_bounds__(256)
k_intensity_ratio(const float* __restrict__ I0, const float* __restrict__ a0,
                  const float* __restrict__ I1, const float* __restrict__ a1, int n, float* ratio) {
    __shared__ float2 s_p[RATIO_CHUNK];
    float sumL = 0.0f, sumR = 0.0f;
    for (int base = 0; base < n; base += RATIO_CHUNK) {
        const int m = min(RATIO_CHUNK, n - base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const float al = fmul(a0[base + i], a1[base + i]);
            s_p[i] = make_float2(fmul(al, I0[base + i]), fmul(al, I1[base + i]));
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < m; ++i) {
                const float2 p = s_p[i];
                sumL = fadd(sumL, p.x);
                sumR = fadd(sumR, p.y);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) ratio[0] = __fdiv_rn(sumL, sumR);
}

// computePatchError, CPU/PixFlow.hpp:157-188 (I1eq = I1 * ratio evaluated on the fly)
__device__ float patch_error(const float* __restrict__ i0, const float* __restrict__ a0, int i0x, int i0y,
                             const float* __restrict__ i1, const float* __restrict__ a1, int i1x, int i1y,
                             int w, int h, float ratio, int dist) {
    float sad = 0.0f, alpha = 0.0f;
    for (int dy = -2; dy <= 2; ++dy) {
        const int d0y = i0y + dy;
        if (0 <= d0y && d0y < h) {
            const int d1y = clampi(i1y + dy, 0, h - 1);
            for (int dx = -2; dx <= 2; ++dx) {
                const int d0x = i0x + dx;
                if (0 <= d0x && d0x < w) {
                    const int d1x = clampi(i1x + dx, 0, w - 1);
                    const float diff = fsub(i0[(size_t)d0y * w + d0x], fmul(i1[(size_t)d1y * w + d1x], ratio));
                    sad = fadd(sad, fabsf(diff));
                    alpha = fadd(alpha, fmul(a0[(size_t)d0y * w + d0x], a1[(size_t)d1y * w + d1x]));
                }
            }
        }
    }
    sad = __fdiv_rn(sad, alpha);
    const double fx = (double)(i1x - i0x), fy = (double)(i1y - i0y);
    const float length = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy))));
    sad = fmul(sad, fadd(1.0f, __fdiv_rn(length, (float)dist)));
    return sad;
}

__global__ void __launch_bounds__(128)
k_adjust_initial_flow(const float* __restrict__ I0, const float* __restrict__ I1,
                      const float* __restrict__ a0, const float* __restrict__ a1, float2* __restrict__ flow,
                      const float* __restrict__ ratio_p, int h, int w, int fp, int bx, int by, int bw, int bh, int dist) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float2 out = make_float2(0.0f, 0.0f);
    if (bw > 0 && a0[(size_t)y * w + x] > PF_ALPHA_THRESHOLD) {
        const float ratio = ratio_p[0];
        float errorBest = fmul(0.8f, patch_error(I0, a0, x, y, I1, a1, x, y, w, h, ratio, dist));
        int bestx = x, besty = y;
        for (int dy = by; dy < by + bh; ++dy)
            for (int dx = bx; dx < bx + bw; ++dx) {
                const int i1x = x + dx, i1y = y + dy;
                if (0 <= i1x && i1x < w && 0 <= i1y && i1y < h) {
                    const float e = patch_error(I0, a0, x, y, I1, a1, i1x, i1y, w, h, ratio, dist);
                    if (errorBest > e) { errorBest = e; bestx = i1x; besty = i1y; }
                }
            }
        out = make_float2((float)(bestx - x), (float)(besty - y));
    }
    flow[(size_t)y * fp + x] = out;
}

void launch_initial_flow(const float* I0, const float* I1, const float* alpha0, const float* alpha1,
                         float2* flow, int fp, float* ratio, int h, int w, int hint, int dist, cudaStream_t st) {
    int bx = 0, by = 0, bw = 0, bh = 0;
    if (dist > 0 && hint >= 1 && hint <= 4) {   // computeSearchBox, CPU/PixFlow.hpp:207-224
        const int ortho = (dist + 8 / 2) / 8, thick = 2 * ortho + 1;
        switch (hint) {
        case 1: bx = 0; by = -ortho; bw = dist + 1; bh = thick; break;        // RIGHT
        case 2: bx = -ortho; by = 0; bw = thick; bh = dist + 1; break;        // DOWN
        case 3: bx = -dist; by = -ortho; bw = dist + 1; bh = thick; break;    // LEFT
        case 4: bx = -ortho; by = -dist; bw = thick; bh = dist + 1; break;    // UP
        }
        k_intensity_ratio<<<1, 256, 0, st>>>(I0, alpha0, I1, alpha1, h * w, ratio);
    }
    dim3 b(32, 4);
    k_adjust_initial_flow<<<grid2d(w, h, b), b, 0, st>>>(I0, I1, alpha0, alpha1, flow, ratio, h, w, fp, bx, by, bw, bh, dist);
}

// ====================================================================================================
// combineNovelViews
// ====================================================================================================
__device__ __forceinline__ uchar4 novel_view_point(const uint8_t* __restrict__ img, size_t stride, float2 f, double t,
                                                   int x, int y, int rows, int cols) {
    int srcx = __double2int_rz(__dadd_rn((double)x, __dmul_rn((double)f.x, t)));
    if (srcx > cols - 1) srcx -= cols;
    if (srcx < 0) srcx += cols;
    int srcy = __double2int_rz(__dadd_rn((double)y, __dmul_rn((double)f.y, t)));
    if (srcy > rows - 1) srcy = rows - 1;
    if (srcy < 0) srcy = 0;
    srcx = clampi(srcx, 0, cols - 1);   // the reference would read out of bounds here; never taken for sane flows
    return *reinterpret_cast<const uchar4*>(img + (size_t)srcy * stride + (size_t)srcx * 4);
}

__global__ void __launch_bounds__(256)
k_combine(const uint8_t* __restrict__ imageL, size_t strideL, const uint8_t* __restrict__ imageR, size_t strideR,
          const float2* __restrict__ flowLR, size_t strideLR, const float2* __restrict__ flowRL, size_t strideRL,
          const float* __restrict__ blend, size_t strideB, int rows, int cols, uint8_t* __restrict__ out, size_t strideOut) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float blendR = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(blend) + (size_t)y * strideB + (size_t)x * 4);
    const float blendL = fsub(1.0f, blendR);
    const float2 fLR = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flowLR) + (size_t)y * strideLR + (size_t)x * 8);
    const float2 fRL = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flowRL) + (size_t)y * strideRL + (size_t)x * 8);
    const uchar4 cL = novel_view_point(imageL, strideL, fRL, (double)blendR, x, y, rows, cols);
    const uchar4 cR = novel_view_point(imageR, strideR, fLR, (double)blendL, x, y, rows, cols);
    uchar4 o = make_uchar4(0, 0, 0, 0);
    if (cL.w != 0 && cR.w != 0) {
        const float fc = (float)cols;
        const float magLR = __fdiv_rn(__fsqrt_rn(fadd(fmul(fLR.x, fLR.x), fmul(fLR.y, fLR.y))), fc);
        const float magRL = __fdiv_rn(__fsqrt_rn(fadd(fmul(fRL.x, fRL.x), fmul(fRL.y, fRL.y))), fc);
        const int di = abs((int)cL.x - (int)cR.x) + abs((int)cL.y - (int)cR.y) + abs((int)cL.z - (int)cR.z);
        const float colorDiff = __fdiv_rn((float)di, 255.0f);
        const float deghost = tanhf(fmul(colorDiff, 10.0f));
        const float alphaL = __fdiv_rn((float)cL.w, 255.0f), alphaR = __fdiv_rn((float)cR.w, 255.0f);
        const double aL = (double)fmul(fmul(10.0f, blendL), alphaL);
        const double aR = (double)fmul(fmul(10.0f, blendR), alphaR);
        const double expL = exp(__dmul_rn(aL, __dadd_rn(1.0, (double)fmul(100.0f, magRL))));
        const double expR = exp(__dmul_rn(aR, __dadd_rn(1.0, (double)fmul(100.0f, magLR))));
        const double sumExp = __dadd_rn(__dadd_rn(expL, expR), 0.00001);
        const float smL = __double2float_rn(__ddiv_rn(expL, sumExp));
        const float smR = __double2float_rn(__ddiv_rn(expR, sumExp));
        const float wL = fadd(fmul(blendL, fsub(1.0f, deghost)), fmul(smL, deghost));   // lerp, util.hpp:98-101
        const float wR = fadd(fmul(blendR, fsub(1.0f, deghost)), fmul(smR, deghost));
        o.x = (unsigned char)__float2int_rz(fadd(fmul((float)cL.x, wL), fmul((float)cR.x, wR)));
        o.y = (unsigned char)__float2int_rz(fadd(fmul((float)cL.y, wL), fmul((float)cR.y, wR)));
        o.z = (unsigned char)__float2int_rz(fadd(fmul((float)cL.z, wL), fmul((float)cR.z, wR)));
        o.w = 255;
    }
    *reinterpret_cast<uchar4*>(out + (size_t)y * strideOut + (size_t)x * 4) = o;
}

void launch_combine(const uint8_t* imageL, size_t strideL, const uint8_t* imageR, size_t strideR,
                    const float2* flowLR, size_t strideLR, const float2* flowRL, size_t strideRL,
                    const float* blend, size_t strideB, int rows, int cols,
                    uint8_t* out, size_t strideOut, cudaStream_t st) {
    dim3 b(32, 8);
    k_combine<<<grid2d(cols, rows, b), b, 0, st>>>(imageL, strideL, imageR, strideR, flowLR, strideLR, flowRL, strideRL,
                                                  blend, strideB, rows, cols, out, strideOut);
}

}  // namespace pf
