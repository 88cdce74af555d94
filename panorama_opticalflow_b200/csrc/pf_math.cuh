// pf_math.cuh -- exact-arithmetic device helpers shared by all kernels.
//
// Everything here reproduces, operation by operation (fp32, separately rounded multiply and add,
// IEEE sqrt/div), the arithmetic of the reference CPU path: the loops of CPU/PixFlow.hpp and the
// OpenCV primitives they call.  The flow iteration is chaotically sensitive to
// rounding, so no FMA contraction is allowed: the library is compiled
// with -fmad=false and the sensitive expressions additionally use explicit __f*_rn intrinsics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pf {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// ---- Blackwell packed fp32x2 arithmetic (FFMA2 / FADD2, sm_100+) -------------------------------------------------------
// Two independent IEEE round-to-nearest fp32 operations per instruction: the (x, y) channels of a flow vector or of a
// gradient pair go through identical arithmetic, so every packed result is bit-identical to the two scalar __f*_rn
// operations it replaces, at half the issue slots (the FMA pipe is occupied for two cycles; measured on B200 with
// tools/microbench_f32x2.cu: same 4.5-cycle dependent latency as the scalar forms).
// ptxas 12.9 contracts mul.f32x2 + add.f32x2 (and __fmul2_rn + __fadd2_rn) into one FFMA2 even under --fmad=false, which
// would change the rounding; the packed multiply is therefore written as fma(a, b, {-0,-0}) with the -0 pair read from
// constant memory, opaque to the compiler: RN(a*b + (-0)) == RN(a*b) for every a, b (signed zeros included), and an fma
// cannot be contracted with a following add.  Checked on 2^24 random bit patterns by the same micro-benchmark.
typedef unsigned long long f2p;            // {lo = x, hi = y}: the register pair of a float2
static __constant__ f2p c_pf_negzero2 = 0x8000000080000000ull;
__device__ __forceinline__ f2p pk(float lo, float hi) { f2p r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2p pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ float2 upk(f2p v) { float2 o; asm("mov.b64 {%0,%1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(v)); return o; }
__device__ __forceinline__ f2p pfma(f2p a, f2p b, f2p c) { f2p r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2p padd(f2p a, f2p b) { f2p r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2p psub(f2p a, f2p b) { f2p r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2p pmul(f2p a, f2p b) { return pfma(a, b, c_pf_negzero2); }
__device__ __forceinline__ f2p pmuls(f2p a, float s) { return pfma(a, pk(s, s), c_pf_negzero2); }
__device__ __forceinline__ f2p pneg(f2p a) { return psub(0ull, a); }      // only used where the sign of a zero is irrelevant

__device__ __forceinline__ int clampi(int x, int a, int b) { return x < a ? a : (x > b ? b : x); }

// BORDER_REFLECT_101
__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// OpenCV resize coordinate: f = float((d+0.5)*scale-0.5) in double; s=floor(f); f-=s
__device__ __forceinline__ void resize_coord(int d, double scale, int& s, float& f) {
    const double t = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    const float fx = __double2float_rn(t);
    const float fl = floorf(fx);
    s = (int)fl;
    f = fsub(fx, fl);
}

// OpenCV interpolateCubic, A = -0.75, evaluated left to right in fp32
__device__ __forceinline__ void cubic_coeffs(float x, float c[4]) {
    const float A = -0.75f;
    const float x1 = fadd(x, 1.0f);
    c[0] = fsub(fmul(fadd(fmul(fsub(fmul(A, x1), 5 * A), x1), 8 * A), x1), 4 * A);
    c[1] = fadd(fmul(fmul(fsub(fmul(A + 2, x), A + 3), x), x), 1.0f);
    const float xm = fsub(1.0f, x);
    c[2] = fadd(fmul(fmul(fsub(fmul(A + 2, xm), A + 3), xm), xm), 1.0f);
    c[3] = fsub(fsub(fsub(1.0f, c[0]), c[1]), c[2]);
}

// Gaussian kernels = float32(w_i / sum w), w in double, as exact hex floats; index 0 is
// the centre tap.  tests/test_host_logic.py checks them against the oracle's kernel generator.
#define PF_GAUSS_TABLES \
    /* (5, 0.25) pre-blur, CPU/PixFlow.hpp:102 */ \
    constexpr float kG5[3] = {0x1.ffa81ep-1f, 0x1.5f85bp-12f, 0x1.c7f7fep-47f}; \
    /* (3, 0.5) gradient blur, :291 */ \
    constexpr float kG3H[2] = {0x1.92efd6p-1f, 0x1.b440aap-4f}; \
    /* (3, 1.0) final flow blur, :130 */ \
    constexpr float kG3O[2] = {0x1.ceb51cp-2f, 0x1.18a572p-2f}; \
    /* (15, 8.0) flow blur, :307/:390 */ \
    constexpr float kG15[8] = {0x1.395eaep-4f, 0x1.36ee62p-4f, 0x1.2fba7ep-4f, 0x1.2417c4p-4f, \
                               0x1.148c44p-4f, 0x1.01c56p-4f, 0x1.d91684p-5f, 0x1.ab663cp-5f};
#define PF_INV255 0x1.010102p-8f      /* float(1/255.)  : Mat /= 255.0f */
#define PF_INV_PYR 0x1.1c71c8p+0f     /* 1.0f / 0.9f    : CPU/PixFlow.hpp:124 */

// median of 25 by a 99-comparator selection network (verified exhaustively with the 0-1 principle,
// tests/test_host_logic.py); pure min/max selection == medianBlur(32FC2, 5) per channel
#define PF_CSWAP(a, b) { const float lo_ = fminf(v[a], v[b]); const float hi_ = fmaxf(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
__device__ __forceinline__ float median25(float v[25]) {
    PF_CSWAP(0,1) PF_CSWAP(3,4) PF_CSWAP(2,4) PF_CSWAP(2,3) PF_CSWAP(6,7) PF_CSWAP(5,7) PF_CSWAP(5,6) PF_CSWAP(9,10)
    PF_CSWAP(8,10) PF_CSWAP(8,9) PF_CSWAP(12,13) PF_CSWAP(11,13) PF_CSWAP(11,12) PF_CSWAP(15,16) PF_CSWAP(14,16)
    PF_CSWAP(14,15) PF_CSWAP(18,19) PF_CSWAP(17,19) PF_CSWAP(17,18) PF_CSWAP(21,22) PF_CSWAP(20,22) PF_CSWAP(20,21)
    PF_CSWAP(23,24) PF_CSWAP(2,5) PF_CSWAP(3,6) PF_CSWAP(0,6) PF_CSWAP(0,3) PF_CSWAP(4,7) PF_CSWAP(1,7) PF_CSWAP(1,4)
    PF_CSWAP(11,14) PF_CSWAP(8,14) PF_CSWAP(8,11) PF_CSWAP(12,15) PF_CSWAP(9,15) PF_CSWAP(9,12) PF_CSWAP(13,16)
    PF_CSWAP(10,16) PF_CSWAP(10,13) PF_CSWAP(20,23) PF_CSWAP(17,23) PF_CSWAP(17,20) PF_CSWAP(21,24) PF_CSWAP(18,24)
    PF_CSWAP(18,21) PF_CSWAP(19,22) PF_CSWAP(8,17) PF_CSWAP(9,18) PF_CSWAP(0,18) PF_CSWAP(0,9) PF_CSWAP(10,19)
    PF_CSWAP(1,19) PF_CSWAP(1,10) PF_CSWAP(11,20) PF_CSWAP(2,20) PF_CSWAP(2,11) PF_CSWAP(12,21) PF_CSWAP(3,21)
    PF_CSWAP(3,12) PF_CSWAP(13,22) PF_CSWAP(4,22) PF_CSWAP(4,13) PF_CSWAP(14,23) PF_CSWAP(5,23) PF_CSWAP(5,14)
    PF_CSWAP(15,24) PF_CSWAP(6,24) PF_CSWAP(6,15) PF_CSWAP(7,16) PF_CSWAP(7,19) PF_CSWAP(13,21) PF_CSWAP(15,23)
    PF_CSWAP(7,13) PF_CSWAP(7,15) PF_CSWAP(1,9) PF_CSWAP(3,11) PF_CSWAP(5,17) PF_CSWAP(11,17) PF_CSWAP(9,17)
    PF_CSWAP(4,10) PF_CSWAP(6,12) PF_CSWAP(7,14) PF_CSWAP(4,6) PF_CSWAP(4,7) PF_CSWAP(12,14) PF_CSWAP(10,14)
    PF_CSWAP(6,7) PF_CSWAP(10,12) PF_CSWAP(6,10) PF_CSWAP(6,17) PF_CSWAP(12,17) PF_CSWAP(7,17) PF_CSWAP(7,10)
    PF_CSWAP(12,18) PF_CSWAP(7,12) PF_CSWAP(10,18) PF_CSWAP(12,20) PF_CSWAP(10,20) PF_CSWAP(10,12)
    return v[12];
}

// The same selection network with its 21 three-element sorts (bubble triples (j,k),(i,k),(i,j) of the list above) done by
// Blackwell's 3-input min/max: lo = min3, hi = max3, mid = a ^ b ^ c ^ lo ^ hi on the bit patterns (lo and hi are bit copies
// of two of the inputs, so the XOR leaves exactly the third).  FMNMX, FMNMX3 and the logic ops all issue on the half-rate ALU
// pipe (tools/microbench_minmax.cu), which bounds the median kernels: 4 ALU operations per sorted triple instead of 6.
// NaN-free inputs only (the XOR identity needs min3/max3 to return input bit patterns); the callers' flows are finite.
#define PF_SORT3(a, b, c) { const float x_ = v[a], y_ = v[b], z_ = v[c]; \
    const float lo_ = fminf(fminf(x_, y_), z_), hi_ = fmaxf(fmaxf(x_, y_), z_); \
    v[b] = __int_as_float(__float_as_int(x_) ^ __float_as_int(y_) ^ __float_as_int(z_) ^ __float_as_int(lo_) ^ __float_as_int(hi_)); \
    v[a] = lo_; v[c] = hi_; }
__device__ __forceinline__ float median25_s3(float v[25]) {
    PF_CSWAP(0,1) PF_SORT3(2,3,4) PF_SORT3(5,6,7) PF_SORT3(8,9,10) PF_SORT3(11,12,13) PF_SORT3(14,15,16)
    PF_SORT3(17,18,19) PF_SORT3(20,21,22) PF_CSWAP(23,24) PF_CSWAP(2,5) PF_SORT3(0,3,6) PF_SORT3(1,4,7)
    PF_SORT3(8,11,14) PF_SORT3(9,12,15) PF_SORT3(10,13,16) PF_SORT3(17,20,23) PF_SORT3(18,21,24) PF_CSWAP(19,22)
    PF_CSWAP(8,17) PF_SORT3(0,9,18) PF_SORT3(1,10,19) PF_SORT3(2,11,20) PF_SORT3(3,12,21) PF_SORT3(4,13,22)
    PF_SORT3(5,14,23) PF_SORT3(6,15,24) PF_CSWAP(7,16) PF_CSWAP(7,19) PF_CSWAP(13,21) PF_CSWAP(15,23) PF_CSWAP(7,13)
    PF_CSWAP(7,15) PF_CSWAP(1,9) PF_CSWAP(3,11) PF_CSWAP(5,17) PF_CSWAP(11,17) PF_CSWAP(9,17) PF_CSWAP(4,10)
    PF_CSWAP(6,12) PF_CSWAP(7,14) PF_CSWAP(4,6) PF_CSWAP(4,7) PF_CSWAP(12,14) PF_CSWAP(10,14) PF_CSWAP(6,7)
    PF_CSWAP(10,12) PF_CSWAP(6,10) PF_CSWAP(6,17) PF_CSWAP(12,17) PF_CSWAP(7,17) PF_CSWAP(7,10) PF_CSWAP(12,18)
    PF_CSWAP(7,12) PF_CSWAP(10,18) PF_SORT3(10,12,20)
    return v[12];
}

// ---- the PixFlow error function, CPU/PixFlow.hpp:407-456 -------------------------------------------
// PixFlow presets (CPU/PixFlow.hpp:461-497): both presets share these values
#define PF_SMOOTHNESS_COEF 0.001f
#define PF_VERT_REG_COEF 0.01f
#define PF_HORZ_REG_COEF 0.01f
#define PF_GRAD_STEP 0.5f
#define PF_GRAD_EPS 0.001f
#define PF_ALPHA_THRESHOLD 0.9f

struct ErrCtx {
    const float2* G1;   // (I1x, I1y) interleaved, row-major, stride = w
    int w, h;
    float wm2, hm2;     // float(w) - 2.0f, float(h) - 2.0f
    float fw;           // float(I0.cols)
};

// getPixBilinear32FExtend on both gradient planes at once (same coordinates, CPU/PixFlow.hpp:445-446)
__device__ __forceinline__ float2 bilinear2(const ErrCtx& c, float x, float y) {
    { const float t = (0.0f < x) ? x : 0.0f; x = (t < c.wm2) ? t : c.wm2; }
    { const float t = (0.0f < y) ? y : 0.0f; y = (t < c.hm2) ? t : c.hm2; }
    const int x0 = __float2int_rz(x), y0 = __float2int_rz(y);
    const float xR = fsub(x, (float)x0), yR = fsub(y, (float)y0);
    const float2* p = c.G1 + (size_t)y0 * c.w + x0;
    const float2 f00 = __ldg(p), f10 = __ldg(p + 1), f01 = __ldg(p + c.w), f11 = __ldg(p + c.w + 1);
    float2 r;
    {
        const float a2 = fsub(f10.x, f00.x), a3 = fsub(f01.x, f00.x);
        const float a4 = fsub(fsub(fadd(f00.x, f11.x), f10.x), f01.x);
        r.x = fadd(fadd(fadd(f00.x, fmul(a2, xR)), fmul(a3, yR)), fmul(fmul(a4, xR), yR));
    }
    {
        const float a2 = fsub(f10.y, f00.y), a3 = fsub(f01.y, f00.y);
        const float a4 = fsub(fsub(fadd(f00.y, f11.y), f10.y), f01.y);
        r.y = fadd(fadd(fadd(f00.y, fmul(a2, xR)), fmul(a3, yR)), fmul(fmul(a4, xR), yR));
    }
    return r;
}

// errorFunction(x, y, flowDir): g0 = (I0x, I0y)(y,x), bl = blurredFlow(y,x)
__device__ __forceinline__ float error_function(const ErrCtx& c, int x, int y, float2 g0, float2 bl, float fx, float fy) {
    const float matchX = fadd((float)x, fx), matchY = fadd((float)y, fy);
    const float2 g1 = bilinear2(c, matchX, matchY);
    const float dX = fsub(bl.x, fx), dY = fsub(bl.y, fy);
    const float smoothness = __fsqrt_rn(fadd(fmul(dX, dX), fmul(dY, dY)));
    const float ex = fsub(g0.x, g1.x), ey = fsub(g0.y, g1.y);
    float err = __fsqrt_rn(fadd(fmul(ex, ex), fmul(ey, ey)));
    err = fadd(err, fmul(smoothness, PF_SMOOTHNESS_COEF));
    err = fadd(err, __fdiv_rn(fmul(PF_VERT_REG_COEF, fabsf(fy)), c.fw));
    err = fadd(err, __fdiv_rn(fmul(PF_HORZ_REG_COEF, fabsf(fx)), c.fw));
    return err;
}

}  // namespace pf

// ---- branch-free exactly-rounded division and square root ------------------------------------------------
// The sweep is latency-bound on one warp per scheduler, so the slow-path branches inside __fdiv_rn/__fsqrt_rn
// cost more than their arithmetic.  These versions give the SAME correctly rounded results on a restricted
// input range (checked exhaustively on the GPU by pf_selftest_exact_math / tests/test_gpu_exact_math.py);
// callers track `bad` and redo the work with the IEEE intrinsics when an input leaves the range.
namespace pf {

// Validity ranges of the branch-free sequences below (checked exhaustively by pf_selftest_exact_math):
//   sqrt_exact_fast(a):     a == 0 or 2^-60 <= a < 2^126
//   div_by_const(x, d, rd): x == 0 or 2^-60 <= |x| < 2^100, for d = 0.001f and every integer d in [PF_EXACT_W_MIN, PF_EXACT_W_MAX]
// Level widths outside that interval (inputs narrower than 48 px, or wider than 16384 px after the half-scale) make the
// launchers select the IEEE-intrinsic path for every pixel (exact_div_width_ok).
#define PF_TINY_BITS 0x21800000u      /* float bits of 2^-60 */
#define PF_EXACT_W_MIN 24
#define PF_EXACT_W_MAX 8192
__host__ __device__ __forceinline__ bool exact_div_width_ok(int w) { return w >= PF_EXACT_W_MIN && w <= PF_EXACT_W_MAX; }
__device__ __forceinline__ bool in_sqrt_range(float a) { return a == 0.0f || (a >= 0x1p-60f && a < 0x1p126f); }
__device__ __forceinline__ bool in_div_range(float x) { const float a = fabsf(x); return a == 0.0f || (a >= 0x1p-60f && a < 0x1p100f); }
// key of a non-negative operand for the "tiny but non-zero" test: min over keys < PF_TINY_BITS-1  <=>  some
// operand lies in (0, 2^-60).  (0 maps to 0xffffffff.)
__device__ __forceinline__ unsigned tiny_key(float x_nonneg) { return __float_as_uint(x_nonneg) - 1u; }

// x / d for a loop-invariant divisor d with rd = RN(1/d): q = RN(x*rd), exact remainder by FMA, one correction
// (Markstein).  FMA is used on purpose here: only the final, correctly rounded quotient matters.
__device__ __forceinline__ float div_by_const(float x, float d, float rd) {
    const float q = __fmul_rn(x, rd);
    const float rem = __fmaf_rn(-q, d, x);
    const float r = __fmaf_rn(rem, rd, q);
    return x == 0.0f ? x : r;          // keeps the sign of a zero numerator (d > 0)
}

// the same two sequences on pairs.  div2_by_const needs both numerators to be >= +0 or non-zero (a -0 numerator would come
// back as +0; the callers divide |x|-products and differences of errors, never -0).
__device__ __forceinline__ f2p div2_by_const(f2p x, float d, float rd) {
    const f2p q = pmuls(x, rd);
    const f2p rem = pfma(q, pk(-d, -d), x);          // x - q*d, exact (FMA)
    return pfma(rem, pk(rd, rd), q);
}
// The seed is taken from max(a, 2^-126): for a == 0 the sequence then yields g = 0 * 2^63 = 0, r = 0, s = 0 (instead of 0 * inf),
// with the sign of the zero preserved, so no select on a == 0 is needed; normal inputs are unaffected.
__device__ __forceinline__ f2p sqrt2_exact_fast(float a0, float a1) {
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fmaxf(a0, 0x1p-126f)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fmaxf(a1, 0x1p-126f)));
    const f2p a = pk(a0, a1), y = pk(y0, y1);
    const f2p g = pmul(a, y), h = pmuls(y, 0.5f);
    const f2p r = pfma(pneg(g), g, a);
    return pfma(r, h, g);
}

// sqrt(a) for a in [2^-60, 2^126) or a == 0: rsqrt seed, one Newton step with exact residual
__device__ __forceinline__ float sqrt_exact_fast(float a) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(a, 0x1p-126f)));
    const float g = __fmul_rn(a, y), h = __fmul_rn(y, 0.5f);
    const float r = __fmaf_rn(-g, g, a);
    return __fmaf_rn(r, h, g);         // a == +-0: g = +-0, r = +-0, s = +-0
}

}  // namespace pf
