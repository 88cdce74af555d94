// pf_prep.cuh -- the per-pixel record of the coming sweep (own-flow terms), shared by the fused stencil kernels
// (pf_fused.cu) and the stand-alone prep kernel used by the diagnostic stage entry (pf_sweep.cu).
#pragma once
#include "pf_kernels.cuh"
#include "pf_math.cuh"

namespace pf {

struct PrepArgs {
    const float* alpha0; const float* alpha1;
    const float2* G0; const float2* G1;      // row-major gradients of image 0 (at the pixel) and image 1 (gathered)
    SweepRec* rec;                           // wavefront-packed output
    int R;                                   // rows per sweep warp
    int dir;                                 // +1 forward sweep, -1 backward sweep
};

__device__ __forceinline__ ErrCtx make_err_ctx(const float2* G1, int w, int h) {
    ErrCtx c;
    c.G1 = G1; c.w = w; c.h = h;
    c.wm2 = fsub((float)w, 2.0f); c.hm2 = fsub((float)h, 2.0f); c.fw = (float)w;
    return c;
}

// Record {E(f0), r0.x, r0.y, - | I0x, I0y, blur.x, blur.y} of pixel (x,y) with old flow f and blurred flow bl
// (CPU/PixFlow.hpp:318 currErr, :321 + :364-386 the gradient step taken when no proposal wins); pixels that the sweep
// must not update (alpha <= 0.9, :317) get {-inf, f}.  Stored in the wavefront-packed order of the sweep kernel:
//     rec[((rowblock * nsteps + step) * R + row_in_block)],  nsteps = w + R - 1, step = logical column + row_in_block
__device__ __forceinline__ void emit_record(const PrepArgs& a, const ErrCtx& c, int x, int y, int w, int h, float2 f, float2 bl) {
    const size_t p = (size_t)y * w + x;
    const float2 g0 = a.G0[p];
    float4 A = make_float4(__int_as_float(0xff800000), f.x, f.y, 0.0f);
    if (a.alpha0[p] > PF_ALPHA_THRESHOLD && a.alpha1[p] > PF_ALPHA_THRESHOLD) {
        const float e0 = error_function(c, x, y, g0, bl, f.x, f.y);
        const float ex = error_function(c, x, y, g0, bl, fadd(f.x, PF_GRAD_EPS), fadd(f.y, 0.0f));
        const float ey = error_function(c, x, y, g0, bl, fadd(f.x, 0.0f), fadd(f.y, PF_GRAD_EPS));
        A.x = e0;
        A.y = fsub(f.x, fmul(PF_GRAD_STEP, __fdiv_rn(fsub(ex, e0), PF_GRAD_EPS)));
        A.z = fsub(f.y, fmul(PF_GRAD_STEP, __fdiv_rn(fsub(ey, e0), PF_GRAD_EPS)));
    }
    const int j = a.dir > 0 ? y : h - 1 - y, i = a.dir > 0 ? x : w - 1 - x;
    const int wb = j / a.R, g = j % a.R;
    const size_t idx = ((size_t)wb * (w + a.R - 1) + (i + g)) * a.R + g;
    SweepRec r;
    r.a = A;
    r.b = make_float4(g0.x, g0.y, bl.x, bl.y);
    a.rec[idx] = r;
}

}  // namespace pf
