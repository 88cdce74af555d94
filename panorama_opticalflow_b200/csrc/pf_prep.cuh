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
    int nsteps_pad;                          // wavefront steps per row group in the record layout (sweep_nsteps_pad(w))
    int dir;                                 // +1 forward sweep, -1 backward sweep
    int slow;                                // 1: level width outside the verified range of div_by_const -> IEEE intrinsics everywhere
};

PrepArgs make_prep_args(const float* alpha0, const float* alpha1, const float2* G0, const float2* G1, SweepRec* rec, int w, int dir);

__device__ __forceinline__ ErrCtx make_err_ctx(const float2* G1, int w, int h) {
    ErrCtx c;
    c.G1 = G1; c.w = w; c.h = h;
    c.wm2 = fsub((float)w, 2.0f); c.hm2 = fsub((float)h, 2.0f); c.fw = (float)w;
    return c;
}

// ---- errorFunction at the three probes f, f+(eps,0), f+(0,eps) of one pixel (CPU/PixFlow.hpp:318, :382-383) ----------
// The probes almost always fall into the same bilinear cell, so the four gradient texels are fetched once and only
// re-fetched when a probe crosses a cell boundary; the IEEE divisions and square roots use the branch-free exactly
// rounded sequences of pf_math.cuh, with a per-thread fallback to the intrinsics outside their validity range.
struct BilCell { int x0, y0; float xR, yR; };

__device__ __forceinline__ BilCell bil_cell_rm(const ErrCtx& c, float x, float y) {
    { const float t = (0.0f < x) ? x : 0.0f; x = (t < c.wm2) ? t : c.wm2; }     // getPixBilinear32FExtend, :409-410
    { const float t = (0.0f < y) ? y : 0.0f; y = (t < c.hm2) ? t : c.hm2; }
    BilCell b;
    b.x0 = __float2int_rz(x); b.y0 = __float2int_rz(y);
    b.xR = fsub(x, (float)b.x0); b.yR = fsub(y, (float)b.y0);
    return b;
}

// the four texels of a bilinear cell as the coefficients of getPixBilinear32FExtend (:415-424), both gradient planes packed
// (fp32x2, pf_math.cuh)
struct BilCoef { f2p f00, a2, a3, a4; };

__device__ __forceinline__ BilCoef load_coef_rm(const ErrCtx& c, int x0, int y0) {
    const float2* p = c.G1 + (size_t)y0 * c.w + x0;
    const f2p F00 = pk(__ldg(p)), F10 = pk(__ldg(p + 1)), F01 = pk(__ldg(p + c.w)), F11 = pk(__ldg(p + c.w + 1));
    BilCoef t;
    t.f00 = F00;
    t.a2 = psub(F10, F00); t.a3 = psub(F01, F00);
    t.a4 = psub(psub(padd(F00, F11), F10), F01);
    return t;
}

__device__ __forceinline__ f2p bil_interp(const BilCoef& t, float xR, float yR) {
    return padd(padd(padd(t.f00, pmuls(t.a2, xR)), pmuls(t.a3, yR)), pmuls(pmuls(t.a4, xR), yR));
}

// errorFunction after the gather (CPU/PixFlow.hpp:447-455) at the three probes (fx0, fy0), (fx1, fy0), (fx0, fy2) of one flow
// vector -- E(f), E(f + (eps,0)), E(f + (0,eps)) of :318 / :382-383 -- given image 1's interpolated gradients G1a/b/c at the three
// matched positions.  The probes share a component pairwise, so the smoothness squares and the two regularisers are computed for
// the FOUR distinct components (two packed operations each) instead of six; every value is produced by the same fp32 operation
// on the same operands as in the reference, so the sharing is exact.  (The reference adds 0.0f to the unchanged component of a
// shifted probe; that only turns a -0 into +0, which none of |f|, (blur - f)^2 and the clamped match position can see.)
// `tiny` collects what the range check of the branch-free exact sequences needs (see emit_record).
template <bool SLOW>
__device__ __forceinline__ void err3_from_g1(float fw, float rcp_w, float2 g0, float2 bl, f2p G1a, f2p G1b, f2p G1c,
                                             float fx0, float fy0, float fx1, float fy2, float v[3], unsigned& tiny) {
    const f2p BL = pk(bl), G0 = pk(g0);
    const f2p Da = psub(BL, pk(fx0, fy0)), Db = psub(BL, pk(fx1, fy2));
    const float2 qa = upk(pmul(Da, Da)), qb = upk(pmul(Db, Db));
    const float ss0 = fadd(qa.x, qa.y), ss1 = fadd(qb.x, qa.y), ss2 = fadd(qa.x, qb.y);
    const f2p Ea = psub(G0, G1a), Eb = psub(G0, G1b), Ec = psub(G0, G1c);
    const float2 ea = upk(pmul(Ea, Ea)), eb = upk(pmul(Eb, Eb)), ec = upk(pmul(Ec, Ec));
    const float gs0 = fadd(ea.x, ea.y), gs1 = fadd(eb.x, eb.y), gs2 = fadd(ec.x, ec.y);
    const float ty0 = fmul(PF_VERT_REG_COEF, fabsf(fy0)), ty2 = fmul(PF_VERT_REG_COEF, fabsf(fy2));
    const float tx0 = fmul(PF_HORZ_REG_COEF, fabsf(fx0)), tx1 = fmul(PF_HORZ_REG_COEF, fabsf(fx1));
    float2 s0, s1, s2, ra, rb;       // {smoothness, gradient} norms per probe; {ry0, rx0}, {ry2, rx1}
    if (SLOW) {
        s0 = make_float2(__fsqrt_rn(ss0), __fsqrt_rn(gs0));
        s1 = make_float2(__fsqrt_rn(ss1), __fsqrt_rn(gs1));
        s2 = make_float2(__fsqrt_rn(ss2), __fsqrt_rn(gs2));
        ra = make_float2(__fdiv_rn(ty0, fw), __fdiv_rn(tx0, fw));
        rb = make_float2(__fdiv_rn(ty2, fw), __fdiv_rn(tx1, fw));
    } else {
        s0 = upk(sqrt2_exact_fast(ss0, gs0)); s1 = upk(sqrt2_exact_fast(ss1, gs1)); s2 = upk(sqrt2_exact_fast(ss2, gs2));
        ra = upk(div2_by_const(pk(ty0, tx0), fw, rcp_w));                         // ty, tx >= +0
        rb = upk(div2_by_const(pk(ty2, tx1), fw, rcp_w));
        tiny = min(tiny, min(min(min(tiny_key(ss0), tiny_key(gs0)), min(tiny_key(ss1), tiny_key(gs1))), min(tiny_key(ss2), tiny_key(gs2))));
        tiny = min(tiny, min(min(tiny_key(ty0), tiny_key(tx0)), min(tiny_key(ty2), tiny_key(tx1))));
    }
    v[0] = fadd(fadd(fadd(s0.y, fmul(s0.x, PF_SMOOTHNESS_COEF)), ra.x), ra.y);
    v[1] = fadd(fadd(fadd(s1.y, fmul(s1.x, PF_SMOOTHNESS_COEF)), ra.x), rb.y);
    v[2] = fadd(fadd(fadd(s2.y, fmul(s2.x, PF_SMOOTHNESS_COEF)), rb.x), ra.y);
}

// Record {E(f0), r0.x, r0.y, - | I0x, I0y, blur.x, blur.y} of pixel (x,y) with old flow f and blurred flow bl
// (CPU/PixFlow.hpp:318 currErr, :321 + :364-386 the gradient step taken when no proposal wins); pixels that the sweep
// must not update (alpha <= 0.9, :317) get {-inf, f}.  Stored in the wavefront-packed order of the sweep kernel:
//     rec[((rowgroup * nsteps_pad + step) * 16 + row_in_group)],  step = logical column + row_in_group
__device__ __forceinline__ void emit_record(const PrepArgs& a, const ErrCtx& c, int x, int y, int w, int h, float2 f, float2 bl) {
    const size_t p = (size_t)y * w + x;
    const float2 g0 = a.G0[p];
    float4 A = make_float4(__int_as_float(0xff800000), f.x, f.y, 0.0f);
    if (a.alpha0[p] > PF_ALPHA_THRESHOLD && a.alpha1[p] > PF_ALPHA_THRESHOLD) {
        const float xf = (float)x, yf = (float)y;
        // (the reference's f.y + 0.0f / f.x + 0.0f of the shifted probes is dropped: see err3_from_g1)
        const float fx1 = fadd(f.x, PF_GRAD_EPS), fy2 = fadd(f.y, PF_GRAD_EPS);
        const BilCell c0 = bil_cell_rm(c, fadd(xf, f.x), fadd(yf, f.y));
        const BilCell c1 = bil_cell_rm(c, fadd(xf, fx1), fadd(yf, f.y));
        const BilCell c2 = bil_cell_rm(c, fadd(xf, f.x), fadd(yf, fy2));
        // every probe gathers its own bilinear cell (no divergent re-gather when a probe crosses a cell boundary, which is common:
        // gradient descent parks many pixels within eps of one)
        const BilCoef t0 = load_coef_rm(c, c0.x0, c0.y0), t1 = load_coef_rm(c, c1.x0, c1.y0), t2 = load_coef_rm(c, c2.x0, c2.y0);
        const f2p g1a = bil_interp(t0, c0.xR, c0.yR), g1b = bil_interp(t1, c1.xR, c1.yR), g1c = bil_interp(t2, c2.xR, c2.yR);
        const float rcp_w = __frcp_rn(c.fw), rcp_eps = __frcp_rn(PF_GRAD_EPS);
        unsigned tiny = 0xffffffffu;
        float ev[3];
        err3_from_g1<false>(c.fw, rcp_w, g0, bl, g1a, g1b, g1c, f.x, f.y, fx1, fy2, ev, tiny);
        float e0 = ev[0], ex = ev[1], ey = ev[2];
        const float2 d = upk(psub(pk(ex, ey), pk(e0, e0)));
        float2 q = upk(div2_by_const(pk(d.x, d.y), PF_GRAD_EPS, rcp_eps));            // a difference of errors is never -0
        tiny = min(tiny, min(tiny_key(fabsf(d.x)), tiny_key(fabsf(d.y))));
        // operands outside the verified range of the branch-free sequences: tiny non-zero (key test), or huge / inf / NaN
        // (every operand is bounded by the errors, so testing those is enough -- same test as the sweep kernel's)
        const float vmax = fmaxf(fmaxf(fabsf(e0), fabsf(ex)), fabsf(ey));
        const bool bad = a.slow || (tiny < PF_TINY_BITS - 1u) || !(vmax < 0x1p50f) || !(e0 == e0) || !(ex == ex) || !(ey == ey);
        if (bad) {                                   // rare: IEEE intrinsics
            unsigned dummy = 0;
            err3_from_g1<true>(c.fw, rcp_w, g0, bl, g1a, g1b, g1c, f.x, f.y, fx1, fy2, ev, dummy);
            e0 = ev[0]; ex = ev[1]; ey = ev[2];
            q.x = __fdiv_rn(fsub(ex, e0), PF_GRAD_EPS);
            q.y = __fdiv_rn(fsub(ey, e0), PF_GRAD_EPS);
        }
        A.x = e0;
        A.y = fsub(f.x, fmul(PF_GRAD_STEP, q.x));
        A.z = fsub(f.y, fmul(PF_GRAD_STEP, q.y));
    }
    const int j = a.dir > 0 ? y : h - 1 - y, i = a.dir > 0 ? x : w - 1 - x;
    const int wb = j >> 4, g = j & 15;                          // sweep warp (16 logical rows) and row within it
    const size_t idx = (((size_t)wb * a.nsteps_pad + (i + g)) << 4) + g;
    SweepRec r;
    r.a = A;
    r.b = make_float4(g0.x, g0.y, bl.x, bl.y);
    a.rec[idx] = r;
}

}  // namespace pf
