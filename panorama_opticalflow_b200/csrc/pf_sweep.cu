// pf_sweep.cu -- the Gauss-Seidel sweeps of PixFlow (CPU/PixFlow.hpp:315-337) as an exact anti-diagonal
// wavefront, laid out for the B200 memory system.
//
// Why a wavefront is exact: the reference visits pixels in (reverse) raster order and updates the flow in
// place; a pixel reads only its already-updated left/up (forward) or right/down (backward) neighbour from the
// array being written.  Any order that respects those two dependencies gives bit-identical results, and the
// anti-diagonal order is the one with the shortest critical path (w+h-1 steps).
//
// Data layouts:
//   * gradients of image 1 ("skewed", anti-diagonal-major): element (x,y) at (x+y)*pitch + pos, pos = x if w<=h else y.
//     All pixels a warp touches in one wavefront step sit on one anti-diagonal, so the bilinear gathers of a step are
//     coalesced (1-2 cache lines per tap) as long as neighbouring rows have similar flow, instead of one line per row.
//   * per-pixel records ("wavefront-packed", written by the prep code fused into the stencil kernels, pf_prep.cuh):
//     one contiguous run of R*32 bytes per warp-step, consumed strictly sequentially -> staged through a shared-memory
//     ring with cp.async several steps ahead, so the step never waits on L2/HBM.
//
// Work split per pixel (i,j):
//   * everything that depends only on the pixel's OWN old flow f0 -- E(f0), E(f0+dx), E(f0+dy) and the result
//     r0 = f0 - step*grad(f0) that is kept when no neighbour proposal wins -- is hoisted out of the dependency chain
//     into the fully parallel prep (record a = {E(f0), r0.x, r0.y}; pixels that must not be updated get {-inf, f0});
//   * the sweep kernel only evaluates the two neighbour candidates -- left (the row's own previous result) and up
//     (previous result of the row above, one shuffle away) -- speculatively at the three probe offsets (0,0), (eps,0),
//     (0,eps) each, finishes BOTH candidates' gradient steps and then does the reference's two compares.
//     P lanes per row (template): P = 2 (default) one candidate per lane -- its three probes each gather their own
//     bilinear cell (12 independent loads, almost always the same L1 lines), the lane finishes its candidate's gradient
//     step and the pair swaps {E, r}; P = 8 one probe per lane (six shuffles collect the errors); P = 4 two probes per
//     lane; P = 1 both candidates on one lane.
//   * the (x, y) channel pairs of gradients and flows go through Blackwell's packed fp32x2 pipe (FFMA2 / FADD2,
//     pf_math.cuh): bit-identical to the scalar operations, half the instructions.
//   * the step body is branch-free: the IEEE divisions (by eps and by cols, both loop-invariant) and square roots
//     use exactly-rounded branchless sequences (pf_math.cuh, verified exhaustively on the GPU); a warp-uniform
//     vote redoes the step with the IEEE intrinsics in the rare case an operand leaves their validity range.
//   * 32/P rows per warp; warps hand the last row's results to the next warp through LL-style lines
//     {fx, flag, fy, flag} (16-byte single-instruction stores, each 8-byte half self-validating, no fences, so L1 is
//     never invalidated): 64-entry shared-memory rings inside a CTA (flag = lap number, back-pressure through a
//     progress counter), full-width arrays in global memory between CTAs, read by a dedicated "poller" warp that
//     forwards them into ring 0 so that no compute warp ever waits on an L2 round trip.
//   * persistent CTAs: the grid is only as wide as the wavefront (front + margin); a CTA takes the next row-block
//     ticket when it finishes one.  Tickets are handed out in row-block order, so the block a CTA waits on is always
//     being processed or done (no deadlock whatever the residency).  Warps queued behind the front sleep-poll.
//   * measured (profiles/r1_sweep_v8_ncu.md): the step is bound by the in-order dependent instruction chain of one warp
//     (~1100 cycles for 360 instructions at P = 2: dependency waits 37 %, issue 33 %, L1 misses of the gather 9 %), not by
//     memory; throughput comes from interleaving independent wavefronts (both directions, several pairs) on the same
//     schedulers -- 23 KB of shared memory and 126 registers per thread keep three sweep CTAs resident per SM.
#include <cstdlib>
#include <type_traits>

#include "pf_kernels.cuh"
#include "pf_math.cuh"
#include "pf_prep.cuh"

namespace pf {

// ---------------------------------------------------------------------------------------------------------
// skewed layout helpers
// ---------------------------------------------------------------------------------------------------------
Skew make_skew(int w, int h) {
    Skew s;
    s.w = w; s.h = h;
    s.posx = (w <= h) ? 1 : 0;
    const int m = w <= h ? w : h;
    s.pitch = (m + 7) & ~7;
    return s;
}
size_t skew_elems(const Skew& s) { return (size_t)(s.w + s.h - 1) * (size_t)s.pitch; }

__device__ __forceinline__ size_t skew_idx(const Skew& s, int x, int y) {
    return (size_t)(x + y) * (size_t)s.pitch + (size_t)(s.posx ? x : y);
}

// Writes a 32x32 row-major tile held in shared memory (tile[ly*32+lx]) to the skewed array: one warp per
// anti-diagonal of the tile, lanes along the diagonal -> contiguous global stores, conflict-free smem reads.
template <class T>
__device__ __forceinline__ void store_tile_skewed(const T* tile, T* __restrict__ out, const Skew& s, int x0, int y0,
                                                  int warp, int nwarps, int lane) {
    for (int ld = warp; ld < 63; ld += nwarps) {
        const int lx = lane, ly = ld - lane;
        if (ly >= 0 && ly < 32 && x0 + lx < s.w && y0 + ly < s.h)
            out[skew_idx(s, x0 + lx, y0 + ly)] = tile[ly * 32 + lx];
    }
}

__global__ void __launch_bounds__(256)
k_skew_copy_f2(const float2* __restrict__ src, float2* __restrict__ dst, Skew s) {
    __shared__ float2 tile[32 * 32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int lx = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ly = threadIdx.y + 8 * k;
        const int x = x0 + lx, y = y0 + ly;
        if (x < s.w && y < s.h) tile[ly * 32 + lx] = src[(size_t)y * s.w + x];
    }
    __syncthreads();
    store_tile_skewed(tile, dst, s, x0, y0, threadIdx.y, 8, lx);
}

void launch_skew_copy_f2(const float2* src, float2* dst, const Skew& s, cudaStream_t st) {
    dim3 b(32, 8), g((s.w + 31) / 32, (s.h + 31) / 32);
    k_skew_copy_f2<<<g, b, 0, st>>>(src, dst, s);
}

// ---------------------------------------------------------------------------------------------------------
// sweep prep: everything that depends only on the pixel's own old flow (fully parallel).
//
// Output layout ("wavefront-packed"): the sweep gives R consecutive logical rows to a warp and at step s row g
// of the warp handles logical column s-g.  Record (A,B) of that pixel is stored at
//     rec[((wb * nsteps + s) * R + g)]            wb = logical row / R, nsteps = w + R - 1
// (R = rows per warp: 4, 16 or 32) so everything one warp needs for one step is ONE contiguous run of R*32 bytes
// and a warp consumes its runs strictly sequentially -- which is what lets the sweep stage them through shared
// memory with cp.async several steps ahead.
// ---------------------------------------------------------------------------------------------------------
size_t sweep_rec_count(int h, int w) {
    // worst case over the supported rows-per-warp values (4, 16, 32), plus the cp.async lookahead
    size_t best = 0;
    for (int R = 4; R <= 32; R *= 2) {
        const size_t nblk = (size_t)(h + R - 1) / R;
        const size_t n = (nblk * (size_t)(w + R - 1) + 16) * R;
        if (n > best) best = n;
    }
    return best;
}

// stand-alone form (the production path fuses this into the blur / median kernels, pf_fused.cu)
__global__ void __launch_bounds__(256)
k_sweep_prep(const float2* __restrict__ blurred, const float2* __restrict__ flow, int h, int w, PrepArgs pa) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const ErrCtx c = make_err_ctx(pa.G1, w, h);
    const size_t p = (size_t)y * w + x;
    emit_record(pa, c, x, y, w, h, flow[p], blurred[p]);
}

void launch_sweep_prep(const float* alpha0, const float* alpha1, const float2* G0, const float2* G1,
                       const float2* blurred, const float2* flow, SweepRec* rec, int h, int w, int dir, cudaStream_t st) {
    dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
    PrepArgs pa;
    pa.alpha0 = alpha0; pa.alpha1 = alpha1; pa.G0 = G0; pa.G1 = G1; pa.rec = rec;
    pa.R = 32 / sweep_lanes_per_row();
    pa.logR = pa.R == 4 ? 2 : (pa.R == 8 ? 3 : (pa.R == 16 ? 4 : 5));
    pa.dir = dir;
    pa.slow = exact_div_width_ok(w) ? 0 : 1;
    k_sweep_prep<<<g, b, 0, st>>>(blurred, flow, h, w, pa);
}

// ---------------------------------------------------------------------------------------------------------
// the wavefront sweep
// ---------------------------------------------------------------------------------------------------------
#ifndef PF_SWEEP_WARPS2
#define PF_SWEEP_WARPS2 4
#endif
#ifndef PF_SW_PREFETCH
#define PF_SW_PREFETCH 4
#endif
constexpr int SW_PREFETCH_GATHER = PF_SW_PREFETCH;          // steps ahead for the L1 warm-up of the gradient gather
constexpr int SW_LL_RING = 64;                 // entries of a shared-memory LL ring (power of two)
constexpr int SW_PROGRESS_EVERY = 8;           // consumer publishes its progress every 8 columns

// Geometry of the sweep kernel for P lanes per row.  P = 8: one error evaluation per lane (shortest chain per
// step); P = 2: one candidate (3 probes) per lane; P = 1: both candidates (6 probes) per lane -- fewer issue slots
// per pixel and independent chains for the in-order scheduler to interleave.
template <int P> struct SweepGeom {
    static constexpr int ROWS = 32 / P;                       // rows per warp
    static constexpr int NQ = P == 8 ? 1 : (P == 4 ? 2 : (P == 2 ? 3 : 6));   // evaluations per lane
    static constexpr int WARPS = P == 8 ? 8 : (P == 4 ? 8 : (P == 2 ? PF_SWEEP_WARPS2 : 2));    // compute warps per CTA (+ 1 poller warp)
    static constexpr int ROWS_PER_CTA = ROWS * WARPS;
    static constexpr int THREADS = (WARPS + 1) * 32;
    static constexpr int DEPTH = P == 8 ? 8 : 4;              // cp.async groups in flight (steps of lookahead)
    static constexpr int SLOTS = 2 * DEPTH;                   // ring slots
    static constexpr int CHUNKS = ROWS * 2;                   // 16-byte chunks per step
};

__device__ __forceinline__ uint4 ll_load_global(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void ll_store_global(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ll_load_shared(unsigned saddr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void ll_store_shared(unsigned saddr, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// predicated forms (no branch around the store: the step body is latency-bound on a single warp)
__device__ __forceinline__ void ll_store_global_if(bool p, uint4* ptr, uint4 v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4}; }"
                 :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void ll_store_shared_if(bool p, unsigned saddr, uint4 v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4}; }"
                 :: "r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void st_volatile_shared_s32_if(bool p, unsigned saddr, int v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.volatile.shared.s32 [%0], %1; }" :: "r"(saddr), "r"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ int ld_volatile_shared_s32(unsigned saddr) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void st_volatile_shared_s32(unsigned saddr, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct SweepConst {
    const float2* G1s;
    unsigned touch;          // this lane's 16-byte scratch slot in shared memory (target of the L1 warm-up copies)
    int g1s_last, dstep, pitch;
    float wm2, hm2, fw, rcp_w, rcp_eps;
};

// The part of errorFunction (CPU/PixFlow.hpp:447-455) after the bilinear gather, with the (x, y) channel pairs on the packed
// fp32x2 pipe (pf_math.cuh): G1 = (I1x, I1y) at the matched position.  Same operations, same order, same roundings as the
// scalar form.  accumulate: fold the operand keys into `tiny` (several probes on one lane) instead of overwriting it.
template <bool SLOW>
__device__ __forceinline__ float err_tail(const SweepConst& k, f2p G1, float2 g0, float2 bl, float fx, float fy, unsigned& tiny, bool accumulate) {
    const f2p D = psub(pk(bl), pk(fx, fy));
    const float2 d2 = upk(pmul(D, D));
    const float ss = fadd(d2.x, d2.y);
    const f2p E = psub(pk(g0), G1);
    const float2 e2 = upk(pmul(E, E));
    const float gs = fadd(e2.x, e2.y);
    const float ty = fmul(PF_VERT_REG_COEF, fabsf(fy)), tx = fmul(PF_HORZ_REG_COEF, fabsf(fx));
    float smooth, grad, ry, rx;
    if (SLOW) {
        smooth = __fsqrt_rn(ss); grad = __fsqrt_rn(gs);
        ry = __fdiv_rn(ty, k.fw); rx = __fdiv_rn(tx, k.fw);
    } else {
        const float2 sq = upk(sqrt2_exact_fast(ss, gs));
        smooth = sq.x; grad = sq.y;
        const float2 rr = upk(div2_by_const(pk(ty, tx), k.fw, k.rcp_w));      // ty, tx >= +0
        ry = rr.x; rx = rr.y;
        const unsigned key = min(min(tiny_key(ss), tiny_key(gs)), min(tiny_key(ty), tiny_key(tx)));
        tiny = accumulate ? min(tiny, key) : key;
    }
    float err = fadd(grad, fmul(smooth, PF_SMOOTHNESS_COEF));
    err = fadd(err, ry);
    err = fadd(err, rx);
    return err;
}

// errorFunction (CPU/PixFlow.hpp:427-456) for ONE flow candidate, gathering I1's gradients from the skewed layout.
// SLOW = false: branch-free exact sequences; `tiny` collects the keys of their operands (see tiny_key).
template <int POSX, bool SLOW>
__device__ __forceinline__ float eval_err(const SweepConst& k, float xf, float yf, float2 g0, float2 bl, float fx, float fy, unsigned& tiny) {
    // getPixBilinear32FExtend, :407-425.  fmaxf/fminf == the std::max/std::min of the reference (NaN -> 0 included)
    const float mx = fminf(fmaxf(fadd(xf, fx), 0.0f), k.wm2);
    const float my = fminf(fmaxf(fadd(yf, fy), 0.0f), k.hm2);
    const int x0 = __float2int_rz(mx), y0 = __float2int_rz(my);
    const float xR = fsub(mx, truncf(mx)), yR = fsub(my, truncf(my));
    const int gi = (x0 + y0) * k.pitch + (POSX ? x0 : y0);
    const float2* p00 = k.G1s + gi;
    const float2* p1 = p00 + k.pitch;            // anti-diagonal +1: (x0,y0+1) and (x0+1,y0) are adjacent
    const float2* p2 = p1 + k.pitch;             // anti-diagonal +2
    const float2 f00 = __ldg(p00);
    const float2 f10 = __ldg(p1 + (POSX ? 1 : 0));
    const float2 f01 = __ldg(p1 + (POSX ? 0 : 1));
    const float2 f11 = __ldg(p2 + 1);
    {   // warm L1 with the anti-diagonal the gather reaches a few steps from now: an asynchronous 16-byte
        // cp.async.ca into a scratch slot allocates the line in L1 and never blocks (its data is not used)
        int pi = gi + 2 * k.pitch + 1 + SW_PREFETCH_GATHER * k.dstep;
        pi = max(0, min(pi, k.g1s_last - 1)) & ~1;
        cp_async16(k.touch, k.G1s + pi);
    }
    const f2p F00 = pk(f00), F10 = pk(f10), F01 = pk(f01), F11 = pk(f11);
    const f2p A2 = psub(F10, F00), A3 = psub(F01, F00), A4 = psub(psub(padd(F00, F11), F10), F01);
    const f2p G1 = padd(padd(padd(F00, pmuls(A2, xR)), pmuls(A3, yR)), pmuls(pmuls(A4, xR), yR));
    return err_tail<SLOW>(k, G1, g0, bl, fx, fy, tiny, false);
}

// The three probes f, f+(eps,0), f+(0,eps) of ONE candidate on one lane (2 or 1 lanes per row).  The probes almost always
// fall into the same bilinear cell: the four texels are gathered once and re-gathered (warp-uniform branch) only when
// some lane's probe crosses a cell boundary.
struct SkewCell { int gi; float xR, yR; };

template <int POSX>
__device__ __forceinline__ SkewCell skew_cell(const SweepConst& k, float mxr, float myr) {
    const float mx = fminf(fmaxf(mxr, 0.0f), k.wm2), my = fminf(fmaxf(myr, 0.0f), k.hm2);
    const int x0 = __float2int_rz(mx), y0 = __float2int_rz(my);
    SkewCell c;
    c.xR = fsub(mx, truncf(mx)); c.yR = fsub(my, truncf(my));
    c.gi = (x0 + y0) * k.pitch + (POSX ? x0 : y0);
    return c;
}

// the four texels of a bilinear cell as the coefficients of getPixBilinear32FExtend (:415-424), both planes packed
struct SkewCoef { f2p f00, a2, a3, a4; };

template <int POSX>
__device__ __forceinline__ SkewCoef skew_gather(const SweepConst& k, int gi) {
    const int i1 = gi + k.pitch;                 // anti-diagonal +1: (x0,y0+1) and (x0+1,y0) are adjacent
    const int i2 = i1 + k.pitch;                 // anti-diagonal +2
    const f2p F00 = pk(__ldg(k.G1s + gi)), F10 = pk(__ldg(k.G1s + i1 + (POSX ? 1 : 0))), F01 = pk(__ldg(k.G1s + i1 + (POSX ? 0 : 1)));
    const f2p F11 = pk(__ldg(k.G1s + i2 + 1));
    SkewCoef c;
    c.f00 = F00;
    c.a2 = psub(F10, F00); c.a3 = psub(F01, F00);
    c.a4 = psub(psub(padd(F00, F11), F10), F01);
    return c;
}

template <bool SLOW>
__device__ __forceinline__ float err_from_taps(const SweepConst& k, const SkewCoef& t, float xR, float yR, float2 g0, float2 bl,
                                               float fx, float fy, unsigned& tiny) {
    const f2p G1 = padd(padd(padd(t.f00, pmuls(t.a2, xR)), pmuls(t.a3, yR)), pmuls(pmuls(t.a4, xR), yR));
    return err_tail<SLOW>(k, G1, g0, bl, fx, fy, tiny, true);
}

// The three probes f, f+(eps,0), f+(0,eps) of ONE candidate on one lane.  Gradient descent on the piecewise-bilinear error
// parks many pixels within eps of a bilinear-cell boundary, so in most warp-steps some probe falls into the neighbouring
// cell (73 % in profiles/r1_sweep_v8_ncu.md).  Every probe therefore gathers its own cell unconditionally: 12 loads issued
// together -- almost always the same one or two L1 lines -- instead of 4 loads, a warp vote, a branch and a second, dependent
// gather on the step's critical path.
template <int POSX, bool SLOW>
__device__ __forceinline__ void eval_err3(const SweepConst& k, float xf, float yf, float2 g0, float2 bl, float2 cand,
                                          float v[3], unsigned& tiny) {
    // flow + Point2f(eps, 0) / (0, eps) of the reference add +0 to the other component; a -0 component becomes +0 there, which
    // changes nothing in errorFunction (x + (-0) == x + 0, |-0| == 0, b - (-0) == b - 0), so the candidate is used as it is
    const float fx0 = cand.x, fy0 = cand.y;
    const float fx1 = fadd(cand.x, PF_GRAD_EPS), fy2 = fadd(cand.y, PF_GRAD_EPS);
    const SkewCell c0 = skew_cell<POSX>(k, fadd(xf, fx0), fadd(yf, fy0));
    const SkewCell c1 = skew_cell<POSX>(k, fadd(xf, fx1), fadd(yf, fy0));
    const SkewCell c2 = skew_cell<POSX>(k, fadd(xf, fx0), fadd(yf, fy2));
    const SkewCoef t0 = skew_gather<POSX>(k, c0.gi), t1 = skew_gather<POSX>(k, c1.gi), t2 = skew_gather<POSX>(k, c2.gi);
    {   // warm L1 with the anti-diagonal the gather reaches a few steps from now (asynchronous copy into a scratch slot)
        int pi = c0.gi + 2 * k.pitch + 1 + SW_PREFETCH_GATHER * k.dstep;
        pi = max(0, min(pi, k.g1s_last - 1)) & ~1;
        cp_async16(k.touch, k.G1s + pi);
    }
    v[0] = err_from_taps<SLOW>(k, t0, c0.xR, c0.yR, g0, bl, fx0, fy0, tiny);
    v[1] = err_from_taps<SLOW>(k, t1, c1.xR, c1.yR, g0, bl, fx1, fy0, tiny);
    v[2] = err_from_taps<SLOW>(k, t2, c2.xR, c2.yR, g0, bl, fx0, fy2, tiny);
}

// P = 4: two probes of ONE candidate on one lane -- (0,0) and (eps,0) on the even lane of the candidate's lane pair, (0,eps)
// twice on the odd one (the duplicate is free in SIMT and keeps the lanes in step)
template <int POSX, bool SLOW>
__device__ __forceinline__ void eval_err2(const SweepConst& k, float xf, float yf, float2 g0, float2 bl, float2 cand, bool odd,
                                          float v[2], unsigned& tiny) {
    const float fxa = fadd(cand.x, 0.0f), fya = fadd(cand.y, odd ? PF_GRAD_EPS : 0.0f);
    const float fxb = fadd(cand.x, odd ? 0.0f : PF_GRAD_EPS), fyb = fya;
    const SkewCell c0 = skew_cell<POSX>(k, fadd(xf, fxa), fadd(yf, fya));
    const SkewCell c1 = skew_cell<POSX>(k, fadd(xf, fxb), fadd(yf, fyb));
    const SkewCoef t0 = skew_gather<POSX>(k, c0.gi);
    {   // warm L1 with the anti-diagonal the gather reaches a few steps from now (asynchronous copy into a scratch slot)
        int pi = c0.gi + 2 * k.pitch + 1 + SW_PREFETCH_GATHER * k.dstep;
        pi = max(0, min(pi, k.g1s_last - 1)) & ~1;
        cp_async16(k.touch, k.G1s + pi);
    }
    const SkewCoef t1 = skew_gather<POSX>(k, c1.gi);
    v[0] = err_from_taps<SLOW>(k, t0, c0.xR, c0.yR, g0, bl, fxa, fya, tiny);
    v[1] = err_from_taps<SLOW>(k, t1, c1.xR, c1.yR, g0, bl, fxb, fyb, tiny);
}

// One candidate's gradient step (CPU/PixFlow.hpp:321, :364-386) from its three errors {E, E(+dx), E(+dy)}: r = cand - step * dE/eps
template <bool SLOW>
__device__ __forceinline__ float2 finish_candidate(const SweepConst& k, const float e3[3], float2 cand, unsigned& tiny) {
    const float2 d = upk(psub(pk(e3[1], e3[2]), pk(e3[0], e3[0])));
    f2p Q;
    if (SLOW) {
        Q = pk(__fdiv_rn(d.x, PF_GRAD_EPS), __fdiv_rn(d.y, PF_GRAD_EPS));
    } else {
        Q = div2_by_const(pk(d.x, d.y), PF_GRAD_EPS, k.rcp_eps);              // a difference of errors is never -0
        tiny = min(tiny_key(fabsf(d.x)), tiny_key(fabsf(d.y)));
    }
    return upk(psub(pk(cand), pmuls(Q, PF_GRAD_STEP)));
}

// From the six errors {L, L+dx, L+dy, U, U+dx, U+dy} of a pixel to its result: finish both candidates' gradient steps
// (CPU/PixFlow.hpp:321, :364-386) and select in the reference's order (:318-320: left proposal first, then up, strict <).
// the reference's two compares (:318-320: left proposal first, then up, strict <) between the pixel's own record A = {E(f0), r0}
// and the finished candidates
__device__ __forceinline__ float2 select_result(float eL, float2 rL, float eU, float2 rU, bool leftValid, bool upValid, float4 A) {
    const float POS_INF = __int_as_float(0x7f800000);
    eL = leftValid ? eL : POS_INF;
    eU = upValid ? eU : POS_INF;
    // decide from the errors alone, pick the vectors last: the candidates' r arrive (shuffles) after their E
    const bool pL = eL < A.x;
    const float cur = pL ? eL : A.x;
    const bool pU = eU < cur;
    float2 out = make_float2(A.y, A.z);
    out = pL ? rL : out;
    out = pU ? rU : out;
    return out;
}

// From the six errors {L, L+dx, L+dy, U, U+dx, U+dy} of a pixel to its result: finish both candidates' gradient steps, then select.
template <bool SLOW>
__device__ __forceinline__ float2 finish_pixel(const SweepConst& k, const float e6[6], float2 left, float2 up,
                                               bool leftValid, bool upValid, float4 A, unsigned& tiny) {
    unsigned t1 = 0xffffffffu, t2 = 0xffffffffu;
    const float2 rL = finish_candidate<SLOW>(k, e6, left, t1);
    const float2 rU = finish_candidate<SLOW>(k, e6 + 3, up, t2);
    tiny = min(t1, t2);
    return select_result(e6[0], rL, e6[3], rU, leftValid, upValid, A);
}

// One step of one lane: evaluate this lane's probes, finish the candidates, select (the body of the wavefront loop).
// SLOW = false: branch-free exact sequences, tkey / vmax collect what their validity check needs; SLOW = true: IEEE intrinsics.
template <int POSX, int P, bool SLOW>
__device__ __forceinline__ float2 step_eval(const SweepConst& k, float xf, float yf, float4 A, float4 B, float2 res, float2 up,
                                            int i, int j, int sub, int gbase, unsigned& tkey, float& vmax) {
    typedef SweepGeom<P> G;
    const unsigned full = 0xffffffffu;
    const float2 g0 = make_float2(B.x, B.y), bl = make_float2(B.z, B.w);
    float2 out = make_float2(A.y, A.z);
    float v[G::NQ];
    // P = 8: this lane's single probe (lanes 0-2 left candidate, 3-5 up candidate, 6-7 duplicates of 3-4)
    const bool candUp8 = sub >= 3;
    const float offx8 = (sub % 3) == 1 ? PF_GRAD_EPS : 0.0f, offy8 = (sub % 3) == 2 ? PF_GRAD_EPS : 0.0f;
    (void)candUp8; (void)offx8; (void)offy8; (void)gbase;
    if constexpr (P == 8) {
        const float2 cand = candUp8 ? up : res;
        unsigned t1 = 0xffffffffu;
        v[0] = eval_err<POSX, SLOW>(k, xf, yf, g0, bl, fadd(cand.x, offx8), fadd(cand.y, offy8), t1);
        tkey = min(tkey, t1);
    } else if constexpr (P == 4) {
        eval_err2<POSX, SLOW>(k, xf, yf, g0, bl, sub >= 2 ? up : res, (sub & 1) != 0, v, tkey);
    } else if constexpr (P == 2) {
        eval_err3<POSX, SLOW>(k, xf, yf, g0, bl, sub != 0 ? up : res, v, tkey);
    } else {
        eval_err3<POSX, SLOW>(k, xf, yf, g0, bl, res, v, tkey);
        eval_err3<POSX, SLOW>(k, xf, yf, g0, bl, up, v + 3, tkey);
    }
#pragma unroll
    for (int q = 0; q < G::NQ; ++q) {
        vmax = fmaxf(vmax, fabsf(v[q]));
        if (!(v[q] == v[q])) vmax = __int_as_float(0x7f800000);      // NaN -> flagged
    }
    unsigned t2 = 0xffffffffu;
    if constexpr (P == 2) {
        // each lane finishes ITS candidate's gradient step, then the pair swaps {E, r.x, r.y}
        const float oe = __shfl_xor_sync(full, v[0], 1);            // E first: it is ready before the gradient step
        const float2 mine = finish_candidate<SLOW>(k, v, sub != 0 ? up : res, t2);
        const float ox = __shfl_xor_sync(full, mine.x, 1), oy = __shfl_xor_sync(full, mine.y, 1);
        const float2 other = make_float2(ox, oy);
        out = select_result(sub == 0 ? v[0] : oe, sub == 0 ? mine : other, sub == 0 ? oe : v[0], sub == 0 ? other : mine,
                            i > 0, j > 0, A);
    } else if constexpr (P == 4) {
        // lanes {0,1} of a row hold the left candidate's {E, E+dx | E+dy}, lanes {2,3} the up candidate's: the even
        // lane of each pair finishes its candidate, then all four lanes fetch both {E, r.x, r.y}
        const float edy = __shfl_xor_sync(full, v[0], 1);
        const float e3[3] = {v[0], v[1], edy};
        const float2 mine = finish_candidate<SLOW>(k, e3, sub >= 2 ? up : res, t2);
        t2 = (sub & 1) ? 0xffffffffu : t2;               // the odd lanes' finish is a don't-care
        const float eL = __shfl_sync(full, v[0], gbase), eU = __shfl_sync(full, v[0], gbase + 2);
        const float2 rL = make_float2(__shfl_sync(full, mine.x, gbase), __shfl_sync(full, mine.y, gbase));
        const float2 rU = make_float2(__shfl_sync(full, mine.x, gbase + 2), __shfl_sync(full, mine.y, gbase + 2));
        out = select_result(eL, rL, eU, rU, i > 0, j > 0, A);
    } else {
        // every lane of the row gets the six errors {L, L+dx, L+dy, U, U+dx, U+dy}
        float e6[6];
        if (P == 8) {
#pragma unroll
            for (int q = 0; q < 6; ++q) e6[q] = __shfl_sync(full, v[0], gbase + q);
        } else {
#pragma unroll
            for (int q = 0; q < 6; ++q) e6[q] = v[q % G::NQ];
        }
        out = finish_pixel<SLOW>(k, e6, res, up, i > 0, j > 0, A, t2);
    }
    tkey = min(tkey, t2);
    return out;
}

template <int POSX, int P>
__device__ __noinline__ float2 step_eval_slow(SweepConst k, float xf, float yf, float4 A, float4 B, float2 res, float2 up,
                                              int i, int j, int sub, int gbase) {
    unsigned tkey = 0xffffffffu;
    float vmax = 0.0f;
    return step_eval<POSX, P, true>(k, xf, yf, A, B, res, up, i, j, sub, gbase, tkey, vmax);
}

// Shared memory of one sweep CTA.
template <int P> struct SweepSmem {
    typedef SweepGeom<P> G;
    int b;                                                       // row block being processed
    int progress[G::WARPS];                                      // columns consumed from ring k
    __align__(16) uint4 llring[G::WARPS][SW_LL_RING];            // [0] inbound via the poller, [k] from warp k-1
    __align__(128) SweepRec ring[G::WARPS][G::SLOTS][G::ROWS];   // cp.async staging of the record stream
    uint4 touch[G::WARPS][32];                                   // targets of the L1 warm-up copies
};

// One row block (ROWS_PER_CTA logical rows) of one sweep, by one CTA.
template <int DIR, int POSX, int P>
__device__ __noinline__ void sweep_block(const Sweep2Args& a, SweepSmem<P>& sm, const int b) {
    typedef SweepGeom<P> G;
    const unsigned full = 0xffffffffu;
    int w = a.s.w, h = a.s.h;
    auto& s_progress = sm.progress;
    auto& s_llring = sm.llring;
    auto& s_ring = sm.ring;
    auto& s_touch = sm.touch;
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (wi == G::WARPS) {
        // ---- poller warp: forward the upstream CTA's global LL lines into ring 0 as they become valid ----
        if (b == 0) return;
        const uint4* src = a.boundary + (size_t)(b - 1) * w;
        const unsigned prog = smem_u32(&s_progress[0]);
        const unsigned ring0 = smem_u32(&s_llring[0][0]);
        for (int base = 0; base < w; base += 32) {
            while (base + 32 > ld_volatile_shared_s32(prog) + SW_LL_RING) { }    // back-pressure: batch must fit
            const int col = base + lane;
            bool done = col >= w;
            bool started = base > 0;
            while (!__all_sync(full, done)) {
                bool got = false;
                if (!done) {
                    const uint4 v = ll_load_global(src + col);
                    if (v.y == 1u && v.w == 1u) {
                        const unsigned e = (unsigned)(col / SW_LL_RING) + 1u;
                        ll_store_shared(ring0 + (col & (SW_LL_RING - 1)) * 16, make_uint4(v.x, e, v.z, e));
                        done = true; got = true;
                    }
                }
                if (!started) {                       // upstream CTA not running yet: back off
                    started = __any_sync(full, got);
                    if (!started) __nanosleep(256);
                }
            }
        }
        return;
    }

    const int g = lane / P, sub = lane % P, gbase = lane - sub;
    const int jw = b * G::ROWS_PER_CTA + wi * G::ROWS;            // first logical row of this warp
    if (jw >= h) return;
    const int j = jw + g;
    const bool rowValid = j < h;
    const int y = DIR > 0 ? j : h - 1 - j;
    const bool has_in = jw > 0;                                    // warp-uniform
    const bool has_out = jw + G::ROWS < h;                         // warp-uniform
    const bool out_global = wi == G::WARPS - 1;
    const unsigned rin = smem_u32(&s_llring[wi][0]);
    const unsigned rout = smem_u32(&s_llring[(wi + 1) % G::WARPS][0]);
    const unsigned prog_in = smem_u32(&s_progress[wi]);
    const unsigned prog_out = smem_u32(&s_progress[(wi + 1) % G::WARPS]);
    uint4* gout = a.boundary + (size_t)b * w;
    int out_limit = SW_LL_RING;                                    // columns < out_limit fit in the out ring unchecked

    SweepConst k;
    k.G1s = a.G1s; k.g1s_last = (int)a.g1s_last;
    k.touch = smem_u32(&s_touch[wi][lane]);
    k.pitch = a.s.pitch;
    k.dstep = DIR * (k.pitch + POSX);
    k.wm2 = fsub((float)w, 2.0f); k.hm2 = fsub((float)h, 2.0f); k.fw = (float)w;
    k.rcp_w = __frcp_rn(k.fw); k.rcp_eps = __frcp_rn(PF_GRAD_EPS);
    asm volatile("" : "+r"(w), "+r"(k.pitch), "+r"(k.dstep), "+r"(k.g1s_last));     // keep loop invariants in registers
    asm volatile("" : "+f"(k.wm2), "+f"(k.hm2), "+f"(k.fw), "+f"(k.rcp_w), "+f"(k.rcp_eps));
    const bool force_slow = !exact_div_width_ok(w);      // level width outside the verified range of div_by_const
    const float yf = (float)y;
    float xf = (float)(DIR > 0 ? -g : w - 1 + g);                  // float(x) of step 0, then +-1 per step (exact)
    float2* flow_row = a.flow + (size_t)y * w;

    // ---- record stream: ROWS*32 contiguous bytes per step, staged through a shared-memory ring with cp.async ----
    const int nsteps = w + G::ROWS - 1;
    const uint4* stream = reinterpret_cast<const uint4*>(a.rec + (size_t)(jw / G::ROWS) * nsteps * G::ROWS);
    const unsigned ring = smem_u32(&s_ring[wi][0][0]);
    auto issue = [&](int t) {
        const unsigned dst = ring + (t % G::SLOTS) * (G::CHUNKS * 16);
        const uint4* src = stream + (size_t)t * G::CHUNKS;
#pragma unroll
        for (int c0 = 0; c0 < G::CHUNKS; c0 += 32)
            if (c0 + lane < G::CHUNKS) cp_async16(dst + (c0 + lane) * 16, src + c0 + lane);
        cp_async_commit();
    };
    for (int t = 0; t < G::DEPTH; ++t) issue(t);                   // prologue: steps 0 .. depth-1
    const uint4* s_slot0 = reinterpret_cast<const uint4*>(&s_ring[wi][0][g]);

    float2 res = make_float2(0.0f, 0.0f);
    uint4 ln = make_uint4(0u, 0u, 0u, 0u);
    if (has_in) {
        // Waiting for this warp's turn (the wavefront reaches row jw after ~jw steps): sleep-poll so that the
        // warps still queued behind the front leave the issue slots to the warps that are working.
        ln = ll_load_shared(rin);
        while (ln.y != 1u || ln.w != 1u) { __nanosleep(128); ln = ll_load_shared(rin); }
    }
    const int in_cols = has_in ? w : 0;           // columns to take from the inbound ring
    const bool ring_out = has_out && !out_global;
    const bool last_row = g == G::ROWS - 1 && sub == 0;
    // records of step 0 (software pipeline: the records of step s+1 are fetched from the ring at the end of step s)
    cp_async_wait<G::DEPTH - 1>();
    __syncwarp();
    float4 A = *reinterpret_cast<const float4*>(s_slot0);
    float4 B = *reinterpret_cast<const float4*>(s_slot0 + 1);

#pragma unroll 2
    for (int s = 0; s < nsteps; ++s) {
        const int i = s - g;                      // logical column of this row at this step
        issue(s + G::DEPTH);                      // keep `depth` runs in flight
        // ---- up neighbour: previous result of the row above (shuffle; first row of the warp: LL ring) ----
        float2 up;
        up.x = __shfl_up_sync(full, res.x, P);
        up.y = __shfl_up_sync(full, res.y, P);
        if (s < in_cols) {                        // warp-uniform: every lane reads the same ring entry
            const unsigned e = (unsigned)(s / SW_LL_RING) + 1u;
            uint4 v = ln;
            while (v.y != e || v.w != e) v = ll_load_shared(rin + (s & (SW_LL_RING - 1)) * 16);
            ln = ll_load_shared(rin + ((s + 1) & (SW_LL_RING - 1)) * 16);
            st_volatile_shared_s32_if((s & (SW_PROGRESS_EVERY - 1)) == SW_PROGRESS_EVERY - 1 && lane == 0, prog_in, s + 1);
            up.x = g == 0 ? __uint_as_float(v.x) : up.x;
            up.y = g == 0 ? __uint_as_float(v.z) : up.y;
        }
        const bool valid = rowValid && (unsigned)i < (unsigned)w;
        const bool active = valid && __float_as_uint(A.x) != 0xff800000u;     // -inf marks "not updatable"; a NaN E(f0) stays active
        float2 out = make_float2(A.y, A.z);
        if (__any_sync(full, active)) {           // warp-uniform: skip fully inactive stretches
            unsigned tkey = 0xffffffffu;
            float vmax = 0.0f;
            out = step_eval<POSX, P, false>(k, xf, yf, A, B, res, up, i, j, sub, gbase, tkey, vmax);
            // operands left the range of the branch-free sequences (tiny non-zero, or huge / inf / NaN)?
            const bool bad = force_slow || (tkey < PF_TINY_BITS - 1u) || !(vmax < 0x1p50f);
            // rare: redo the step with the IEEE intrinsics -- out of line, so that the hot loop stays compact in the instruction cache
            if (__any_sync(full, bad && active)) out = step_eval_slow<POSX, P>(k, xf, yf, A, B, res, up, i, j, sub, gbase);
        }
        res.x = valid ? out.x : res.x;
        res.y = valid ? out.y : res.y;
        // ---- results: flow (row-major, only where alpha > 0.9) and the hand-off of the warp's last row ----
        const int x = DIR > 0 ? i : w - 1 - i;
        if (active && sub == 0) flow_row[x] = out;
        const int i_last = s - (G::ROWS - 1);                 // column of the warp's last row (warp-uniform)
        if (ring_out && i_last >= out_limit)                  // back-pressure, rare: wait until the slot is free
            while (i_last >= out_limit) out_limit = ld_volatile_shared_s32(prog_out) + SW_LL_RING;
        {
            const unsigned e = out_global ? 1u : (unsigned)(i_last / SW_LL_RING) + 1u;
            const uint4 lv = make_uint4(__float_as_uint(out.x), e, __float_as_uint(out.y), e);
            const bool doit = last_row && valid && has_out;
            ll_store_global_if(doit && out_global, gout + i, lv);
            ll_store_shared_if(doit && !out_global, rout + (i_last & (SW_LL_RING - 1)) * 16, lv);
        }
        xf = fadd(xf, (float)DIR);
        // ---- records of the next step ----
        cp_async_wait<G::DEPTH - 1>();
        __syncwarp();
        const uint4* slot = s_slot0 + ((s + 1) % G::SLOTS) * G::CHUNKS;
        A = *reinterpret_cast<const float4*>(slot);
        B = *reinterpret_cast<const float4*>(slot + 1);
    }
    cp_async_wait<0>();
}

// Persistent sweep kernel: the grid holds only about as many CTAs as the wavefront is wide (front + margin); a CTA
// that finishes its row block takes the next ticket.  Tickets are handed out in row-block order, so the block a CTA
// waits on is always being processed (or done) -- no deadlock whatever the residency -- and CTAs far behind the
// front do not occupy registers and shared memory while they would only be waiting for their turn.
template <int DIR, int POSX, int P>
__global__ void __launch_bounds__(SweepGeom<P>::THREADS)
k_sweep6(Sweep2Args a) {
    typedef SweepGeom<P> G;
    __shared__ SweepSmem<P> sm;
    const int nblocks = (a.s.h + G::ROWS_PER_CTA - 1) / G::ROWS_PER_CTA;
    for (;;) {
        __syncthreads();                                  // every warp is done with the previous block
        if (threadIdx.x == 0) sm.b = atomicAdd(a.ticket, 1);
        for (int i = threadIdx.x; i < G::WARPS * SW_LL_RING; i += G::THREADS) (&sm.llring[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < G::WARPS) sm.progress[threadIdx.x] = 0;
        __syncthreads();
        const int b = sm.b;
        if (b >= nblocks) break;
        sweep_block<DIR, POSX, P>(a, sm, b);
    }
}

// lanes per row of the sweep kernel: 2 (default: measured equal step latency to 8 with a quarter of the warps and
// a third of the issue slots per pixel, hence the best throughput when several pairs are in flight), 8 or 1.
// PF_SWEEP_LANES overrides; read once.
int sweep_lanes_per_row() {
    static int p = 0;
    if (p == 0) {
        const char* e = getenv("PF_SWEEP_LANES");
        const int v = e ? atoi(e) : 2;
        p = (v == 1 || v == 2 || v == 4 || v == 8) ? v : 2;
    }
    return p;
}

static int rows_per_cta() {
    const int p = sweep_lanes_per_row();
    return p == 8 ? SweepGeom<8>::ROWS_PER_CTA : (p == 4 ? SweepGeom<4>::ROWS_PER_CTA : (p == 2 ? SweepGeom<2>::ROWS_PER_CTA : SweepGeom<1>::ROWS_PER_CTA));
}

size_t sweep2_boundary_lines(int h, int w, bool) {
    // sized for the smallest CTA (32 rows), whatever PF_SWEEP_LANES says
    const int ncta = (h + 31) / 32;
    return (size_t)(ncta > 1 ? ncta - 1 : 0) * (size_t)w + 1;
}

bool sweep2_use_smem(int) { return true; }

// Experiment knob: a dummy dynamic shared-memory request caps the sweep CTAs per SM (PF_SWEEP_CTAS_PER_SM=n).
// Measured on B200 with 16-32 pairs in flight: capping at 3 or 2 is 9-16 % SLOWER than no cap, so the default is
// no cap (profiles/r1_batch_scaling.md).
static size_t sweep_smem_pad(size_t static_smem) {
    static int per_sm = -1;
    if (per_sm < 0) {
        const char* e = getenv("PF_SWEEP_CTAS_PER_SM");
        per_sm = e ? atoi(e) : 0;
        if (per_sm < 1 || per_sm > 8) per_sm = 0;     // 0: no cap
    }
    if (per_sm == 0) return 0;
    const size_t budget = (size_t)227 * 1024 / per_sm - 1024;    // per CTA, incl. the 1 KB the driver reserves
    return budget > static_smem + 1024 ? ((budget - static_smem) & ~(size_t)127) : 0;
}

template <int P>
static void launch_sweep_p(const Sweep2Args& a, int dir, cudaStream_t st) {
    typedef SweepGeom<P> G;
    const int nblocks = (a.s.h + G::ROWS_PER_CTA - 1) / G::ROWS_PER_CTA;
    const int front = (a.s.w + 2 * G::ROWS_PER_CTA - 1) / G::ROWS_PER_CTA + 2;   // row blocks working at the same time
    const int ncta = nblocks < front ? nblocks : front;
    static size_t pad = (size_t)-1;
    static bool attr_done[64] = {};
    if (pad == (size_t)-1) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_sweep6<1, 1, P>);
        pad = sweep_smem_pad(fa.sharedSizeBytes);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (pad > 0 && dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaFuncSetAttribute(k_sweep6<1, 1, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        cudaFuncSetAttribute(k_sweep6<1, 0, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        cudaFuncSetAttribute(k_sweep6<-1, 1, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        cudaFuncSetAttribute(k_sweep6<-1, 0, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
        attr_done[dev] = true;
    }
    if (dir > 0) {
        if (a.s.posx) k_sweep6<1, 1, P><<<ncta, G::THREADS, pad, st>>>(a); else k_sweep6<1, 0, P><<<ncta, G::THREADS, pad, st>>>(a);
    } else {
        if (a.s.posx) k_sweep6<-1, 1, P><<<ncta, G::THREADS, pad, st>>>(a); else k_sweep6<-1, 0, P><<<ncta, G::THREADS, pad, st>>>(a);
    }
}

void launch_sweep2(const Sweep2Args& a, int dir, cudaStream_t st) {
    (void)rows_per_cta;
    static int skip = -1;
    if (skip < 0) skip = getenv("PF_EXP_SKIP_SWEEP") ? 1 : 0;    // timing experiment only: results are wrong
    if (skip) return;
    switch (sweep_lanes_per_row()) {
    case 1: launch_sweep_p<1>(a, dir, st); break;
    case 2: launch_sweep_p<2>(a, dir, st); break;
    case 4: launch_sweep_p<4>(a, dir, st); break;
    default: launch_sweep_p<8>(a, dir, st); break;
    }
}

}  // namespace pf
