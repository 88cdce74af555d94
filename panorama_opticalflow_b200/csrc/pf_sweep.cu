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
//     a warp (16 logical rows) consumes ONE contiguous run of 16 x 32 bytes per wavefront step, strictly sequentially ->
//     staged through shared memory by the TMA engine (1-D bulk copies of 4 steps = 2 KB, completion on an mbarrier, issued by
//     one lane every fourth step), two stages ahead, so the step never waits on L2/HBM and spends two shared-memory loads on
//     its records.
//
// Work split per pixel (i,j):
//   * everything that depends only on the pixel's OWN old flow f0 -- E(f0), E(f0+dx), E(f0+dy) and the result
//     r0 = f0 - step*grad(f0) that is kept when no neighbour proposal wins -- is hoisted out of the dependency chain
//     into the fully parallel prep (record a = {E(f0), r0.x, r0.y}; pixels that must not be updated get {-inf, f0});
//   * the sweep kernel evaluates the two neighbour candidates -- left (the row's own previous result) and up (previous
//     result of the row above, one shuffle away) -- speculatively at the three probe offsets (0,0), (eps,0), (0,eps)
//     each, finishes BOTH candidates' gradient steps and then does the reference's two compares.
//
// Lanes: two lanes per row (16 rows per warp).  Lane 0 of a pair evaluates the LEFT candidate's three probes, lane 1 the UP
// candidate's; each finishes its candidate's gradient step and the pair swaps {E, r} with three shuffles.  The step is bound by the
// LATENCY of its dependent chain (ring / shuffle in -> clamp, floor, address -> L1 gather -> bilinear -> exact square roots -> sums ->
// exact division -> exchange -> select), not by its instruction count (profiles/r2_sweep_v12_ncu.md): what pays is a shorter chain and
// fewer L1 misses of the gather, not fewer instructions in the chain's shadow.  Spreading the probes over more lanes (eight lanes per
// row, one probe each) or over several warps was built and measured and is not faster: the parts of the step without any parallelism
// (the way in, the exchange, the select, the votes) are the same (profiles/r2_notes.md, profiles/r2_sweep_v10_experiment.md).
//   * the (x, y) channel pairs of gradients and flows go through Blackwell's packed fp32x2 pipe (FFMA2 / FADD2,
//     pf_math.cuh): bit-identical to the scalar operations, half the instructions.
//   * the step body is branch-free: the IEEE divisions (by eps and by cols, both loop-invariant) and square roots
//     use exactly-rounded branchless sequences (pf_math.cuh, verified exhaustively on the GPU); a warp-uniform
//     vote redoes the evaluation / the finish with the IEEE intrinsics in the rare case an operand leaves their range.
//   * warps hand the last row's results down through LL-style lines {fx, flag, fy, flag} (16-byte single-instruction
//     stores, each 8-byte half self-validating, no fences, so L1 is never invalidated): 64-entry shared-memory rings
//     inside a CTA (flag = lap number, back-pressure through a progress counter), full-width arrays in global
//     memory between CTAs, read by a dedicated "poller" warp that forwards them into ring 0 so that no compute warp ever
//     waits on an L2 round trip.
//   * persistent CTAs: the grid is only as wide as the wavefront (front + margin); a CTA takes the next row-block
//     ticket when it finishes one.  Tickets are handed out in row-block order, so the block a CTA waits on is always
//     being processed or done (no deadlock whatever the residency).  Warps queued behind the front sleep-poll.
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
// Output layout ("wavefront-packed"): a sweep warp owns 16 consecutive logical rows; at step s row g of the warp handles
// logical column s - g.  Record (A,B) of that pixel is stored at
//     rec[((rowgroup * nsteps_pad + s) * 16 + g)]      rowgroup = logical row / 16, nsteps_pad = (w + 15) rounded up to 4
// so everything a warp needs for one step is ONE contiguous run of 512 bytes, four steps are one 2 KB bulk copy, and a
// warp consumes its runs strictly sequentially.
// ---------------------------------------------------------------------------------------------------------
// (rounded up to 4 steps whatever the stage size, so that the record layout does not depend on a tuning knob)
__host__ __device__ __forceinline__ int sweep_nsteps_pad_dev(int w) { return (w + SWEEP_GROUP_ROWS - 1 + 3) & ~3; }
int sweep_nsteps_pad(int w) { return sweep_nsteps_pad_dev(w); }

size_t sweep_rec_count(int h, int w) {
    const size_t ngroups = (size_t)(h + SWEEP_GROUP_ROWS - 1) / SWEEP_GROUP_ROWS;
    return ngroups * (size_t)sweep_nsteps_pad(w) * SWEEP_GROUP_ROWS;
}

// stand-alone form (the production path fuses this into the blur / median kernels, pf_fused.cu)
__global__ void __launch_bounds__(256)
k_sweep_prep(const float2* __restrict__ blurred, const float2* __restrict__ flow, int fp, int h, int w, PrepArgs pa) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const ErrCtx c = make_err_ctx(pa.G1, w, h);
    const size_t p = (size_t)y * fp + x;
    emit_record(pa, c, x, y, w, h, flow[p], blurred[p]);
}

PrepArgs make_prep_args(const float* alpha0, const float* alpha1, const float2* G0, const float2* G1, SweepRec* rec, int w, int dir) {
    PrepArgs pa;
    pa.alpha0 = alpha0; pa.alpha1 = alpha1; pa.G0 = G0; pa.G1 = G1; pa.rec = rec;
    pa.nsteps_pad = sweep_nsteps_pad(w);
    pa.dir = dir;
    pa.slow = exact_div_width_ok(w) ? 0 : 1;
    return pa;
}

void launch_sweep_prep(const float* alpha0, const float* alpha1, const float2* G0, const float2* G1,
                       const float2* blurred, const float2* flow, int fp, SweepRec* rec, int h, int w, int dir, cudaStream_t st) {
    dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
    k_sweep_prep<<<g, b, 0, st>>>(blurred, flow, fp, h, w, make_prep_args(alpha0, alpha1, G0, G1, rec, w, dir));
}

// ---------------------------------------------------------------------------------------------------------
// the wavefront sweep
// ---------------------------------------------------------------------------------------------------------
#ifndef PF_SW_PREFETCH
#define PF_SW_PREFETCH 4
#endif
#ifndef PF_SWEEP_WARPS
#define PF_SWEEP_WARPS 4
#endif
constexpr int SW_PREFETCH_GATHER = PF_SW_PREFETCH;   // steps ahead for the L1 warm-up of the gradient gather
constexpr int SW_LL_RING = 64;                       // entries of a shared-memory LL ring (power of two)
#ifndef PF_SW_STAGE_STEPS
#define PF_SW_STAGE_STEPS 4
#endif
#ifndef PF_SW_NSTAGES
#define PF_SW_NSTAGES 3
#endif
constexpr int SW_STAGE_STEPS = PF_SW_STAGE_STEPS;    // wavefront steps per TMA bulk copy (4 -> 2 KB); also the unroll factor
constexpr int SW_NSTAGES = PF_SW_NSTAGES;            // stages of a warp's record ring (one being consumed, the others in flight)
constexpr int SW_ROWS = SWEEP_GROUP_ROWS;            // 16 rows per warp, two lanes per row
constexpr int SW_WARPS = PF_SWEEP_WARPS;             // compute warps per CTA (+ 1 poller warp)
constexpr int SW_ROWS_PER_CTA = SW_ROWS * SW_WARPS;
constexpr int SW_THREADS = (SW_WARPS + 1) * 32;
constexpr unsigned SW_STAGE_BYTES = SW_STAGE_STEPS * SW_ROWS * sizeof(SweepRec);

__device__ __forceinline__ uint4 ll_load_global(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ll_load_shared(unsigned saddr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void ll_store_shared(unsigned saddr, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// The hand-off of a warp's last row: one line {fx, flag, fy, flag}, to the global boundary array (last warp of a CTA) or to the next
// warp's shared-memory ring.  ONE predicated store through a generic address (the target is a loop invariant of the warp), so that the
// line is built once; no branch around it (the step body is latency-bound).
__device__ __forceinline__ void ll_store_generic_if(bool p, unsigned long long addr, uint4 v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0; @q st.volatile.v4.u32 [%0], {%1,%2,%3,%4}; }"
                 :: "l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void st_volatile_shared_s32(unsigned saddr, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_shared_s32(unsigned saddr) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// Waits until the LL line at `saddr` carries flag e in both halves; v holds a copy read earlier.  Every lane of the warp reads the SAME
// line, so the loop is warp-uniform: bra.uni tells the assembler so (no convergence barriers around the common, already-valid case).
__device__ __forceinline__ void ll_wait_shared_uniform(uint4& v, unsigned saddr, unsigned e, bool need) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.u32 p, %1, %5;\n\t"
                 "setp.ne.or.u32 p, %3, %5, p;\n\t"
                 "setp.ne.and.u32 p, %6, 0, p;\n\t"
                 "@!p bra.uni LL_DONE;\n\t"
                 "LL_SPIN:\n\t"
                 "ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t"
                 "setp.ne.u32 p, %1, %5;\n\t"
                 "setp.ne.or.u32 p, %3, %5, p;\n\t"
                 "@p bra.uni LL_SPIN;\n\t"
                 "LL_DONE:\n\t"
                 "}" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"(saddr), "r"(e), "r"((unsigned)need) : "memory");
}
// the exchange words {error bits, tag}: one 8-byte scalar access each (single-copy atomic)
__device__ __forceinline__ void xchg_store(unsigned saddr, float e, unsigned tag) {
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" :: "r"(saddr), "r"(__float_as_uint(e)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 xchg_load(unsigned saddr) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D TMA bulk copy (the record stream) ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void tma_bulk_g2s(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

struct SweepConst {
    const float2* G1s;
    unsigned touch;          // this lane's 16-byte scratch slot in shared memory (target of the L1 warm-up copies)
    int g1s_last, warm, pitch;          // warm: this lane's prefetch offset (elements of G1s ahead of the cell it gathers now)
    float wm2, hm2, fw, rcp_w, rcp_eps;
};

// getPixBilinear32FExtend (CPU/PixFlow.hpp:407-425) on the skewed layout: the clamped cell and the fractional parts
struct SkewPos { int gi; float xR, yR; };

template <int POSX>
__device__ __forceinline__ SkewPos skew_cell(const SweepConst& k, float mxr, float myr) {
    // fmaxf/fminf == the std::max/std::min of the reference (NaN -> 0 included)
    const float mx = fminf(fmaxf(mxr, 0.0f), k.wm2), my = fminf(fmaxf(myr, 0.0f), k.hm2);
    SkewPos c;
    const int x0 = __float2int_rz(mx), y0 = __float2int_rz(my);
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

__device__ __forceinline__ f2p skew_interp(const SkewCoef& t, float xR, float yR) {
    return padd(padd(padd(t.f00, pmuls(t.a2, xR)), pmuls(t.a3, yR)), pmuls(pmuls(t.a4, xR), yR));
}

// warm L1 with the anti-diagonal the gather reaches a few steps from now: an asynchronous 16-byte cp.async.ca into a
// scratch slot allocates the line in L1 and never blocks (its data is not used).  The skewed array has holes that no pixel maps to;
// a warm-up copy may read them (compute-sanitizer --tool initcheck reports exactly these reads and nothing else): the bytes go to
// the scratch slot and are never looked at.
__device__ __forceinline__ void warm_gather(const SweepConst& k, int gi) {
    int pi = gi + k.warm;
    pi = max(0, min(pi, k.g1s_last - 1)) & ~1;
    cp_async16(k.touch, k.G1s + pi);
}

// The three probes of ONE candidate on one lane.  Gradient descent on the piecewise-bilinear error parks many pixels within eps
// of a bilinear-cell boundary, so in most warp-steps some probe falls into the neighbouring cell (73 % in
// profiles/r1_sweep_v8_ncu.md).  Every probe therefore gathers its own cell unconditionally: 12 loads issued together --
// almost always the same one or two L1 lines -- instead of 4 loads, a warp vote, a branch and a second, dependent gather.
template <int POSX, bool SLOW>
__device__ __forceinline__ void eval_probes3(const SweepConst& k, float xf, float yf, float2 g0, float2 bl, float2 cand, bool warm,
                                             float v[3], unsigned& tiny) {
    const float fx0 = cand.x, fy0 = cand.y;
    const float fx1 = fadd(cand.x, PF_GRAD_EPS), fy2 = fadd(cand.y, PF_GRAD_EPS);
    const SkewPos c0 = skew_cell<POSX>(k, fadd(xf, fx0), fadd(yf, fy0));
    const SkewPos c1 = skew_cell<POSX>(k, fadd(xf, fx1), fadd(yf, fy0));
    const SkewPos c2 = skew_cell<POSX>(k, fadd(xf, fx0), fadd(yf, fy2));
    const SkewCoef t0 = skew_gather<POSX>(k, c0.gi), t1 = skew_gather<POSX>(k, c1.gi), t2 = skew_gather<POSX>(k, c2.gi);
    if (warm) warm_gather(k, c0.gi);
    // the part of errorFunction after the gather, for the three probes together (pf_prep.cuh)
    err3_from_g1<SLOW>(k.fw, k.rcp_w, g0, bl, skew_interp(t0, c0.xR, c0.yR), skew_interp(t1, c1.xR, c1.yR), skew_interp(t2, c2.xR, c2.yR),
                       fx0, fy0, fx1, fy2, v, tiny);
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
        tiny = min(tiny, min(tiny_key(fabsf(d.x)), tiny_key(fabsf(d.y))));
    }
    return upk(psub(pk(cand), pmuls(Q, PF_GRAD_STEP)));
}

// the reference's two compares (:318-320: left proposal first, then up, strict <) between the pixel's own record
// A = {E(f0), r0} and the finished candidates
__device__ __forceinline__ float2 select_result(float eL, float2 rL, float eU, float2 rU, bool leftValid, bool upValid, float4 A) {
    const float POS_INF = __int_as_float(0x7f800000);
    eL = leftValid ? eL : POS_INF;
    eU = upValid ? eU : POS_INF;
    const bool pL = eL < A.x;
    const float cur = pL ? eL : A.x;
    const bool pU = eU < cur;
    float2 out = make_float2(A.y, A.z);
    out = pL ? rL : out;
    out = pU ? rU : out;
    return out;
}

// The rare redo of a lane's step with the IEEE intrinsics (an operand left the verified range of the branch-free sequences):
// out of line, so that the hot loop stays compact in the instruction cache.  -> {E(cand), r.x, r.y}
template <int POSX>
__device__ __noinline__ float4 step_slow(SweepConst k, float xf, float yf, float2 g0, float2 bl, float2 cand) {
    unsigned dummy = 0xffffffffu;
    float v[3];
    eval_probes3<POSX, true>(k, xf, yf, g0, bl, cand, false, v, dummy);
    const float2 r = finish_candidate<true>(k, v, cand, dummy);
    return make_float4(v[0], r.x, r.y, 0.0f);
}

// Shared memory of one sweep CTA.
struct SweepSmem {
    int b;                                                        // row block being processed
    int progress[SW_WARPS];                                       // columns consumed from ring k
    __align__(8) unsigned long long full[SW_WARPS][SW_NSTAGES];   // mbarriers: stage filled by the TMA engine
    __align__(16) uint4 llring[SW_WARPS][SW_LL_RING];             // [0] inbound via the poller, [k] from warp k-1
    __align__(128) SweepRec rec[SW_WARPS][SW_NSTAGES][SW_STAGE_STEPS][SW_ROWS];   // TMA staging of the record streams
    uint4 touch[SW_WARPS][32];                                    // targets of the L1 warm-up copies
};

// One row block (SW_ROWS_PER_CTA logical rows) of one sweep: the part of compute warp wi.
template <int DIR, int POSX>
__device__ __forceinline__ void sweep_rows(const Sweep2Args& a, SweepSmem& sm, const int b, const int wi) {
    const unsigned full = 0xffffffffu;
    int w = a.s.w;
    const int h = a.s.h;
    int lane = threadIdx.x & 31;
    // Loop invariants must stay in registers: ptxas otherwise rematerialises them from the special registers / kernel
    // parameters inside the step (S2R, LDC, I2F ...) on a warp that is bound by its instruction count
    // (profiles/r2_sweep_v10_experiment.md).  A shuffle result is opaque to it.
    lane = __shfl_sync(full, lane, lane);
    auto pin_i = [&](int v) { return __shfl_sync(full, v, lane); };
    auto pin_f = [&](float v) { return __shfl_sync(full, v, lane); };
    const int g = lane >> 1, sub = lane & 1;                      // row of the warp, candidate of the row (0 left, 1 up)
    const int jw = b * SW_ROWS_PER_CTA + wi * SW_ROWS;            // first logical row of this warp
    if (jw >= h) return;
    const int j = jw + g;
    const bool rowValid = j < h;
    const int y = DIR > 0 ? j : h - 1 - j;
    const bool has_in = jw > 0;                                    // warp-uniform
    const bool has_out = jw + SW_ROWS < h;                         // warp-uniform
    const bool out_global = wi == SW_WARPS - 1;
    const unsigned rin = pin_i((int)smem_u32(&sm.llring[wi][0]));
    // hand-off target of the warp's last row as one generic address: hbase + (column & hmask) * 16, flag = ((column & emask) / ring) + 1
    // (global boundary array: no wrap, flag 1; shared ring of the next warp: wraps, flag = lap number)
    unsigned long long hbase = out_global ? reinterpret_cast<unsigned long long>(a.boundary + (size_t)b * a.s.w)
                                          : static_cast<unsigned long long>(__cvta_generic_to_shared(&sm.llring[(wi + 1) % SW_WARPS][0]));
    if (!out_global) {       // shared window -> generic address
        asm("cvta.shared.u64 %0, %0;" : "+l"(hbase));
    }
    hbase = ((unsigned long long)(unsigned)pin_i((int)(hbase >> 32)) << 32) | (unsigned)pin_i((int)(unsigned)hbase);
    const unsigned hmask = pin_i(out_global ? 0x7fffffff : SW_LL_RING - 1), emask = pin_i(out_global ? 0 : -1);
    const unsigned prog_in = smem_u32(&sm.progress[wi]);
    const unsigned prog_out = smem_u32(&sm.progress[(wi + 1) % SW_WARPS]);
    int out_limit = SW_LL_RING;                                    // columns < out_limit fit in the out ring unchecked

    SweepConst k;
    k.G1s = a.G1s; k.g1s_last = (int)a.g1s_last;
    k.touch = smem_u32(&sm.touch[wi][lane]);
    k.pitch = a.s.pitch;
    // L1 warm-up of the gather: the third anti-diagonal of the bilinear cell, SW_PREFETCH_GATHER steps ahead on lane 0 of a row and
    // twice as far on lane 1 (the near one catches what the far one mispredicted or lost: 40.2 -> 38.9 ms for one 4000 x 2000 pair)
        k.warm = 2 * k.pitch + 1 + SW_PREFETCH_GATHER * DIR * (k.pitch + POSX) * (1 + sub);
    k.wm2 = fsub((float)w, 2.0f); k.hm2 = fsub((float)h, 2.0f); k.fw = (float)w;
    k.rcp_w = __frcp_rn(k.fw); k.rcp_eps = __frcp_rn(PF_GRAD_EPS);
    w = pin_i(w); k.pitch = pin_i(k.pitch); k.warm = pin_i(k.warm); k.g1s_last = pin_i(k.g1s_last); k.touch = pin_i(k.touch);
    k.wm2 = pin_f(k.wm2); k.hm2 = pin_f(k.hm2); k.fw = pin_f(k.fw); k.rcp_w = pin_f(k.rcp_w); k.rcp_eps = pin_f(k.rcp_eps);
    {
        unsigned long long p = reinterpret_cast<unsigned long long>(k.G1s);
        p = ((unsigned long long)(unsigned)pin_i((int)(p >> 32)) << 32) | (unsigned)pin_i((int)(unsigned)p);
        k.G1s = reinterpret_cast<const float2*>(p);
    }
    // level width outside the verified range of div_by_const: every key test fails -> IEEE intrinsics everywhere
    const unsigned tkey0 = exact_div_width_ok(w) ? 0xffffffffu : 0u;
    const float yf = pin_f((float)y);
    float xf = (float)(DIR > 0 ? -g : w - 1 + g);                  // float(x) of step 0, then +-1 per step (exact)
    int i = rowValid ? -g : -0x40000000;                           // logical column of this row, +1 per step; rows past h never become valid
    float2* fptr = a.flow + (size_t)y * a.fp + (DIR > 0 ? -g : w - 1 + g);          // &flow(y, x) of the current step (dereferenced only where valid)

    // ---- record stream: 512 bytes per step, 4 steps per TMA bulk copy, SW_NSTAGES stages, fed by this warp's lane 0 ----
    const int nstages = sweep_nsteps_pad_dev(w) / SW_STAGE_STEPS;
    const char* stream = reinterpret_cast<const char*>(a.rec + (size_t)(jw / SW_ROWS) * (nstages * SW_STAGE_STEPS) * SW_ROWS);
    const unsigned rec0 = smem_u32(&sm.rec[wi][0][0][0]);
    const unsigned full0 = smem_u32(&sm.full[wi][0]);
    auto issue = [&](int t) {       // one lane: arm the stage's barrier with the byte count, then start the copy
        const int slot = t % SW_NSTAGES;
        mbar_arrive_expect_tx(full0 + slot * 8, SW_STAGE_BYTES);
        tma_bulk_g2s(rec0 + slot * SW_STAGE_BYTES, stream + (size_t)t * SW_STAGE_BYTES, SW_STAGE_BYTES, full0 + slot * 8);
    };
    if (lane == 0)
        for (int t = 0; t < SW_NSTAGES && t < nstages; ++t) issue(t);            // prologue: every slot
    const unsigned my_rec = pin_i((int)(rec0 + g * (unsigned)sizeof(SweepRec)));

    float2 res = make_float2(0.0f, 0.0f);
    if (has_in) {
        // Waiting for this warp's turn (the wavefront reaches row jw after ~jw steps): sleep-poll so that the
        // warps still queued behind the front leave the issue slots to the warps that are working.
        uint4 ln = ll_load_shared(rin);
        while (ln.y != 1u || ln.w != 1u) { __nanosleep(128); ln = ll_load_shared(rin); }
    }
    const int in_cols = has_in ? w : 0;           // columns to take from the inbound ring
    const bool ring_out = has_out && !out_global;
    const bool last_row = lane == 2 * (SW_ROWS - 1);               // lane 0 of the warp's last row
    const bool st_out = has_out && last_row;
    const bool storer = sub == 0;
    int s = 0;                                    // wavefront step

    // software pipeline: the records of step s + 1 are fetched at the end of step s (in the shadow of the exchange shuffles)
    auto load_rec = [&](unsigned addr, float4& A, float4& B) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(A.x), "=f"(A.y), "=f"(A.z), "=f"(A.w) : "r"(addr));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(B.x), "=f"(B.y), "=f"(B.z), "=f"(B.w) : "r"(addr + 16u));
    };
    constexpr unsigned STEP_BYTES = SW_ROWS * (unsigned)sizeof(SweepRec);
    float4 A, B;
    mbar_wait(full0, 0);
    load_rec(my_rec, A, B);
    uint4 rv = ll_load_shared(rin);

    for (int t = 0; t < nstages; ++t) {
        const unsigned srec = my_rec + (t % SW_NSTAGES) * SW_STAGE_BYTES;
        // back-pressure, once per stage (rare): the out ring must have room for the columns this stage hands down, s - 15 .. s - 15 + 3
        // (the consumer publishes its progress at its own stage ends and can always consume what has been handed down so far)
        if (ring_out)
            while (s + SW_STAGE_STEPS - SW_ROWS >= out_limit) out_limit = ld_volatile_shared_s32(prog_out) + SW_LL_RING;
#pragma unroll
        for (int u = 0; u < SW_STAGE_STEPS; ++u, ++s) {
            // steps past w + 14 (padding of the last stage) have no valid pixel and fall through the skip below
            const bool valid = (unsigned)i < (unsigned)w;
            const bool active = valid && __float_as_uint(A.x) != 0xff800000u;   // -inf marks "not updatable"; a NaN E(f0) stays active
            float2 out = make_float2(A.y, A.z);
            const float4 Ac = A;
            // Warp-uniform skip of steps in which no pixel of the warp is updatable (on the reference's real canvases ~85 % of the
            // steps: the overlap is sparse).  Such a step only carries the rows' old flows forward; in particular it does not
            // need the upper neighbours, so neither the shuffles nor the hand-off ring are touched (a ring entry is
            // self-validating and is simply never read).
            if (__any_sync(full, active)) {
                // ---- up neighbour: previous result of the row above (shuffle; first row of the warp: LL ring) ----
                float2 up;
                up.x = __shfl_up_sync(full, res.x, 2);
                up.y = __shfl_up_sync(full, res.y, 2);
                {   // warp-uniform: every lane reads the same ring entry (columns past in_cols: read, not waited for, not used)
                    const bool need = s < in_cols;
                    const unsigned e = ((unsigned)s / SW_LL_RING) + 1u;
                    const unsigned ra = rin + ((unsigned)s & (SW_LL_RING - 1)) * 16u;
                    uint4 v = rv;                                  // fetched at the end of the previous step (38.2 -> 37.5 ms per pair)
                    ll_wait_shared_uniform(v, ra, e, need);
                    up.x = (need && g == 0) ? __uint_as_float(v.x) : up.x;
                    up.y = (need && g == 0) ? __uint_as_float(v.z) : up.y;
                }
                const float2 g0 = make_float2(B.x, B.y), bl = make_float2(B.z, B.w);
                const float2 cand = make_float2(sub ? up.x : res.x, sub ? up.y : res.y);
                // ---- this lane's candidate: three probes, then its gradient step ----
                float v[3];
                unsigned tkey = tkey0;
                eval_probes3<POSX, false>(k, xf, yf, g0, bl, cand, true, v, tkey);
                const float oe = __shfl_xor_sync(full, v[0], 1);           // E first: it is ready before the gradient step
                float2 mine = finish_candidate<false>(k, v, cand, tkey);
                const bool bad = tkey < PF_TINY_BITS - 1u || !(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fabsf(v[2])) < 0x1p50f);
                float eM = v[0];
                float eO = oe;
                // the pair exchanges its finished candidates and selects BEFORE the range check of the branch-free sequences is known
                // (vote + branch off the dependent chain); the rare redo with the IEEE intrinsics repeats exchange and select
                float2 other = make_float2(__shfl_xor_sync(full, mine.x, 1), __shfl_xor_sync(full, mine.y, 1));
                if (u + 1 < SW_STAGE_STEPS) load_rec(srec + (u + 1) * STEP_BYTES, A, B);      // records of the next step
                out = select_result(sub ? eO : eM, sub ? other : mine, sub ? eM : eO, sub ? mine : other, i > 0, j > 0, Ac);
                if (__any_sync(full, bad && active)) {                     // out of line
                    const float4 sl = step_slow<POSX>(k, xf, yf, g0, bl, cand);
                    eM = sl.x; mine = make_float2(sl.y, sl.z);
                    eO = __shfl_xor_sync(full, eM, 1);
                    other = make_float2(__shfl_xor_sync(full, mine.x, 1), __shfl_xor_sync(full, mine.y, 1));
                    out = select_result(sub ? eO : eM, sub ? other : mine, sub ? eM : eO, sub ? mine : other, i > 0, j > 0, Ac);
                }
            } else {
                if (u + 1 < SW_STAGE_STEPS) load_rec(srec + (u + 1) * STEP_BYTES, A, B);
            }
            // the ring entry of the next step, early: when the upstream warp is a step ahead it is valid already and the wait below never loads
            if (u + 1 < SW_STAGE_STEPS) rv = ll_load_shared(rin + ((unsigned)(s + 1) & (SW_LL_RING - 1)) * 16u);
            res.x = valid ? out.x : res.x;
            res.y = valid ? out.y : res.y;
            // ---- results: flow (row-major, only where alpha > 0.9) and the hand-off of the warp's last row ----
            if (active && storer) *fptr = out;
            {
                const int i_last = s - (SW_ROWS - 1);                 // column of the warp's last row (warp-uniform)
                const unsigned e = (((unsigned)i_last & emask) / SW_LL_RING) + 1u;
                const uint4 lv = make_uint4(__float_as_uint(out.x), e, __float_as_uint(out.y), e);
                ll_store_generic_if(st_out && valid, hbase + (unsigned long long)((unsigned)i_last & hmask) * 16ull, lv);
            }
            xf = fadd(xf, (float)DIR);
            ++i;
            fptr += DIR;
        }
        // ---- stage boundary: every lane holds its last records of stage t in registers ----
        __syncwarp();
        if (lane == 0) {
            if (s <= in_cols) st_volatile_shared_s32(prog_in, s);          // done with the ring entries of columns < s
            const int tn = t + SW_NSTAGES;                                 // refill the slot just consumed (the proxy fence orders
            if (tn < nstages) {                                            // the warp's reads of it before the TMA's writes)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(tn);
            }
        }
        if (t + 1 < nstages) {
            const int sn = (t + 1) % SW_NSTAGES;
            mbar_wait(full0 + sn * 8, ((t + 1) / SW_NSTAGES) & 1);
            load_rec(my_rec + sn * SW_STAGE_BYTES, A, B);
            rv = ll_load_shared(rin + ((unsigned)s & (SW_LL_RING - 1)) * 16u);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// forwards the upstream CTA's global LL lines into ring 0 as they become valid (one warp per CTA)
__device__ __forceinline__ void sweep_poller(const Sweep2Args& a, SweepSmem& sm, const int b) {
    const unsigned full = 0xffffffffu;
    const int w = a.s.w, lane = threadIdx.x & 31;
    if (b == 0) return;
    const uint4* src = a.boundary + (size_t)(b - 1) * w;
    const unsigned prog = smem_u32(&sm.progress[0]);
    const unsigned ring0 = smem_u32(&sm.llring[0][0]);
    for (int base = 0; base < w; base += 32) {
        while (base + 32 > ld_volatile_shared_s32(prog) + SW_LL_RING) { }    // back-pressure: the batch must fit in the ring
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
}

// Persistent sweep kernel: the grid holds only about as many CTAs as the wavefront is wide (front + margin); a CTA
// that finishes its row block takes the next ticket.  Tickets are handed out in row-block order, so the block a CTA
// waits on is always being processed (or done) -- no deadlock whatever the residency -- and CTAs far behind the
// front do not occupy registers and shared memory while they would only be waiting for their turn.
#ifndef PF_SWEEP_MIN_CTAS
#define PF_SWEEP_MIN_CTAS 1
#endif
template <int DIR, int POSX>
__global__ void __launch_bounds__(SW_THREADS, PF_SWEEP_MIN_CTAS)
k_sweep(Sweep2Args a) {
    __shared__ SweepSmem sm;
    const int nblocks = (a.s.h + SW_ROWS_PER_CTA - 1) / SW_ROWS_PER_CTA;
    const int wi = threadIdx.x >> 5;
    bool first = true;
    for (;;) {
        __syncthreads();                                  // every warp is done with the previous block
        if (threadIdx.x == 0) {
            sm.b = atomicAdd(a.ticket, 1);
            for (int k = 0; k < SW_WARPS; ++k)
                for (int t = 0; t < SW_NSTAGES; ++t) {
                    if (!first) mbar_inval(smem_u32(&sm.full[k][t]));
                    mbar_init(smem_u32(&sm.full[k][t]), 1);
                }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        first = false;
        for (int i = threadIdx.x; i < SW_WARPS * SW_LL_RING; i += SW_THREADS) (&sm.llring[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < SW_WARPS) sm.progress[threadIdx.x] = 0;
        __syncthreads();
        const int b = sm.b;
        if (b >= nblocks) break;
        if (wi == SW_WARPS) sweep_poller(a, sm, b);
        else sweep_rows<DIR, POSX>(a, sm, b, wi);
    }
}

size_t sweep2_boundary_lines(int h, int w) {
    const int ncta = (h + SW_ROWS_PER_CTA - 1) / SW_ROWS_PER_CTA;
    return (size_t)(ncta > 1 ? ncta - 1 : 0) * (size_t)w + 1;
}

// Row blocks that work at the same time: block b starts SW_ROWS_PER_CTA steps after block b-1 and lasts w + SW_ROWS_PER_CTA - 1
// steps, so ceil((w + R - 1) / R) blocks overlap; `margin` more CTAs keep a free CTA ready when a block could start (latency)
// at the price of CTAs that sit waiting for their turn (SM slots, which is what bounds the throughput with many pairs in flight).
static int sweep_margin() {
    static int m = -1;
    if (m < 0) {
        const char* e = getenv("PF_SWEEP_MARGIN");
        m = e ? atoi(e) : 0;         // measured (profiles/r2_notes.md): 0 / 1 / 2 -> 1253 / 1240 / 1227 Mpix/s at 16 pairs, single pair unchanged
        if (m < 0 || m > 8) m = 0;
    }
    return m;
}

void launch_sweep2(const Sweep2Args& a, int dir, cudaStream_t st) {
    const int nblocks = (a.s.h + SW_ROWS_PER_CTA - 1) / SW_ROWS_PER_CTA;
    const int front = (a.s.w + 2 * SW_ROWS_PER_CTA - 1) / SW_ROWS_PER_CTA + sweep_margin();
    int ncta = nblocks < front ? nblocks : front;
    // Throughput flavour (many pairs in flight): CTA k of a launch cannot start before step k * SW_ROWS_PER_CTA, so a full front
    // of persistent CTAs spends ~front^2 / 2 block-steps resident but idle while the pipeline fills -- SM slots that, with many
    // wavefronts in flight, are what bounds the device.  A fraction of the front (each CTA then runs its row blocks back to back,
    // never waiting for its upstream) makes one sweep longer and the device fuller.
    if (a.cta_divisor > 1) ncta = (ncta + a.cta_divisor - 1) / a.cta_divisor;
    if (dir > 0) {
        if (a.s.posx) k_sweep<1, 1><<<ncta, SW_THREADS, 0, st>>>(a); else k_sweep<1, 0><<<ncta, SW_THREADS, 0, st>>>(a);
    } else {
        if (a.s.posx) k_sweep<-1, 1><<<ncta, SW_THREADS, 0, st>>>(a); else k_sweep<-1, 0><<<ncta, SW_THREADS, 0, st>>>(a);
    }
}

}  // namespace pf
