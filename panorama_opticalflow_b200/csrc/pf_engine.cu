// pf_engine.cu -- host-side engine and the C-ABI of libpixflow_b200.so (include/pixflow_b200.h).
//
// The engine owns per-pair workspaces in HBM (pyramids, gradients, per-direction flow buffers, boundary
// arenas), three streams per workspace (shared front end + one per flow direction) and enqueues the whole
// coarse-to-fine loop without any host round trip; the two directions L->R and R->L share the front end,
// pyramids and gradients (the reference recomputes them per direction, CPU/OpticalFlow.cpp:130-139).
#include <atomic>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pixflow_b200.h"
#include "pf_kernels.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
// A stream capture is invalidated by device-wide synchronising calls made anywhere in the process while it is open
// (cudaFree, cudaMalloc, cudaDeviceSynchronize ...).  All of this library's captures, allocations and teardowns take
// this mutex, so engines used from different threads cannot break each other's captures; a capture that is broken by
// foreign code falls back to plain stream launches for that call.
std::mutex g_capture_mu;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define PF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---- geometry of one flow problem (CPU/PixFlow.hpp:80-81, :137-151) -----------------------------------
struct Plan {
    int rows = 0, cols = 0, pad = 0, pcols = 0;
    int dh = 0, dw = 0, L = 0;
    std::vector<int> ws, hs;
    std::vector<size_t> off;       // pixel offset of each level in the pyramid arrays
    size_t total_px = 0;
    std::vector<size_t> bnd_off;   // per (level, sweep) offset in uint4 lines
    size_t bnd_lines = 0;
    std::vector<pf::Skew> skew;    // per-level skewed layout
    std::vector<size_t> skew_off;  // element offset of each level in the skewed gradient pyramids
    size_t skew_total = 0;

    void build(int rows_, int cols_, int pad_) {
        rows = rows_; cols = cols_; pad = pad_; pcols = cols + 2 * pad;
        dw = (int)((float)pcols * 0.5f);
        dh = (int)((float)rows * 0.5f);
        ws.clear(); hs.clear(); off.clear(); bnd_off.clear(); skew.clear(); skew_off.clear();
        ws.push_back(dw); hs.push_back(dh);
        while (ws.size() < 1000) {
            const int nw = (int)((float)ws.back() * 0.9f + 0.5f);
            const int nh = (int)((float)hs.back() * 0.9f + 0.5f);
            if (nh <= 24 || nw <= 24) break;
            ws.push_back(nw); hs.push_back(nh);
        }
        L = (int)ws.size();
        total_px = 0;
        bnd_lines = 0;
        skew_total = 0;
        for (int l = 0; l < L; ++l) {
            off.push_back(total_px);
            total_px += ((size_t)ws[l] * hs[l] + 63) & ~(size_t)63;   // keep every level 256-byte aligned
            skew.push_back(pf::make_skew(ws[l], hs[l]));
            skew_off.push_back(skew_total);
            skew_total += (pf::skew_elems(skew.back()) + 63) & ~(size_t)63;
            for (int s = 0; s < 2; ++s) {
                bnd_off.push_back(bnd_lines);
                bnd_lines += pf::sweep2_boundary_lines(hs[l], ws[l]);
            }
        }
    }
};

struct Workspace {
    Plan plan;
    int ndir = 2;
    uint8_t* in[2] = {nullptr, nullptr};          // staged BGRA inputs (rows x cols x 4, dense)
    float* I[2] = {nullptr, nullptr};
    float* A[2] = {nullptr, nullptr};
    float* Ipre = nullptr;
    float2* G[2] = {nullptr, nullptr};            // row-major gradient pyramids (Ix, Iy)
    float2* Gs[2] = {nullptr, nullptr};           // the same in the skewed layout (gathered by the sweeps)
    pf::SweepRec* rec[2] = {nullptr, nullptr};    // per direction: wavefront-packed records of the current sweep
    float2* bufA[2] = {nullptr, nullptr};         // per direction: flow ping
    float2* bufB[2] = {nullptr, nullptr};         // flow pong
    float2* blurred[2] = {nullptr, nullptr};
    // TMA tile maps of the flow buffers, per direction and level: blur box on bufA, median box on bufA and on bufB
    std::vector<pf::FlowTileMap> tmBlurA[2], tmMedA[2], tmMedB[2];
    float* ratio[2] = {nullptr, nullptr};
    uint4* bnd[2] = {nullptr, nullptr};
    int* tickets[2] = {nullptr, nullptr};
    float2* out[2] = {nullptr, nullptr};          // rows x cols flows (dense)
    uint8_t* merged = nullptr;                    // rows x cols BGRA (novel view)
    float* blend = nullptr;
    cudaStream_t sMain = nullptr, sDir[2] = {nullptr, nullptr};
    cudaEvent_t evReady = nullptr, evDone[2] = {nullptr, nullptr};
    // Latency flavour (a pair alone on the device): the gradient pyramids are built on a third stream, coarsest level first, while the
    // directions already work on the coarse levels -- evPyr: image pyramids complete, evG[l]: gradients of level l (both images) complete
    cudaStream_t sAux = nullptr;
    cudaEvent_t evPyr = nullptr;
    std::vector<cudaEvent_t> evG;
    bool overlap_front = false;                   // for the next enqueue
    cudaEvent_t evIn = nullptr;                   // host inputs staged (recorded on the engine's copy stream)
    cudaEvent_t evOut = nullptr;                  // host outputs copied back (recorded on the engine's download stream)
    bool outPending = false;                      // evOut must be waited for before the results are complete
    cudaEvent_t preWait = nullptr;                // (borrowed) event the pair's first operation must wait for, consumed by enqueue_pair
    cudaEvent_t evNovel = nullptr;                // novel view complete (recorded on sMain by the asynchronous form)
    std::vector<cudaEvent_t> sweepEv[2];          // optional timing events
    size_t nSweepEv[2] = {0, 0};
    // the whole pair pipeline (front end -> both directions -> join) captured once as a CUDA graph and replayed
    // with a single launch: ~1000 kernel launches per pair would otherwise cost ~3 ms of host time each call
    cudaGraphExec_t graph = nullptr;
    uint64_t graph_launches = 0;                  // kernels per replay
    int graph_key = -1;                           // ndir | hint0 << 4 | hint1 << 8 | search_dist << 12 | sweep_cta_divisor << 20 | overlap_front << 28
    int sweep_cta_divisor = 1;                    // for the next enqueue: 1 latency flavour, > 1 throughput flavour (launch_sweep2)

    ~Workspace() { release(); }
    void release() {
        if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; graph_key = -1; }
        for (int k = 0; k < 2; ++k) {
            cudaFree(in[k]); cudaFree(I[k]); cudaFree(A[k]); cudaFree(G[k]); cudaFree(Gs[k]); cudaFree(rec[k]);
            Gs[k] = nullptr; rec[k] = nullptr;
            cudaFree(bufA[k]); cudaFree(bufB[k]); cudaFree(blurred[k]);
            cudaFree(ratio[k]); cudaFree(bnd[k]); cudaFree(tickets[k]); cudaFree(out[k]);
            in[k] = nullptr; I[k] = A[k] = nullptr; G[k] = nullptr; bufA[k] = bufB[k] = blurred[k] = nullptr;
            ratio[k] = nullptr; bnd[k] = nullptr; tickets[k] = nullptr; out[k] = nullptr;
            if (sDir[k]) cudaStreamDestroy(sDir[k]);
            if (evDone[k]) cudaEventDestroy(evDone[k]);
            sDir[k] = nullptr; evDone[k] = nullptr;
            for (auto e : sweepEv[k]) cudaEventDestroy(e);
            sweepEv[k].clear();
        }
        cudaFree(Ipre); cudaFree(merged); cudaFree(blend);
        Ipre = nullptr; merged = nullptr; blend = nullptr;
        if (evReady) cudaEventDestroy(evReady);
        if (evPyr) cudaEventDestroy(evPyr);
        for (auto e : evG) cudaEventDestroy(e);
        evG.clear();
        if (sAux) cudaStreamDestroy(sAux);
        sAux = nullptr; evPyr = nullptr;
        if (evIn) cudaEventDestroy(evIn);
        if (evOut) cudaEventDestroy(evOut);
        if (evNovel) cudaEventDestroy(evNovel);
        evNovel = nullptr; preWait = nullptr;
        sMain = nullptr; evReady = nullptr; evIn = nullptr; evOut = nullptr; outPending = false;
    }

    int init(int rows, int cols, int pad, int idx = 0) {
        plan.build(rows, cols, pad);
        const Plan& p = plan;
        const size_t px0 = (size_t)p.dw * p.dh;
        const size_t fpx0 = (size_t)pf::flow_pitch(p.dw) * p.dh;       // pitched flow buffers (every level fits: pitch and rows shrink together)
        for (int k = 0; k < 2; ++k) {
            PF_CUDA(cudaMalloc(&in[k], (size_t)rows * cols * 4));
            PF_CUDA(cudaMalloc(&I[k], p.total_px * sizeof(float)));
            PF_CUDA(cudaMalloc(&A[k], p.total_px * sizeof(float)));
            PF_CUDA(cudaMalloc(&G[k], p.total_px * sizeof(float2)));
            PF_CUDA(cudaMalloc(&Gs[k], p.skew_total * sizeof(float2)));
            PF_CUDA(cudaMalloc(&rec[k], pf::sweep_rec_count(p.dh, p.dw) * sizeof(pf::SweepRec)));
            PF_CUDA(cudaMalloc(&bufA[k], fpx0 * sizeof(float2)));
            PF_CUDA(cudaMalloc(&bufB[k], fpx0 * sizeof(float2)));
            PF_CUDA(cudaMalloc(&blurred[k], fpx0 * sizeof(float2)));
            tmBlurA[k].resize(p.L); tmMedA[k].resize(p.L); tmMedB[k].resize(p.L);
            for (int l = 0; l < p.L; ++l) {
                const int fp = pf::flow_pitch(p.ws[l]);
                pf::make_blur_tile_map(&tmBlurA[k][l], bufA[k], p.hs[l], p.ws[l], fp);
                pf::make_median_tile_map(&tmMedA[k][l], bufA[k], p.hs[l], p.ws[l], fp);
                pf::make_median_tile_map(&tmMedB[k][l], bufB[k], p.hs[l], p.ws[l], fp);
            }
            PF_CUDA(cudaMalloc(&ratio[k], 256));
            PF_CUDA(cudaMalloc(&bnd[k], (p.bnd_lines + 1) * sizeof(uint4)));
            PF_CUDA(cudaMalloc(&tickets[k], (size_t)p.L * 2 * sizeof(int)));
            PF_CUDA(cudaMalloc(&out[k], (size_t)rows * cols * sizeof(float2)));
            {
                // Pairs in flight get DIFFERENT stream priorities (PF_STREAM_PRIO: 0 off, 1 round-robin over the
                // device's priority levels, 2 contiguous groups of 4; measured on B200: no gain, default off): identical pairs launched together would
                // otherwise march in lockstep -- all in their throughput-bound stencil kernels at the same time, then all
                // in their latency-bound sweeps -- and the two kinds of work would never overlap.
                static int mode = -1, least = 0, greatest = 0;
                if (mode < 0) {
                    const char* ev = getenv("PF_STREAM_PRIO");
                    mode = ev ? atoi(ev) : 0;
                    cudaDeviceGetStreamPriorityRange(&least, &greatest);
                }
                const int nlev = least - greatest + 1;
                int prio = least;
                if (mode == 1 && nlev > 1) prio = greatest + idx % nlev;
                if (mode == 2 && nlev > 1) prio = greatest + (idx / 4) % nlev;
                PF_CUDA(cudaStreamCreateWithPriority(&sDir[k], cudaStreamNonBlocking, prio));
            }
            PF_CUDA(cudaEventCreateWithFlags(&evDone[k], cudaEventDisableTiming));
        }
        PF_CUDA(cudaMalloc(&Ipre, px0 * sizeof(float)));
        sMain = sDir[0];   // the shared front end runs on direction 0's stream: two streams (hardware queues) per pair
        PF_CUDA(cudaEventCreateWithFlags(&evReady, cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&evPyr, cudaEventDisableTiming));
        PF_CUDA(cudaStreamCreateWithFlags(&sAux, cudaStreamNonBlocking));
        evG.resize(plan.L);
        for (auto& ev : evG) PF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&evIn, cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&evOut, cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&evNovel, cudaEventDisableTiming));
        return PF_OK;
    }
};

}  // namespace

struct pf_engine {
    int device = 0;
    int max_percentage = 0;     // template parameter of PixFlow<MaxPercentage>
    int search_dist = 0;        // computeSearchDistance, CPU/PixFlow.hpp:153-155
    bool time_sweeps = false;
    bool use_graphs = true;     // PF_NO_GRAPHS=1 disables (per-kernel stream launches, used when timing the sweeps)
    double last_sweep_ms = 0.0;
    uint64_t last_sweep_launches = 0;
    std::mutex mu;
    std::vector<Workspace*> pool;
    cudaStream_t sTimer = nullptr;
    // Host inputs of all pairs are staged through ONE copy stream, in pair order: issued on the pairs' own streams the 2n
    // uploads of a batch interleave on the DMA engine and every pair's inputs complete only when all have, which idles the
    // GPU for the whole upload; in FIFO order pair 0 starts computing after its own 2 images have arrived.
    cudaStream_t sCopy = nullptr;
    // ... and host OUTPUTS leave through one download stream of their own (second DMA engine): the D2H of a pair never sits
    // in front of another pair's kernels, and with the asynchronous calls the downloads of batch k overlap the compute of
    // batch k+1 (which runs on the other slot's workspaces).
    cudaStream_t sOut = nullptr;
    std::vector<Workspace*> inflight[2];          // workspaces of the last asynchronous batch of each slot (pf_wait)
    cudaEvent_t evT0 = nullptr, evT1 = nullptr;

    ~pf_engine() {
        for (auto* w : pool) delete w;
        if (evT0) cudaEventDestroy(evT0);
        if (evT1) cudaEventDestroy(evT1);
        if (sTimer) cudaStreamDestroy(sTimer);
        if (sCopy) cudaStreamDestroy(sCopy);
        if (sOut) cudaStreamDestroy(sOut);
    }

    // workspace #idx for the given geometry (created or re-created on demand)
    int workspace(int idx, int rows, int cols, int pad, Workspace** out) {
        while ((int)pool.size() <= idx) pool.push_back(nullptr);
        Workspace* w = pool[idx];
        if (w && (w->plan.rows != rows || w->plan.cols != cols || w->plan.pad != pad)) {
            std::lock_guard<std::mutex> cg(g_capture_mu);
            delete w; w = nullptr; pool[idx] = nullptr;
        }
        if (!w) {
            std::lock_guard<std::mutex> cg(g_capture_mu);
            w = new Workspace();
            const int rc = w->init(rows, cols, pad, idx);
            if (rc != PF_OK) { delete w; pool[idx] = nullptr; return rc; }
            pool[idx] = w;
        }
        *out = w;
        return PF_OK;
    }
};

namespace {

#define LAUNCHED(n) g_launches.fetch_add((n), std::memory_order_relaxed)

cudaEvent_t next_sweep_event(Workspace& w, int d) {
    if (w.nSweepEv[d] == w.sweepEv[d].size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        w.sweepEv[d].push_back(e);
    }
    return w.sweepEv[d][w.nSweepEv[d]++];
}

// front end + pyramids + gradients for both images, on sMain (CPU/PixFlow.hpp:78-110, :284-294)
int enqueue_shared(pf_engine* e, Workspace& w, const uint8_t* img[2], const size_t stride[2]) {
    const Plan& p = w.plan;
    (void)e;
    for (int k = 0; k < 2; ++k) {
        pf::launch_frontend_resize(img[k], stride[k], p.rows, p.cols, p.pad, w.Ipre, w.A[k] + p.off[0], p.dh, p.dw, w.sMain);
        pf::launch_gauss5(w.Ipre, w.I[k] + p.off[0], p.dh, p.dw, w.sMain);
        LAUNCHED(2);
    }
    for (int l = 1; l < p.L; ++l) {
        pf::PlaneSet ps;
        for (int k = 0; k < 2; ++k) {
            ps.src[k] = w.I[k] + p.off[l - 1]; ps.dst[k] = w.I[k] + p.off[l];
            ps.src[2 + k] = w.A[k] + p.off[l - 1]; ps.dst[2 + k] = w.A[k] + p.off[l];
        }
        pf::launch_pyr_down(ps, 4, p.hs[l - 1], p.ws[l - 1], p.hs[l], p.ws[l], w.sMain);
        LAUNCHED(1);
    }
    if (w.overlap_front) {
        // gradients on their own stream, coarsest level first: the directions start as soon as the coarsest level's are there and
        // the rest (90 % of the gradient work) runs under the coarse levels' sweeps, which leave the device almost empty
        PF_CUDA(cudaEventRecord(w.evPyr, w.sMain));
        PF_CUDA(cudaStreamWaitEvent(w.sAux, w.evPyr, 0));
        for (int l = p.L - 1; l >= 0; --l) {
            for (int k = 0; k < 2; ++k) {
                pf::launch_gradient(w.I[k] + p.off[l], w.G[k] + p.off[l], p.hs[l], p.ws[l], w.sAux);
                pf::launch_skew_copy_f2(w.G[k] + p.off[l], w.Gs[k] + p.skew_off[l], p.skew[l], w.sAux);
                LAUNCHED(2);
            }
            PF_CUDA(cudaEventRecord(w.evG[l], w.sAux));
        }
        PF_CUDA(cudaGetLastError());
        return PF_OK;
    }
    for (int l = 0; l < p.L; ++l)
        for (int k = 0; k < 2; ++k) {
            pf::launch_gradient(w.I[k] + p.off[l], w.G[k] + p.off[l], p.hs[l], p.ws[l], w.sMain);
            pf::launch_skew_copy_f2(w.G[k] + p.off[l], w.Gs[k] + p.skew_off[l], p.skew[l], w.sMain);
            LAUNCHED(2);
        }
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaEventRecord(w.evReady, w.sMain));
    return PF_OK;
}

// coarse-to-fine loop of one direction on sDir[d] (CPU/PixFlow.hpp:112-134); i0 = index of image I0
int enqueue_direction(pf_engine* e, Workspace& w, int d, int i0, int hint, float2* out, size_t out_stride) {
    const Plan& p = w.plan;
    const int i1 = 1 - i0;
    cudaStream_t st = w.sDir[d];
    PF_CUDA(cudaStreamWaitEvent(st, w.overlap_front ? w.evPyr : w.evReady, 0));
    PF_CUDA(cudaMemsetAsync(w.bnd[d], 0, (p.bnd_lines + 1) * sizeof(uint4), st));
    PF_CUDA(cudaMemsetAsync(w.tickets[d], 0, (size_t)p.L * 2 * sizeof(int), st));
    float2* flow = w.bufA[d];
    float2* other = w.bufB[d];
    for (int l = p.L - 1; l >= 0; --l) {
        const int h = p.hs[l], wd = p.ws[l], fp = pf::flow_pitch(wd);
        if (w.overlap_front) PF_CUDA(cudaStreamWaitEvent(st, w.evG[l], 0));
        const float* I0 = w.I[i0] + p.off[l];
        const float* I1 = w.I[i1] + p.off[l];
        const float* A0 = w.A[i0] + p.off[l];
        const float* A1 = w.A[i1] + p.off[l];
        if (l == p.L - 1) {
            pf::launch_initial_flow(I0, I1, A0, A1, flow, fp, w.ratio[d], h, wd, hint, e->search_dist, st);
            LAUNCHED(e->search_dist > 0 && hint != PF_HINT_UNKNOWN ? 2 : 1);
        }
        const float2* G0 = w.G[i0] + p.off[l];
        const float2* G1 = w.G[i1] + p.off[l];
        pf::Sweep2Args sa;
        sa.rec = w.rec[d];
        sa.G1s = w.Gs[i1] + p.skew_off[l];
        sa.s = p.skew[l];
        sa.g1s_last = (long long)pf::skew_elems(p.skew[l]) - 1;
        // blur of the incoming flow (the sweeps regularise against it) + records of the forward sweep
        pf::launch_blur15_prep(flow, w.blurred[d], h, wd, fp, A0, A1, G0, G1, w.rec[d], +1, &w.tmBlurA[d][l], st);
        // forward sweep, in place on `flow`
        sa.flow = flow;
        sa.fp = fp;
        sa.cta_divisor = w.sweep_cta_divisor;
        sa.boundary = w.bnd[d] + p.bnd_off[2 * l];
        sa.ticket = w.tickets[d] + 2 * l;
        if (e->time_sweeps) cudaEventRecord(next_sweep_event(w, d), st);
        pf::launch_sweep2(sa, +1, st);
        if (e->time_sweeps) cudaEventRecord(next_sweep_event(w, d), st);
        // median + records of the backward sweep
        pf::launch_median5_prep(flow, other, w.blurred[d], h, wd, fp, A0, A1, G0, G1, w.rec[d], -1, &w.tmMedA[d][l], st);
        // backward sweep, in place on `other`
        sa.flow = other;
        sa.boundary = w.bnd[d] + p.bnd_off[2 * l + 1];
        sa.ticket = w.tickets[d] + 2 * l + 1;
        if (e->time_sweeps) cudaEventRecord(next_sweep_event(w, d), st);
        pf::launch_sweep2(sa, -1, st);
        if (e->time_sweeps) cudaEventRecord(next_sweep_event(w, d), st);
        pf::launch_median5(other, flow, h, wd, fp, &w.tmMedB[d][l], st);
        // lowAlphaFlowDiffusion: blur + blend, written to `other`
        pf::launch_blur15_diffuse(flow, other, h, wd, fp, A0, A1, &w.tmBlurA[d][l], st);
        LAUNCHED(6);
        if (l > 0) {
            pf::launch_upsample_cubic(other, h, wd, fp, flow, p.hs[l - 1], p.ws[l - 1], pf::flow_pitch(p.ws[l - 1]), st);
            LAUNCHED(1);
        } else {
            pf::launch_tail(other, h, wd, fp, p.rows, p.pcols, p.pad, p.cols, out, out_stride, st);
            LAUNCHED(1);
        }
    }
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

int collect_sweep_timing(pf_engine* e, std::vector<Workspace*>& used) {
    e->last_sweep_ms = 0.0;
    e->last_sweep_launches = 0;
    if (!e->time_sweeps) return PF_OK;
    for (Workspace* w : used)
        for (int d = 0; d < 2; ++d) {
            for (size_t i = 0; i + 1 < w->nSweepEv[d]; i += 2) {
                float ms = 0.0f;
                PF_CUDA(cudaEventElapsedTime(&ms, w->sweepEv[d][i], w->sweepEv[d][i + 1]));
                e->last_sweep_ms += ms;
                e->last_sweep_launches += 1;
            }
            w->nSweepEv[d] = 0;
        }
    return PF_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_image_args(const void* p, size_t stride, int rows, int cols, size_t elem, const char* what) {
    if (!p) return fail(PF_ERR_INVALID_ARGUMENT, "%s is NULL", what);
    if (rows <= 0 || cols <= 0) return fail(PF_ERR_INVALID_ARGUMENT, "rows/cols must be positive");
    if (stride < (size_t)cols * elem) return fail(PF_ERR_INVALID_ARGUMENT, "%s stride %zu < cols*%zu", what, stride, elem);
    if (stride % 4 != 0 || ((uintptr_t)p) % 4 != 0) return fail(PF_ERR_INVALID_ARGUMENT, "%s must be 4-byte aligned", what);
    return PF_OK;
}

// stage an input image: device pointers are used in place, host pointers are copied into ws.in[k]
int stage_input(Workspace& w, int k, const void* img, size_t stride, const uint8_t** dptr, size_t* dstride, cudaStream_t st,
                bool force_copy = false) {
    if (!force_copy && is_device_ptr(img)) { *dptr = (const uint8_t*)img; *dstride = stride; return PF_OK; }
    const size_t dense = (size_t)w.plan.cols * 4;
    PF_CUDA(cudaMemcpy2DAsync(w.in[k], dense, img, stride, dense, w.plan.rows, cudaMemcpyDefault, st));
    *dptr = w.in[k]; *dstride = dense;
    return PF_OK;
}

// the body shared by compute_flow / prepare / batch / novel_view for ONE pair on workspace w (asynchronous)
int enqueue_pair(pf_engine* e, Workspace& w, const void* imgL, size_t strideL, const void* imgR, size_t strideR,
                 int ndir, const int hints[2], void* outs[2], const size_t ostrides[2],
                 const uint8_t* dimg[2], size_t dstride[2], float2* dflow[2], size_t dfstride[2]) {
    int rc;
    if (w.preWait) {            // inputs produced on another stream of this library (the stitching step)
        PF_CUDA(cudaStreamWaitEvent(w.sDir[0], w.preWait, 0));
        w.preWait = nullptr;
    }
    if (e->use_graphs && !e->time_sweeps) {
        // ---- graph path: inputs -> staging buffers, one graph launch, outputs <- workspace flow buffers ----
        cudaStream_t st = w.sDir[0];
        const bool hostL = !is_device_ptr(imgL), hostR = !is_device_ptr(imgR);
        if ((hostL || hostR) && !e->sCopy) PF_CUDA(cudaStreamCreateWithFlags(&e->sCopy, cudaStreamNonBlocking));
        if ((rc = stage_input(w, 0, imgL, strideL, &dimg[0], &dstride[0], hostL ? e->sCopy : st, true)) != PF_OK) return rc;
        if ((rc = stage_input(w, 1, imgR, strideR, &dimg[1], &dstride[1], hostR ? e->sCopy : st, true)) != PF_OK) return rc;
        if (hostL || hostR) {
            PF_CUDA(cudaEventRecord(w.evIn, e->sCopy));
            PF_CUDA(cudaStreamWaitEvent(st, w.evIn, 0));
        }
        const int key = ndir | (hints[0] << 4) | (hints[1] << 8) | (e->search_dist << 12) | (w.sweep_cta_divisor << 20) | ((int)w.overlap_front << 28);
        bool have_graph = w.graph_key == key;
        if (!have_graph) {
            std::lock_guard<std::mutex> cg(g_capture_mu);
            if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; w.graph_key = -1; }
            const uint64_t launched_before = g_launches.load();
            PF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            rc = enqueue_shared(e, w, dimg, dstride);
            for (int d = 0; d < ndir && rc == PF_OK; ++d)
                rc = enqueue_direction(e, w, d, d == 0 ? 0 : 1, hints[d], w.out[d], (size_t)w.plan.cols * sizeof(float2));
            if (rc == PF_OK && ndir == 2) {      // join direction 1 back into the capturing stream
                if (cudaEventRecord(w.evDone[1], w.sDir[1]) != cudaSuccess || cudaStreamWaitEvent(st, w.evDone[1], 0) != cudaSuccess)
                    rc = PF_ERR_CUDA;
            }
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(st, &g);
            if (rc == PF_OK && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&w.graph, g, 0);
            if (g) cudaGraphDestroy(g);
            const uint64_t captured = g_launches.load() - launched_before;
            g_launches.fetch_sub(captured);
            if (rc == PF_OK && ce == cudaSuccess && w.graph) {
                w.graph_key = key;
                w.graph_launches = captured;   // kernels per replay
                have_graph = true;
            } else {
                // the capture was invalidated (e.g. a device-wide synchronisation issued by other code in the process):
                // clear the error state and run this call with plain stream launches; capture is retried next call
                cudaGetLastError();
                if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; }
                w.graph_key = -1;
                g_err.clear();
            }
        }
        if (!have_graph) {
            if ((rc = enqueue_shared(e, w, dimg, dstride)) != PF_OK) return rc;
            for (int d = 0; d < ndir; ++d)
                if ((rc = enqueue_direction(e, w, d, d == 0 ? 0 : 1, hints[d], w.out[d], (size_t)w.plan.cols * sizeof(float2))) != PF_OK) return rc;
            if (ndir == 2) {
                PF_CUDA(cudaEventRecord(w.evDone[1], w.sDir[1]));
                PF_CUDA(cudaStreamWaitEvent(st, w.evDone[1], 0));
            }
        } else {
            PF_CUDA(cudaGraphLaunch(w.graph, st));
            LAUNCHED(w.graph_launches);
        }
        bool host_out = false;
        for (int d = 0; d < ndir; ++d) {
            dflow[d] = w.out[d];
            dfstride[d] = (size_t)w.plan.cols * sizeof(float2);
            if (outs[d] && is_device_ptr(outs[d]))
                PF_CUDA(cudaMemcpy2DAsync(outs[d], ostrides[d], w.out[d], dfstride[d], (size_t)w.plan.cols * sizeof(float2),
                                          w.plan.rows, cudaMemcpyDeviceToDevice, st));
            else if (outs[d]) host_out = true;
        }
        PF_CUDA(cudaEventRecord(w.evDone[0], st));
        if (ndir == 2) PF_CUDA(cudaEventRecord(w.evDone[1], st));
        if (host_out) {      // downloads on the engine's own D2H stream, behind this pair's compute only
            if (!e->sOut) PF_CUDA(cudaStreamCreateWithFlags(&e->sOut, cudaStreamNonBlocking));
            PF_CUDA(cudaStreamWaitEvent(e->sOut, w.evDone[0], 0));
            for (int d = 0; d < ndir; ++d)
                if (outs[d] && !is_device_ptr(outs[d]))
                    PF_CUDA(cudaMemcpy2DAsync(outs[d], ostrides[d], w.out[d], dfstride[d], (size_t)w.plan.cols * sizeof(float2),
                                              w.plan.rows, cudaMemcpyDeviceToHost, e->sOut));
            PF_CUDA(cudaEventRecord(w.evOut, e->sOut));
            w.outPending = true;
        }
        return PF_OK;
    }
    if ((rc = stage_input(w, 0, imgL, strideL, &dimg[0], &dstride[0], w.sMain)) != PF_OK) return rc;
    if ((rc = stage_input(w, 1, imgR, strideR, &dimg[1], &dstride[1], w.sMain)) != PF_OK) return rc;
    if ((rc = enqueue_shared(e, w, dimg, dstride)) != PF_OK) return rc;
    for (int d = 0; d < ndir; ++d) {
        const bool dev_out = outs[d] && is_device_ptr(outs[d]);
        dflow[d] = dev_out ? (float2*)outs[d] : w.out[d];
        dfstride[d] = dev_out ? ostrides[d] : (size_t)w.plan.cols * sizeof(float2);
        if ((rc = enqueue_direction(e, w, d, d == 0 ? 0 : 1, hints[d], dflow[d], dfstride[d])) != PF_OK) return rc;
        if (outs[d] && !dev_out)
            PF_CUDA(cudaMemcpy2DAsync(outs[d], ostrides[d], w.out[d], dfstride[d], (size_t)w.plan.cols * sizeof(float2),
                                      w.plan.rows, cudaMemcpyDeviceToHost, w.sDir[d]));
        PF_CUDA(cudaEventRecord(w.evDone[d], w.sDir[d]));
    }
    return PF_OK;
}

int sync_pair(Workspace& w, int ndir) {
    for (int d = 0; d < ndir; ++d) PF_CUDA(cudaStreamSynchronize(w.sDir[d]));
    PF_CUDA(cudaStreamSynchronize(w.sMain));
    if (w.outPending) { PF_CUDA(cudaEventSynchronize(w.evOut)); w.outPending = false; }
    return PF_OK;
}

}  // namespace

// ---- small RAII device buffer for the synchronous helper entry points ----
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { std::lock_guard<std::mutex> cg(g_capture_mu); cudaFree(p); }
    int alloc(size_t n) { std::lock_guard<std::mutex> cg(g_capture_mu); PF_CUDA(cudaMalloc(&p, n ? n : 1)); return PF_OK; }
    int upload(const void* h, size_t n) { int rc = alloc(n); if (rc) return rc; PF_CUDA(cudaMemcpy(p, h, n, cudaMemcpyHostToDevice)); return PF_OK; }
    int download(void* h, size_t n) { PF_CUDA(cudaDeviceSynchronize()); PF_CUDA(cudaGetLastError()); PF_CUDA(cudaMemcpy(h, p, n, cudaMemcpyDeviceToHost)); return PF_OK; }
    template <class T> T* as() { return (T*)p; }
};
#define RC(x) do { int rc_ = (x); if (rc_ != PF_OK) return rc_; } while (0)
}  // namespace

// ======================================================================================================
// C-ABI
// ======================================================================================================
extern "C" {

const char* pf_last_error(void) { return g_err.c_str(); }
const char* pf_version(void) { return "pixflow_b200 0.1 (sm_100a)"; }
uint64_t pf_kernel_launch_count(void) { return g_launches.load(); }

int pf_engine_create(const char* name, int device, pf_engine** out) {
    if (!out) return fail(PF_ERR_INVALID_ARGUMENT, "out_engine is NULL");
    *out = nullptr;
    if (!name) return fail(PF_ERR_INVALID_ARGUMENT, "flow_alg_name is NULL");
    int pct;
    if (strcmp(name, "pixflow_low") == 0) pct = 0;
    else if (strcmp(name, "pixflow_search_20") == 0) pct = 20;
    else return fail(PF_ERR_UNKNOWN_ALGORITHM, "unrecognized flow algorithm name: %s", name);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PF_ERR_NO_DEVICE, "no CUDA device available: libpixflow_b200 has no CPU fallback");
    }
    if (device < 0) PF_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return fail(PF_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    pf_engine* e = new pf_engine();
    e->device = device;
    e->max_percentage = pct;
    e->search_dist = (24 * pct + 50) / 100;
    e->use_graphs = getenv("PF_NO_GRAPHS") == nullptr;
    *out = e;
    g_err.clear();
    return PF_OK;
}

void release_stitch_bufs(pf_engine* e);

void pf_engine_destroy(pf_engine* e) {
    if (!e) return;
    {
        DeviceGuard g(e->device);
        release_stitch_bufs(e);
    }
    {
        std::lock_guard<std::mutex> cg(g_capture_mu);
        DeviceGuard g(e->device);
        for (Workspace* w : e->pool) {
            if (!w) continue;
            for (int d = 0; d < 2; ++d) if (w->sDir[d]) cudaStreamSynchronize(w->sDir[d]);
        }
        if (e->sOut) cudaStreamSynchronize(e->sOut);
        if (e->sCopy) cudaStreamSynchronize(e->sCopy);
        delete e;
    }
}

int pf_set_sweep_timing(pf_engine* e, int enabled) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    e->time_sweeps = enabled != 0;
    return PF_OK;
}
double pf_last_sweep_ms(pf_engine* e) { return e ? e->last_sweep_ms : 0.0; }
uint64_t pf_last_sweep_launches(pf_engine* e) { return e ? e->last_sweep_launches : 0; }

int pf_timer_start(pf_engine* e) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    if (!e->sTimer) {
        PF_CUDA(cudaStreamCreateWithFlags(&e->sTimer, cudaStreamNonBlocking));
        PF_CUDA(cudaEventCreate(&e->evT0));
        PF_CUDA(cudaEventCreate(&e->evT1));
    }
    PF_CUDA(cudaDeviceSynchronize());
    PF_CUDA(cudaEventRecord(e->evT0, e->sTimer));
    return PF_OK;
}

int pf_timer_stop(pf_engine* e, double* ms) {
    if (!e || !ms) return fail(PF_ERR_INVALID_ARGUMENT, "engine or elapsed_ms is NULL");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    if (!e->sTimer) return fail(PF_ERR_INVALID_ARGUMENT, "pf_timer_start was not called");
    for (Workspace* w : e->pool) {
        if (!w) continue;
        PF_CUDA(cudaStreamSynchronize(w->sMain));
        for (int d = 0; d < 2; ++d) PF_CUDA(cudaStreamSynchronize(w->sDir[d]));
    }
    if (e->sCopy) PF_CUDA(cudaStreamSynchronize(e->sCopy));
    if (e->sOut) PF_CUDA(cudaStreamSynchronize(e->sOut));
    PF_CUDA(cudaEventRecord(e->evT1, e->sTimer));
    PF_CUDA(cudaEventSynchronize(e->evT1));
    float f = 0.0f;
    PF_CUDA(cudaEventElapsedTime(&f, e->evT0, e->evT1));
    *ms = f;
    return PF_OK;
}

static int wait_slot_locked(pf_engine* e, int slot);
static bool front_overlap_for_pairs(int n);

int pf_compute_flow(pf_engine* e, const void* i0, size_t s0, const void* i1, size_t s1, int rows, int cols, int hint,
                    void* flow_out, size_t flow_stride) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(i0, s0, rows, cols, 4, "I0BGRA")) != PF_OK) return rc;
    if ((rc = check_image_args(i1, s1, rows, cols, 4, "I1BGRA")) != PF_OK) return rc;
    if ((rc = check_image_args(flow_out, flow_stride, rows, cols, 8, "flow")) != PF_OK) return rc;
    if (hint < 0 || hint > 4) return fail(PF_ERR_INVALID_ARGUMENT, "unexpected direction %d", hint);
    if ((int)((float)cols * 0.5f) < 4 || (int)((float)rows * 0.5f) < 4) return fail(PF_ERR_INVALID_ARGUMENT, "image too small");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    Workspace* w;
    if ((rc = wait_slot_locked(e, 0)) != PF_OK) return rc;        // an asynchronous batch may still own workspace 0
    if ((rc = e->workspace(0, rows, cols, 0, &w)) != PF_OK) return rc;
    w->sweep_cta_divisor = 1;
    w->overlap_front = front_overlap_for_pairs(1);
    const int hints[2] = {hint, 0};
    void* outs[2] = {flow_out, nullptr};
    const size_t ostr[2] = {flow_stride, 0};
    const uint8_t* dimg[2]; size_t dstr[2]; float2* dflow[2]; size_t dfs[2];
    if ((rc = enqueue_pair(e, *w, i0, s0, i1, s1, 1, hints, outs, ostr, dimg, dstr, dflow, dfs)) != PF_OK) return rc;
    if ((rc = sync_pair(*w, 1)) != PF_OK) return rc;
    std::vector<Workspace*> used{w};
    return collect_sweep_timing(e, used);
}

// Sweep launch flavour for n pairs in flight on the device: alone (or nearly), a pair waits for its wavefronts' dependent chains ->
// full front of CTAs; with many pairs in flight SM slots are the scarce resource -> a fraction of the front (launch_sweep2).
// PF_SWEEP_CTA_DIVISOR sets the throughput flavour's divisor (default 1 = off: measured 1247 / 1246 / 1118 / 993 Mpix/s for 1 / 2 / 3 / 4), PF_LATENCY_MAX_PAIRS the switch-over (default 2).
static int sweep_cta_divisor_for_pairs(int n) {
    static int max_pairs = -1, div = 2;
    if (max_pairs < 0) {
        const char* a = getenv("PF_LATENCY_MAX_PAIRS");
        const char* b = getenv("PF_SWEEP_CTA_DIVISOR");
        max_pairs = a ? atoi(a) : 2;
        div = b ? atoi(b) : 1;
        if (div < 1 || div > 16) div = 1;
    }
    return n <= max_pairs ? 1 : div;
}
// the same switch-over for the overlapped gradient chain (Workspace::overlap_front); PF_NO_FRONT_OVERLAP=1 turns it off
static bool front_overlap_for_pairs(int n) {
    static int max_pairs = -1;
    if (max_pairs < 0) {
        const char* a = getenv("PF_LATENCY_MAX_PAIRS");
        const char* off = getenv("PF_NO_FRONT_OVERLAP");
        max_pairs = (off && atoi(off) != 0) ? 0 : (a ? atoi(a) : 2);
    }
    return n <= max_pairs;
}

// workspace index of pair i of slot s: the two slots own disjoint workspaces (streams, graphs, staging and output buffers)
static inline int ws_index(int slot, int i) { return 2 * i + slot; }

static int wait_slot_locked(pf_engine* e, int slot) {
    int rc = PF_OK;
    for (Workspace* w : e->inflight[slot]) {
        const int r = sync_pair(*w, 2);
        if (r != PF_OK && rc == PF_OK) rc = r;
    }
    if (rc == PF_OK) rc = collect_sweep_timing(e, e->inflight[slot]);
    e->inflight[slot].clear();
    return rc;
}

static int batch_enqueue_locked(pf_engine* e, int slot, int n, const void* const* Ls, size_t sl, const void* const* Rs, size_t sr,
                                int rows, int cols, void* const* lr, size_t slr, void* const* rl, size_t srl) {
    int rc;
    if (!e->inflight[slot].empty() && (rc = wait_slot_locked(e, slot)) != PF_OK) return rc;    // the slot's buffers are re-used
    const int pad = cols / 20;   // CPU/OpticalFlow.cpp:113
    const int hints[2] = {PF_HINT_LEFT, PF_HINT_RIGHT};   // CPU/OpticalFlow.cpp:130-139
    for (int i = 0; i < n; ++i) {
        Workspace* w;
        if ((rc = e->workspace(ws_index(slot, i), rows, cols, pad, &w)) != PF_OK) return rc;
        e->inflight[slot].push_back(w);
        w->sweep_cta_divisor = sweep_cta_divisor_for_pairs(n);
        w->overlap_front = front_overlap_for_pairs(n);
        void* outs[2] = {lr[i], rl[i]};
        const size_t ostr[2] = {slr, srl};
        const uint8_t* dimg[2]; size_t dstr[2]; float2* dflow[2]; size_t dfs[2];
        if ((rc = enqueue_pair(e, *w, Ls[i], sl, Rs[i], sr, 2, hints, outs, ostr, dimg, dstr, dflow, dfs)) != PF_OK) return rc;
    }
    return PF_OK;
}

static int batch_check_args(int n, const void* const* Ls, size_t sl, const void* const* Rs, size_t sr, int rows, int cols,
                            void* const* lr, size_t slr, void* const* rl, size_t srl) {
    if (n <= 0 || !Ls || !Rs || !lr || !rl) return fail(PF_ERR_INVALID_ARGUMENT, "bad batch arguments");
    int rc;
    for (int i = 0; i < n; ++i) {
        if ((rc = check_image_args(Ls[i], sl, rows, cols, 4, "imageL")) != PF_OK) return rc;
        if ((rc = check_image_args(Rs[i], sr, rows, cols, 4, "imageR")) != PF_OK) return rc;
        if ((rc = check_image_args(lr[i], slr, rows, cols, 8, "flowLtoR")) != PF_OK) return rc;
        if ((rc = check_image_args(rl[i], srl, rows, cols, 8, "flowRtoL")) != PF_OK) return rc;
    }
    const int pad = cols / 20;
    if ((int)((float)(cols + 2 * pad) * 0.5f) < 4 || (int)((float)rows * 0.5f) < 4) return fail(PF_ERR_INVALID_ARGUMENT, "image too small");
    return PF_OK;
}

int pf_prepare_bidirectional_batch_async(pf_engine* e, int slot, int n, const void* const* Ls, size_t sl, const void* const* Rs, size_t sr,
                                         int rows, int cols, void* const* lr, size_t slr, void* const* rl, size_t srl) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (slot < 0 || slot > 1) return fail(PF_ERR_INVALID_ARGUMENT, "slot must be 0 or 1");
    int rc;
    if ((rc = batch_check_args(n, Ls, sl, Rs, sr, rows, cols, lr, slr, rl, srl)) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    return batch_enqueue_locked(e, slot, n, Ls, sl, Rs, sr, rows, cols, lr, slr, rl, srl);
}

int pf_wait(pf_engine* e, int slot) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (slot < 0 || slot > 1) return fail(PF_ERR_INVALID_ARGUMENT, "slot must be 0 or 1");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    return wait_slot_locked(e, slot);
}

int pf_prepare_bidirectional_batch(pf_engine* e, int n, const void* const* Ls, size_t sl, const void* const* Rs, size_t sr,
                                   int rows, int cols, void* const* lr, size_t slr, void* const* rl, size_t srl) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = batch_check_args(n, Ls, sl, Rs, sr, rows, cols, lr, slr, rl, srl)) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    const auto t_begin = std::chrono::steady_clock::now();
    if ((rc = batch_enqueue_locked(e, 0, n, Ls, sl, Rs, sr, rows, cols, lr, slr, rl, srl)) != PF_OK) return rc;
    const auto t_enq = std::chrono::steady_clock::now();
    rc = wait_slot_locked(e, 0);
    if (getenv("PF_DEBUG_TIMING")) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[pf] batch n=%d: enqueue %.2f ms, wait %.2f ms\n", n,
                std::chrono::duration<double, std::milli>(t_enq - t_begin).count(),
                std::chrono::duration<double, std::milli>(t_end - t_enq).count());
    }
    return rc;
}


int pf_prepare_bidirectional(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, int rows, int cols,
                             void* lr, size_t slr, void* rl, size_t srl) {
    const void* Ls[1] = {L};
    const void* Rs[1] = {R};
    void* lrs[1] = {lr};
    void* rls[1] = {rl};
    return pf_prepare_bidirectional_batch(e, 1, Ls, sl, Rs, sr, rows, cols, lrs, slr, rls, srl);
}

static int combine_impl(pf_engine* e, Workspace* w, cudaStream_t st, const uint8_t* dL, size_t sL, const uint8_t* dR, size_t sR,
                        const float2* dLR, size_t sLR, const float2* dRL, size_t sRL, const void* blend, size_t sB,
                        int rows, int cols, void* out, size_t sOut) {
    (void)e;
    const float* dB;
    size_t dsB;
    if (is_device_ptr(blend)) { dB = (const float*)blend; dsB = sB; }
    else {
        if (!w->blend) { std::lock_guard<std::mutex> cg(g_capture_mu); PF_CUDA(cudaMalloc(&w->blend, (size_t)rows * cols * 4)); }
        PF_CUDA(cudaMemcpy2DAsync(w->blend, (size_t)cols * 4, blend, sB, (size_t)cols * 4, rows, cudaMemcpyHostToDevice, st));
        dB = w->blend; dsB = (size_t)cols * 4;
    }
    const bool dev_out = is_device_ptr(out);
    uint8_t* dO;
    size_t dsO;
    if (dev_out) { dO = (uint8_t*)out; dsO = sOut; }
    else {
        if (!w->merged) { std::lock_guard<std::mutex> cg(g_capture_mu); PF_CUDA(cudaMalloc(&w->merged, (size_t)rows * cols * 4)); }
        dO = w->merged; dsO = (size_t)cols * 4;
    }
    pf::launch_combine(dL, sL, dR, sR, dLR, sLR, dRL, sRL, dB, dsB, rows, cols, dO, dsO, st);
    LAUNCHED(1);
    PF_CUDA(cudaGetLastError());
    if (!dev_out) PF_CUDA(cudaMemcpy2DAsync(out, sOut, dO, dsO, (size_t)cols * 4, rows, cudaMemcpyDeviceToHost, st));
    return PF_OK;
}

int pf_combine_novel_views(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, const void* lr, size_t slr,
                           const void* rl, size_t srl, const void* blend, size_t sb, int rows, int cols, void* out, size_t so) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(L, sl, rows, cols, 4, "imageL")) != PF_OK) return rc;
    if ((rc = check_image_args(R, sr, rows, cols, 4, "imageR")) != PF_OK) return rc;
    if ((rc = check_image_args(lr, slr, rows, cols, 8, "flowLtoR")) != PF_OK) return rc;
    if ((rc = check_image_args(rl, srl, rows, cols, 8, "flowRtoL")) != PF_OK) return rc;
    if ((rc = check_image_args(blend, sb, rows, cols, 4, "blend")) != PF_OK) return rc;
    if ((rc = check_image_args(out, so, rows, cols, 4, "out")) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    Workspace* w;
    if ((rc = wait_slot_locked(e, 0)) != PF_OK) return rc;
    if ((rc = e->workspace(0, rows, cols, cols / 20, &w)) != PF_OK) return rc;
    cudaStream_t st = w->sMain;
    const uint8_t* dimg[2]; size_t dstr[2];
    if ((rc = stage_input(*w, 0, L, sl, &dimg[0], &dstr[0], st)) != PF_OK) return rc;
    if ((rc = stage_input(*w, 1, R, sr, &dimg[1], &dstr[1], st)) != PF_OK) return rc;
    const float2* dfl[2]; size_t dfs[2];
    const void* fl[2] = {lr, rl};
    const size_t fs[2] = {slr, srl};
    for (int d = 0; d < 2; ++d) {
        if (is_device_ptr(fl[d])) { dfl[d] = (const float2*)fl[d]; dfs[d] = fs[d]; }
        else {
            PF_CUDA(cudaMemcpy2DAsync(w->out[d], (size_t)cols * 8, fl[d], fs[d], (size_t)cols * 8, rows, cudaMemcpyHostToDevice, st));
            dfl[d] = w->out[d]; dfs[d] = (size_t)cols * 8;
        }
    }
    if ((rc = combine_impl(e, w, st, dimg[0], dstr[0], dimg[1], dstr[1], dfl[0], dfs[0], dfl[1], dfs[1], blend, sb, rows, cols, out, so)) != PF_OK) return rc;
    PF_CUDA(cudaStreamSynchronize(st));
    return PF_OK;
}

// body of pf_novel_view; e->mu held by the caller, arguments already checked
// inputs_ready / blend_ready: optional events (the images / the blend map are produced on another stream); done: when given,
// the call does not wait -- *done is the event to wait for (device-side chaining by the stitching step)
static int novel_view_locked(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, const void* blend, size_t sb,
                             int rows, int cols, void* out, size_t so, void* lr, size_t slr, void* rl, size_t srl,
                             cudaEvent_t inputs_ready = nullptr, cudaEvent_t blend_ready = nullptr, cudaEvent_t* done = nullptr) {
    int rc;
    Workspace* w;
    const int pad = cols / 20;
    if ((rc = wait_slot_locked(e, 0)) != PF_OK) return rc;
    if ((rc = e->workspace(0, rows, cols, pad, &w)) != PF_OK) return rc;
    w->sweep_cta_divisor = 1;
    w->overlap_front = front_overlap_for_pairs(1);
    w->preWait = inputs_ready;
    const int hints[2] = {PF_HINT_LEFT, PF_HINT_RIGHT};
    void* outs[2] = {lr, rl};
    const size_t ostr[2] = {slr, srl};
    const uint8_t* dimg[2]; size_t dstr[2]; float2* dflow[2]; size_t dfs[2];
    if ((rc = enqueue_pair(e, *w, L, sl, R, sr, 2, hints, outs, ostr, dimg, dstr, dflow, dfs)) != PF_OK) return rc;
    // join both directions into sMain, then blend there
    for (int d = 0; d < 2; ++d) PF_CUDA(cudaStreamWaitEvent(w->sMain, w->evDone[d], 0));
    if (blend_ready) PF_CUDA(cudaStreamWaitEvent(w->sMain, blend_ready, 0));
    if ((rc = combine_impl(e, w, w->sMain, dimg[0], dstr[0], dimg[1], dstr[1], dflow[0], dfs[0], dflow[1], dfs[1],
                           blend, sb, rows, cols, out, so)) != PF_OK) return rc;
    if (done) {
        PF_CUDA(cudaEventRecord(w->evNovel, w->sMain));
        *done = w->evNovel;
        return PF_OK;
    }
    if ((rc = sync_pair(*w, 2)) != PF_OK) return rc;
    std::vector<Workspace*> used{w};
    return collect_sweep_timing(e, used);
}

int pf_novel_view(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, const void* blend, size_t sb,
                  int rows, int cols, void* out, size_t so, void* lr, size_t slr, void* rl, size_t srl) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(L, sl, rows, cols, 4, "imageL")) != PF_OK) return rc;
    if ((rc = check_image_args(R, sr, rows, cols, 4, "imageR")) != PF_OK) return rc;
    if ((rc = check_image_args(blend, sb, rows, cols, 4, "blend")) != PF_OK) return rc;
    if ((rc = check_image_args(out, so, rows, cols, 4, "out")) != PF_OK) return rc;
    if (lr && (rc = check_image_args(lr, slr, rows, cols, 8, "flowLtoR")) != PF_OK) return rc;
    if (rl && (rc = check_image_args(rl, srl, rows, cols, 8, "flowRtoL")) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    return novel_view_locked(e, L, sl, R, sr, blend, sb, rows, cols, out, so, lr, slr, rl, srl);
}

// ---- the stitching step around the flow path -----------------------------------------

// device buffers of the stitching entry points, kept across calls (a 36 Mpx canvas needs ~1.7 GB of them)
struct StitchBufs {
    int rows = 0, cols = 0;
    uint8_t *L = nullptr, *R = nullptr, *map = nullptr, *oL = nullptr, *oR = nullptr, *merged = nullptr, *gmap = nullptr, *result = nullptr;
    float *braw = nullptr, *mdis = nullptr, *blend = nullptr;
    void* scratch = nullptr;
    cudaStream_t st = nullptr;                 // the stitching kernels' own (non-blocking) stream
    cudaEvent_t evMasked = nullptr, evBlend = nullptr;     // OverlappedL/R ready; Blend ready
    void release() {
        std::lock_guard<std::mutex> cg(g_capture_mu);
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; }
        if (evMasked) { cudaEventDestroy(evMasked); evMasked = nullptr; }
        if (evBlend) { cudaEventDestroy(evBlend); evBlend = nullptr; }
        cudaFree(L); cudaFree(R); cudaFree(map); cudaFree(oL); cudaFree(oR); cudaFree(merged); cudaFree(gmap); cudaFree(result);
        cudaFree(braw); cudaFree(mdis); cudaFree(blend); cudaFree(scratch);
        L = R = map = oL = oR = merged = gmap = result = nullptr; braw = mdis = blend = nullptr; scratch = nullptr;
        rows = cols = 0;
    }
    size_t cap_px = 0, cap_scratch = 0;          // capacity: canvases of different sizes alternate without re-allocation
    int ensure(int r, int c) {
        if (r == rows && c == cols) return PF_OK;
        const size_t need_px = (size_t)r * c, need_scratch = pf::stitch_smooth_scratch_bytes(r, c);
        if (st && need_px <= cap_px && need_scratch <= cap_scratch) { rows = r; cols = c; return PF_OK; }
        release();
        std::lock_guard<std::mutex> cg(g_capture_mu);
        const size_t n = need_px > cap_px ? need_px : cap_px;
        PF_CUDA(cudaMalloc(&L, n * 4)); PF_CUDA(cudaMalloc(&R, n * 4)); PF_CUDA(cudaMalloc(&map, n));
        PF_CUDA(cudaMalloc(&oL, n * 4)); PF_CUDA(cudaMalloc(&oR, n * 4)); PF_CUDA(cudaMalloc(&merged, n * 4));
        PF_CUDA(cudaMalloc(&gmap, n)); PF_CUDA(cudaMalloc(&result, n * 4));
        PF_CUDA(cudaMalloc(&braw, n * 4)); PF_CUDA(cudaMalloc(&mdis, n * 4)); PF_CUDA(cudaMalloc(&blend, n * 4));
        const size_t sb = need_scratch > cap_scratch ? need_scratch : cap_scratch;
        PF_CUDA(cudaMalloc(&scratch, sb ? sb : 16));
        cap_px = n; cap_scratch = sb;
        PF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        PF_CUDA(cudaEventCreateWithFlags(&evMasked, cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&evBlend, cudaEventDisableTiming));
        rows = r; cols = c;
        return PF_OK;
    }
};
static std::mutex g_stitch_mu;
static std::map<pf_engine*, StitchBufs> g_stitch_bufs;      // released by pf_engine_destroy

static StitchBufs& stitch_bufs(pf_engine* e) {
    std::lock_guard<std::mutex> lk(g_stitch_mu);
    return g_stitch_bufs[e];
}
void release_stitch_bufs(pf_engine* e) {
    std::lock_guard<std::mutex> lk(g_stitch_mu);
    auto it = g_stitch_bufs.find(e);
    if (it != g_stitch_bufs.end()) { it->second.release(); g_stitch_bufs.erase(it); }
}

static int stitch_size_check(int rows, int cols, bool need_smooth) {
    if (((cols <= rows) ? cols / 200 : rows / 200) < 1)
        return fail(PF_ERR_INVALID_ARGUMENT, "image too small for Stitchtools::countblend (shorter side < 200)");
    if (need_smooth) {
        int step, k1, k2; size_t smem;
        const int g = pf::stitch_smooth_geometry(rows, cols, &step, &k1, &k2, &smem);
        if (g == 1) return fail(PF_ERR_INVALID_ARGUMENT, "image too small for GenerateBlend's blur (rows < 400: cv::blur kernel rows/400 is empty)");
        if (g == 2) return fail(PF_ERR_INVALID_ARGUMENT, "image shape not supported by the blend smoothing (window exceeds shared memory)");
    }
    return PF_OK;
}

// Stitchtools::prepare on device buffers held in sb (inputs already in sb.L/sb.R or given in place); asynchronous on st
static int stitch_prepare_dev(StitchBufs& sb, const uint8_t* dL, size_t sl, const uint8_t* dR, size_t sr, int rows, int cols,
                              bool smooth, cudaStream_t st) {
    const size_t c1 = (size_t)cols, c4 = (size_t)cols * 4;
    pf::launch_stitch_match_mask(dL, sl, dR, sr, rows, cols, sb.map, c1, sb.oL, c4, sb.oR, c4, st);
    PF_CUDA(cudaEventRecord(sb.evMasked, st));
    pf::launch_stitch_blend_raw(sb.map, c1, rows, cols, sb.braw, c4, sb.mdis, c4, st);
    LAUNCHED(2);
    if (smooth) {
        PF_CUDA(cudaMemcpyAsync(sb.blend, sb.braw, (size_t)rows * c4, cudaMemcpyDeviceToDevice, st));
        const int n = pf::launch_stitch_blend_smooth(sb.blend, c4, sb.mdis, c4, rows, cols, sb.scratch, st);
        if (n == 0) return fail(PF_ERR_INVALID_ARGUMENT, "blend smoothing: unsupported image shape");
        LAUNCHED(n);
    }
    PF_CUDA(cudaEventRecord(sb.evBlend, st));
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

static int stage_u8(const void* src, size_t stride, int rows, size_t row_bytes, uint8_t* dev, const uint8_t** p, size_t* ps, cudaStream_t st) {
    if (is_device_ptr(src)) { *p = (const uint8_t*)src; *ps = stride; return PF_OK; }
    PF_CUDA(cudaMemcpy2DAsync(dev, row_bytes, src, stride, row_bytes, rows, cudaMemcpyHostToDevice, st));
    *p = dev; *ps = row_bytes;
    return PF_OK;
}
static int copy_out(void* user, size_t ustride, const void* dev, size_t row_bytes, int rows, cudaStream_t st) {
    if (!user) return PF_OK;
    PF_CUDA(cudaMemcpy2DAsync(user, ustride, dev, row_bytes, row_bytes, rows, cudaMemcpyDefault, st));
    return PF_OK;
}

int pf_stitch_prepare(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, int rows, int cols,
                      void* map, size_t sm, void* oL, size_t sol, void* oR, size_t sor, void* braw, size_t sbr, void* mdis, size_t sd,
                      void* blend, size_t sbl) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(L, sl, rows, cols, 4, "colorImageL")) != PF_OK) return rc;
    if ((rc = check_image_args(R, sr, rows, cols, 4, "colorImageR")) != PF_OK) return rc;
    if ((rc = stitch_size_check(rows, cols, blend != nullptr)) != PF_OK) return rc;
    if (map && sm < (size_t)cols) return fail(PF_ERR_INVALID_ARGUMENT, "map stride < cols");
    if (oL && (rc = check_image_args(oL, sol, rows, cols, 4, "OverlappedL")) != PF_OK) return rc;
    if (oR && (rc = check_image_args(oR, sor, rows, cols, 4, "OverlappedR")) != PF_OK) return rc;
    if (braw && (rc = check_image_args(braw, sbr, rows, cols, 4, "blend (un-smoothed)")) != PF_OK) return rc;
    if (mdis && (rc = check_image_args(mdis, sd, rows, cols, 4, "MergedDis")) != PF_OK) return rc;
    if (blend && (rc = check_image_args(blend, sbl, rows, cols, 4, "Blend")) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    StitchBufs& sb = stitch_bufs(e);
    RC(sb.ensure(rows, cols));
    cudaStream_t st = sb.st;
    const size_t c4 = (size_t)cols * 4;
    const uint8_t *pL, *pR; size_t psl, psr;
    RC(stage_u8(L, sl, rows, c4, sb.L, &pL, &psl, st));
    RC(stage_u8(R, sr, rows, c4, sb.R, &pR, &psr, st));
    RC(stitch_prepare_dev(sb, pL, psl, pR, psr, rows, cols, blend != nullptr, st));
    RC(copy_out(map, sm, sb.map, (size_t)cols, rows, st));
    RC(copy_out(oL, sol, sb.oL, c4, rows, st));
    RC(copy_out(oR, sor, sb.oR, c4, rows, st));
    RC(copy_out(braw, sbr, sb.braw, c4, rows, st));
    RC(copy_out(mdis, sd, sb.mdis, c4, rows, st));
    RC(copy_out(blend, sbl, sb.blend, c4, rows, st));
    PF_CUDA(cudaStreamSynchronize(st));
    return PF_OK;
}

int pf_stitch_gather(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, const void* merged, size_t smg,
                     const void* map, size_t sm, int rows, int cols, void* out, size_t so) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(L, sl, rows, cols, 4, "ImageL")) != PF_OK) return rc;
    if ((rc = check_image_args(R, sr, rows, cols, 4, "ImageR")) != PF_OK) return rc;
    if ((rc = check_image_args(merged, smg, rows, cols, 4, "Mergedmiddle")) != PF_OK) return rc;
    if ((rc = check_image_args(out, so, rows, cols, 4, "FinalResult")) != PF_OK) return rc;
    if (!map || sm < (size_t)cols) return fail(PF_ERR_INVALID_ARGUMENT, "Map is NULL or its stride < cols");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    StitchBufs& sb = stitch_bufs(e);
    RC(sb.ensure(rows, cols));
    cudaStream_t st = sb.st;
    const size_t c4 = (size_t)cols * 4;
    const uint8_t *pL, *pR, *pM, *pMap; size_t psl, psr, psm, psmap;
    RC(stage_u8(L, sl, rows, c4, sb.L, &pL, &psl, st));
    RC(stage_u8(R, sr, rows, c4, sb.R, &pR, &psr, st));
    RC(stage_u8(merged, smg, rows, c4, sb.merged, &pM, &psm, st));
    RC(stage_u8(map, sm, rows, (size_t)cols, sb.map, &pMap, &psmap, st));
    pf::launch_stitch_gather(pL, psl, pR, psr, pM, psm, pMap, psmap, rows, cols, sb.gmap, sb.result, c4, st);
    LAUNCHED(2);
    PF_CUDA(cudaGetLastError());
    RC(copy_out(out, so, sb.result, c4, rows, st));
    PF_CUDA(cudaStreamSynchronize(st));
    return PF_OK;
}

int pf_four_input_frontend(pf_engine* e, const void* const images[4], size_t stride_in, int rows, int cols,
                           void* outL, size_t sl, void* outR, size_t sr) {
    if (!e || !images) return fail(PF_ERR_INVALID_ARGUMENT, "engine or images is NULL");
    int rc;
    for (int k = 0; k < 4; ++k)
        if ((rc = check_image_args(images[k], stride_in, rows, cols, 4, "colorImage1..4")) != PF_OK) return rc;
    if ((rc = check_image_args(outL, sl, rows, cols, 4, "colorImageL")) != PF_OK) return rc;
    if ((rc = check_image_args(outR, sr, rows, cols, 4, "colorImageR")) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    StitchBufs& sb = stitch_bufs(e);
    RC(sb.ensure(rows, cols));
    cudaStream_t st = sb.st;
    const size_t c4 = (size_t)cols * 4;
    // inputs: in place when on the device, else staged into the (otherwise idle) buffers oL, oR, merged, result
    uint8_t* stage[4] = {sb.oL, sb.oR, sb.merged, sb.result};
    const uint8_t* p[4]; size_t ps[4];
    for (int k = 0; k < 4; ++k) RC(stage_u8(images[k], stride_in, rows, c4, stage[k], &p[k], &ps[k], st));
    const bool dl = is_device_ptr(outL), dr = is_device_ptr(outR);
    uint8_t* oL = dl ? (uint8_t*)outL : sb.L;
    uint8_t* oR = dr ? (uint8_t*)outR : sb.R;
    pf::launch_four_input(p, ps, rows, cols, oL, dl ? sl : c4, oR, dr ? sr : c4, st);
    LAUNCHED(1);
    PF_CUDA(cudaGetLastError());
    if (!dl) RC(copy_out(outL, sl, sb.L, c4, rows, st));
    if (!dr) RC(copy_out(outR, sr, sb.R, c4, rows, st));
    PF_CUDA(cudaStreamSynchronize(st));
    return PF_OK;
}

int pf_stitch_iteration(pf_engine* e, const void* L, size_t sl, const void* R, size_t sr, int rows, int cols,
                        void* out, size_t so, void* blend, size_t sbl, void* merged, size_t smg, void* map, size_t sm) {
    if (!e) return fail(PF_ERR_INVALID_ARGUMENT, "engine is NULL");
    int rc;
    if ((rc = check_image_args(L, sl, rows, cols, 4, "colorImageL")) != PF_OK) return rc;
    if ((rc = check_image_args(R, sr, rows, cols, 4, "colorImageR")) != PF_OK) return rc;
    if ((rc = check_image_args(out, so, rows, cols, 4, "FinalResult")) != PF_OK) return rc;
    if ((rc = stitch_size_check(rows, cols, true)) != PF_OK) return rc;
    if (blend && (rc = check_image_args(blend, sbl, rows, cols, 4, "Blend")) != PF_OK) return rc;
    if (merged && (rc = check_image_args(merged, smg, rows, cols, 4, "Mergedmiddle")) != PF_OK) return rc;
    if (map && sm < (size_t)cols) return fail(PF_ERR_INVALID_ARGUMENT, "map stride < cols");
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    StitchBufs& sb = stitch_bufs(e);
    RC(sb.ensure(rows, cols));
    cudaStream_t st = sb.st;
    const size_t c4 = (size_t)cols * 4;
    const uint8_t *pL, *pR; size_t psl, psr;
    RC(stage_u8(L, sl, rows, c4, sb.L, &pL, &psl, st));
    RC(stage_u8(R, sr, rows, c4, sb.R, &pR, &psr, st));
    // Stitchtools::prepare (CPU/main.cpp:72-73)
    RC(stitch_prepare_dev(sb, pL, psl, pR, psr, rows, cols, true, st));
    // NovelViewGeneratorAsymmetricFlow::prepare + setBlend + generateNovelView (:82-89), everything device-resident and chained
    // on the device: the two flows start as soon as OverlappedL/R exist and run concurrently with GenerateBlend (the blend map
    // is only needed by the final warp + blend); no host synchronisation until the result is complete
    cudaEvent_t novel_done = nullptr;
    RC(novel_view_locked(e, sb.oL, c4, sb.oR, c4, sb.blend, c4, rows, cols, sb.merged, c4, nullptr, 0, nullptr, 0,
                         sb.evMasked, sb.evBlend, &novel_done));
    PF_CUDA(cudaStreamWaitEvent(st, novel_done, 0));
    // setMergedmiddle + Gather (:93-95)
    pf::launch_stitch_gather(pL, psl, pR, psr, sb.merged, c4, sb.map, (size_t)cols, rows, cols, sb.gmap, sb.result, c4, st);
    LAUNCHED(2);
    PF_CUDA(cudaGetLastError());
    RC(copy_out(out, so, sb.result, c4, rows, st));
    RC(copy_out(blend, sbl, sb.blend, c4, rows, st));
    RC(copy_out(merged, smg, sb.merged, c4, rows, st));
    RC(copy_out(map, sm, sb.map, (size_t)cols, rows, st));
    PF_CUDA(cudaStreamSynchronize(st));
    return PF_OK;
}

int pf_selftest_exact_math(int wmin, int wmax, uint64_t* out4) {
    if (!out4 || wmin < 1 || wmax < wmin) return fail(PF_ERR_INVALID_ARGUMENT, "bad selftest arguments");
    unsigned long long a = 0, b = 0, c = 0;
    int first = 0;
    if (pf::selftest_exact_math(wmin, wmax, &a, &b, &c, &first) != 0) return fail(PF_ERR_CUDA, "selftest failed to run");
    out4[0] = a; out4[1] = b; out4[2] = c; out4[3] = (uint64_t)first;
    return PF_OK;
}

int pf_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(PF_ERR_INVALID_ARGUMENT, "ptr is NULL");
    PF_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return PF_OK;
}
int pf_host_free(void* p) {
    PF_CUDA(cudaFreeHost(p));
    return PF_OK;
}

}  // extern "C"

// ---- diagnostic single-stage entry points ---------------------------------------------------------------

extern "C" {

int pf_stage_blend_smooth(pf_engine* e, float* blend, const float* mdis, int rows, int cols) {
    if (!e || !blend || !mdis) return fail(PF_ERR_INVALID_ARGUMENT, "NULL argument");
    int rc;
    if ((rc = stitch_size_check(rows, cols, true)) != PF_OK) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard g(e->device);
    StitchBufs& sb = stitch_bufs(e);
    RC(sb.ensure(rows, cols));
    const size_t n = (size_t)rows * cols * 4;
    PF_CUDA(cudaMemcpyAsync(sb.blend, blend, n, cudaMemcpyHostToDevice, sb.st));
    PF_CUDA(cudaMemcpyAsync(sb.mdis, mdis, n, cudaMemcpyHostToDevice, sb.st));
    const int nk = pf::launch_stitch_blend_smooth(sb.blend, (size_t)cols * 4, sb.mdis, (size_t)cols * 4, rows, cols, sb.scratch, sb.st);
    if (nk == 0) return fail(PF_ERR_INVALID_ARGUMENT, "blend smoothing: unsupported image shape");
    LAUNCHED(nk);
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaMemcpyAsync(blend, sb.blend, n, cudaMemcpyDeviceToHost, sb.st));
    PF_CUDA(cudaStreamSynchronize(sb.st));
    return PF_OK;
}

int pf_stage_frontend(const void* bgra, int rows, int cols, int pad, float* grey, float* alpha, int dh, int dw) {
    DevBuf in, g, a;
    const size_t n = (size_t)dh * dw * 4;
    RC(in.upload(bgra, (size_t)rows * cols * 4)); RC(g.alloc(n)); RC(a.alloc(n));
    pf::launch_frontend_resize(in.as<uint8_t>(), (size_t)cols * 4, rows, cols, pad, g.as<float>(), a.as<float>(), dh, dw, 0);
    LAUNCHED(1);
    RC(g.download(grey, n));
    return a.download(alpha, n);
}
int pf_stage_gauss5(const float* src, float* dst, int h, int w) {
    DevBuf s, d;
    const size_t n = (size_t)h * w * 4;
    RC(s.upload(src, n)); RC(d.alloc(n));
    pf::launch_gauss5(s.as<float>(), d.as<float>(), h, w, 0);
    LAUNCHED(1);
    return d.download(dst, n);
}
int pf_stage_pyr_down(const float* src, int sh, int sw, float* dst, int dh, int dw) {
    DevBuf s, d;
    RC(s.upload(src, (size_t)sh * sw * 4)); RC(d.alloc((size_t)dh * dw * 4));
    pf::PlaneSet ps{};
    ps.src[0] = s.as<float>(); ps.dst[0] = d.as<float>();
    pf::launch_pyr_down(ps, 1, sh, sw, dh, dw, 0);
    LAUNCHED(1);
    return d.download(dst, (size_t)dh * dw * 4);
}
int pf_stage_gradient(const float* I, float* G, int h, int w) {
    DevBuf s, d;
    RC(s.upload(I, (size_t)h * w * 4)); RC(d.alloc((size_t)h * w * 8));
    pf::launch_gradient(s.as<float>(), d.as<float2>(), h, w, 0);
    LAUNCHED(1);
    return d.download(G, (size_t)h * w * 8);
}
int pf_stage_blur15(const float* flow, float* dst, int h, int w, const float* alpha0, const float* alpha1) {
    DevBuf s, d, a0, a1;
    const size_t n = (size_t)h * w * 8;
    RC(s.upload(flow, n)); RC(d.alloc(n));
    pf::FlowTileMap tm;                      // dense rows: describable when w is even (row bytes a multiple of 16)
    pf::make_blur_tile_map(&tm, s.as<float2>(), h, w, w);
    if (alpha0) {
        RC(a0.upload(alpha0, n / 2)); RC(a1.upload(alpha1, n / 2));
        pf::launch_blur15_diffuse(s.as<float2>(), d.as<float2>(), h, w, w, a0.as<float>(), a1.as<float>(), &tm, 0);
    } else {
        pf::launch_blur15(s.as<float2>(), d.as<float2>(), h, w, w, &tm, 0);
    }
    LAUNCHED(1);
    return d.download(dst, n);
}
int pf_stage_median5(const float* flow, float* dst, int h, int w) {
    DevBuf s, d;
    const size_t n = (size_t)h * w * 8;
    RC(s.upload(flow, n)); RC(d.alloc(n));
    pf::FlowTileMap tm;
    pf::make_median_tile_map(&tm, s.as<float2>(), h, w, w);
    pf::launch_median5(s.as<float2>(), d.as<float2>(), h, w, w, &tm, 0);
    LAUNCHED(1);
    return d.download(dst, n);
}
int pf_stage_sweep(const float* alpha0, const float* alpha1, const float* G0, const float* G1, const float* blurred,
                   float* flow, int h, int w, int dir) {
    DevBuf a0, a1, g0, g1, g1s, bl, f, ra, bnd, tk;
    const size_t n = (size_t)h * w;
    const pf::Skew sk = pf::make_skew(w, h);
    const size_t ne = pf::skew_elems(sk);
    RC(a0.upload(alpha0, n * 4)); RC(a1.upload(alpha1, n * 4));
    RC(g0.upload(G0, n * 8)); RC(g1.upload(G1, n * 8)); RC(bl.upload(blurred, n * 8)); RC(f.upload(flow, n * 8));
    RC(g1s.alloc(ne * 8)); RC(ra.alloc(pf::sweep_rec_count(h, w) * sizeof(pf::SweepRec)));
    const size_t lines = pf::sweep2_boundary_lines(h, w);
    RC(bnd.alloc(lines * 16)); RC(tk.alloc(16));
    PF_CUDA(cudaMemset(bnd.p, 0, lines * 16)); PF_CUDA(cudaMemset(tk.p, 0, 16));
    pf::launch_skew_copy_f2(g1.as<float2>(), g1s.as<float2>(), sk, 0);
    pf::launch_sweep_prep(a0.as<float>(), a1.as<float>(), g0.as<float2>(), g1.as<float2>(), bl.as<float2>(), f.as<float2>(), w,
                          ra.as<pf::SweepRec>(), h, w, dir, 0);
    pf::Sweep2Args sa;
    sa.rec = ra.as<pf::SweepRec>(); sa.G1s = g1s.as<float2>(); sa.flow = f.as<float2>(); sa.fp = w;
    sa.s = sk; sa.g1s_last = (long long)ne - 1; sa.boundary = bnd.as<uint4>(); sa.ticket = tk.as<int>(); sa.cta_divisor = 1;
    pf::launch_sweep2(sa, dir, 0);
    LAUNCHED(3);
    return f.download(flow, n * 8);
}
int pf_stage_upsample_cubic(const float* src, int sh, int sw, float* dst, int dh, int dw) {
    DevBuf s, d;
    RC(s.upload(src, (size_t)sh * sw * 8)); RC(d.alloc((size_t)dh * dw * 8));
    pf::launch_upsample_cubic(s.as<float2>(), sh, sw, sw, d.as<float2>(), dh, dw, dw, 0);
    LAUNCHED(1);
    return d.download(dst, (size_t)dh * dw * 8);
}
int pf_stage_tail(const float* flow0, int sh, int sw, int rows, int pcols, int pad, int cols, float* out) {
    DevBuf s, d;
    RC(s.upload(flow0, (size_t)sh * sw * 8)); RC(d.alloc((size_t)rows * cols * 8));
    pf::launch_tail(s.as<float2>(), sh, sw, sw, rows, pcols, pad, cols, d.as<float2>(), (size_t)cols * 8, 0);
    LAUNCHED(1);
    return d.download(out, (size_t)rows * cols * 8);
}
int pf_stage_initial_flow(const float* I0, const float* I1, const float* alpha0, const float* alpha1, float* flow,
                          int h, int w, int hint, int dist) {
    DevBuf i0, i1, a0, a1, f, r;
    const size_t n = (size_t)h * w;
    RC(i0.upload(I0, n * 4)); RC(i1.upload(I1, n * 4)); RC(a0.upload(alpha0, n * 4)); RC(a1.upload(alpha1, n * 4));
    RC(f.alloc(n * 8)); RC(r.alloc(16));
    pf::launch_initial_flow(i0.as<float>(), i1.as<float>(), a0.as<float>(), a1.as<float>(), f.as<float2>(), w, r.as<float>(), h, w, hint, dist, 0);
    LAUNCHED(2);
    return f.download(flow, n * 8);
}

}  // extern "C"
