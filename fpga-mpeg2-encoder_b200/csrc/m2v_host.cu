// m2v_host.cu - C-ABI (include/m2venc.h) and host engine of the B200 MPEG-2 encoder.
//
// Mirrors the sequence FSM of the reference (RTL/mpeg2encoder.v:1027-1095: IDLE -> DURING -> ENDING
// -> ENDED -> IDLE) and its word-oriented output port (RTL:2924-2994), but batches whole closed GOPs
// (RTL:2645-2656: closed_gop=1; RTL:1820-1825: I-frames ignore the reference frame) so that one K1
// launch covers frame t of every GOP in the batch.
//
// One handle owns 1..8 devices (m2v_create / m2v_create_multi).  Every device has a context (streams,
// buffers, launch configuration) and a worker thread; the streaming calls deal whole-GOP batches to the
// workers round-robin and the finished bodies come back as ordered segments in pinned host memory, which
// m2v_pull / m2v_drain read in stream order.  On each device three streams overlap the host->device copy
// of batch i+1, the kernels of batch i and the device->host copy of batch i-1; nothing between the first
// K1 launch and the last K2 launch of a batch waits for the host.
// Product code: there is no CPU fallback and nothing here touches oracle/.
#include "../../include/m2venc.h"
#include "m2v_kernels.cuh"
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <stdlib.h>
#include <deque>
#include <mutex>
#include <new>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

namespace {

// M2V_TRACE=1 in the environment: the workers print a timestamped line per pipeline event (stderr)
static const bool g_trace = getenv("M2V_TRACE") != nullptr;
static double now_ms() {
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
#define TRACE(dev, ...) do { if (g_trace) { fprintf(stderr, "[m2v dev%d %10.3f ms] ", dev, now_ms()); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

struct Range {                                    // NVTX range around the host side of a phase (enqueue / wait)
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t reserve(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct PinBuf {                                   // pinned host memory: device<->host copies at PCIe rate, no bounce buffer
    uint8_t *p = nullptr; size_t cap = 0;
    bool reserve(size_t want, bool keep = false, size_t used = 0) {
        if (want <= cap) return true;
        const size_t nc = std::max(want + want / 4, (size_t)1 << 20);
        uint8_t *q = nullptr;
        if (cudaHostAlloc((void **)&q, nc, cudaHostAllocPortable) != cudaSuccess) return false;
        if (keep && used) memcpy(q, p, used);
        if (p) cudaFreeHost(p);
        p = q; cap = nc;
        return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Geom { int mbw = 0, mbh = 0, P = 0; };

// ---- per-device context ---------------------------------------------------------------------------
struct DevCtx {
    int dev = 0, VL = 3, Q = 2;
    cudaStream_t st = nullptr, st_in = nullptr, st_out = nullptr;      // kernels / host->device / device->host
    int grid_cap_i = 0, grid_cap_p = 0;
    DevBuf<uint8_t> d_in[2], d_recon0, d_recon1, d_body;
    DevBuf<uint4> d_out[2];
    DevBuf<int16_t> d_coefs;
    DevBuf<uint32_t> d_mbinfo, d_mb_bits, d_mb_code, d_mb_off, d_slice_off, d_frame_bytes;
    DevBuf<unsigned long long> d_frame_off;
    DevBuf<unsigned> d_k1ctr; unsigned k1_seq = 0;
    unsigned long long *h_total = nullptr;                              // pinned [2]: body bytes of the batch in each slot
    cudaEvent_t ev_in[2] = {}, ev_size[2] = {}, ev_enc[2] = {}, ev_out[2] = {}, ev_t[2][5] = {};
    struct SlotJob { Geom g; const uint8_t *d_in = nullptr; long F = 0, n0 = 0; } sj[2];   // what is in flight in each slot (for a re-run)
    size_t reserve_per_mb = 192;                                        // body bytes reserved per macroblock (grows after an overflow)
    long last_F = 0; int last_nmb = 0;
    long launches = 0;
    bool timing = false; float kms[5] = {0, 0, 0, 0, 0};
    char err[256] = {0};
};

#define CKC(call)                                                                         \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            snprintf(c.err, sizeof c.err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return M2V_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

int ctx_init(DevCtx &c, int dev, int VL, int Q) {
    c.dev = dev; c.VL = VL; c.Q = Q;
    cudaDeviceProp prop;
    if (cudaSetDevice(dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) return M2V_ENODEV;   // sm_100a image only
    CKC(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c.st_in, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c.st_out, cudaStreamNonBlocking));
    CKC(m2v_upload_tables(Q));
    CKC(m2v_k1_setup(VL, &c.grid_cap_i, &c.grid_cap_p));
    CKC(c.d_k1ctr.reserve(2));
    CKC(cudaMemset(c.d_k1ctr.p, 0, 2 * sizeof(unsigned)));
    CKC(cudaDeviceSynchronize());                    // table uploads and the memset ride the legacy stream: finish them before any launch
    CKC(cudaHostAlloc((void **)&c.h_total, 2 * sizeof(unsigned long long), cudaHostAllocPortable));
    for (int i = 0; i < 2; i++) {
        CKC(cudaEventCreateWithFlags(&c.ev_in[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&c.ev_size[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&c.ev_enc[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&c.ev_out[i], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; k++) for (int i = 0; i < 5; i++) CKC(cudaEventCreate(&c.ev_t[k][i]));
    return M2V_OK;
}

void ctx_destroy(DevCtx &c) {
    if (cudaSetDevice(c.dev) != cudaSuccess) return;
    for (cudaStream_t s : {c.st_in, c.st, c.st_out}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (int i = 0; i < 2; i++) for (cudaEvent_t ev : {c.ev_in[i], c.ev_size[i], c.ev_enc[i], c.ev_out[i]}) if (ev) cudaEventDestroy(ev);
    for (int k = 0; k < 2; k++) for (int i = 0; i < 5; i++) if (c.ev_t[k][i]) cudaEventDestroy(c.ev_t[k][i]);
    for (int i = 0; i < 2; i++) { c.d_in[i].release(); c.d_out[i].release(); }
    c.d_recon0.release(); c.d_recon1.release(); c.d_body.release(); c.d_coefs.release();
    c.d_mbinfo.release(); c.d_mb_bits.release(); c.d_mb_code.release(); c.d_mb_off.release(); c.d_slice_off.release();
    c.d_frame_bytes.release(); c.d_frame_off.release(); c.d_k1ctr.release();
    if (c.h_total) cudaFreeHost(c.h_total);
    c.h_total = nullptr;
}

// Enqueues the whole hot path for frames [n0, n0+F) (whole GOPs, n0 on a GOP boundary) from device memory into d_out[slot] on
// c.st and returns without waiting: K1 per frame index, K2 count, K3 scans, total -> pinned host (event ev_size), device-sized
// zeroing, K4 headers, K2 write (event ev_enc).
int encode_enqueue(DevCtx &c, const Geom &g, const uint8_t *d_in, long F, long n0, int slot) {
    M2VBatch b;
    b.g.mbw = g.mbw; b.g.mbh = g.mbh; b.g.W = g.mbw * 16; b.g.H = g.mbh * 16; b.g.nmb = g.mbw * g.mbh; b.g.P = g.P; b.g.VL = c.VL; b.g.Q = c.Q;
    b.F = F; b.n0 = n0; b.in = d_in;
    const long gop = g.P + 1, G = (F + gop - 1) / gop;
    b.CWp = ((b.g.W / 2) + 15) & ~15; b.fsz420 = (size_t)b.g.W * b.g.H + (size_t)2 * b.CWp * (b.g.H / 2);
    const size_t fsz420 = b.fsz420, nmbF = (size_t)F * b.g.nmb;
    if (G * b.g.nmb >= M2V_K1_MAX_MBS) { snprintf(c.err, sizeof c.err, "chunk too large for one K1 launch"); return M2V_EINVAL; }
    CKC(c.d_recon0.reserve(G * fsz420)); CKC(c.d_recon1.reserve(g.P ? G * fsz420 : 16));
    CKC(c.d_coefs.reserve(nmbF * 384)); CKC(c.d_mbinfo.reserve(nmbF)); CKC(c.d_mb_bits.reserve(nmbF)); CKC(c.d_mb_code.reserve((nmbF + 31) / 32 * 32 * M2V_MB_SLOT)); CKC(c.d_mb_off.reserve(nmbF));
    CKC(c.d_slice_off.reserve((size_t)F * g.mbh)); CKC(c.d_frame_bytes.reserve(F)); CKC(c.d_frame_off.reserve(F + 1));
    CKC(c.d_out[slot].reserve((nmbF * c.reserve_per_mb + 15) / 16 + 4));
    b.recon[0] = c.d_recon0.p; b.recon[1] = g.P ? c.d_recon1.p : c.d_recon0.p;
    b.coefs = c.d_coefs.p; b.mbinfo = c.d_mbinfo.p; b.mb_bits = c.d_mb_bits.p; b.mb_code = c.d_mb_code.p; b.mb_off = c.d_mb_off.p;
    b.slice_off = c.d_slice_off.p; b.frame_bytes = c.d_frame_bytes.p; b.frame_off = c.d_frame_off.p;
    b.out_words = (uint32_t *)c.d_out[slot].p; b.out_cap_words = c.d_out[slot].n * 4;
    b.k1_ctr = c.d_k1ctr.p; b.k1_grid_cap_i = c.grid_cap_i; b.k1_grid_cap_p = c.grid_cap_p;
    c.last_F = F; c.last_nmb = b.g.nmb;
    c.sj[slot].g = g; c.sj[slot].d_in = d_in; c.sj[slot].F = F; c.sj[slot].n0 = n0;
    if (!m2v_make_tmaps(b)) { snprintf(c.err, sizeof c.err, "cuTensorMapEncodeTiled failed"); return M2V_ECUDA; }

    if (c.timing) CKC(cudaEventRecord(c.ev_t[slot][0], c.st));
    {
        Range r("m2v K1 mb_encode");
        for (int t = 0; t <= g.P && t < F; t++) {                 // frame t of every GOP that has one
            const long ng = (F - t + gop - 1) / gop;
            m2v_launch_k1(b, t, ng, c.k1_seq++, c.st); c.launches++;
        }
        CKC(cudaGetLastError());                                 // a failed K1 launch must not feed stale levels to the scans
    }
    if (c.timing) CKC(cudaEventRecord(c.ev_t[slot][1], c.st));
    // (Measured and rejected: the count pass of frame t on a second stream beside K1 of frame t+1 - one K2 CTA fits next to the three
    //  K1 CTAs of an SM.  The count pass leaves the critical path (0.355 -> 0.068 ms) but K1, bound by the pipes K2 also uses, slows down
    //  by the same amount: 9.083 -> 9.067 ms per step.)
    { Range r("m2v K2 vlc count"); m2v_launch_k2(b, false, c.st); c.launches++; }
    if (c.timing) CKC(cudaEventRecord(c.ev_t[slot][2], c.st));
    {
        Range r("m2v K3 scans + K4 headers");
        m2v_launch_k3_scan(b, c.st); c.launches += 2;
        // the size goes to the host as soon as the scans are done (the host can place the body while the write pass runs) ...
        CKC(cudaMemcpyAsync(&c.h_total[slot], b.frame_off + F, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st));
        CKC(cudaEventRecord(c.ev_size[slot], c.st));
        // ... but nothing waits for it: the body is zeroed by a kernel that reads the total on the device
        m2v_launch_zero_body(b, c.st); c.launches++;
        m2v_launch_headers(b, c.st); c.launches++;
    }
    if (c.timing) CKC(cudaEventRecord(c.ev_t[slot][3], c.st));
    { Range r("m2v K2 vlc write"); m2v_launch_k2(b, true, c.st); c.launches++; }
    if (c.timing) CKC(cudaEventRecord(c.ev_t[slot][4], c.st));
    CKC(cudaGetLastError());
    CKC(cudaEventRecord(c.ev_enc[slot], c.st));
    return M2V_OK;
}

// Size of the batch in `slot` (waits for the scans only).  A body that did not fit its buffer was not written: grow the
// buffer to what the scan asked for and run the batch again (rare: the reservation is several times a typical body).
int encode_size(DevCtx &c, int slot, size_t *len) {
    const Geom g = c.sj[slot].g; const uint8_t *d_in = c.sj[slot].d_in; const long F = c.sj[slot].F, n0 = c.sj[slot].n0;
    for (int attempt = 0;; attempt++) {
        CKC(cudaEventSynchronize(c.ev_size[slot]));
        const unsigned long long total = c.h_total[slot];
        const size_t nmbF = (size_t)F * g.mbw * g.mbh;
        if (total > nmbF * 1300 + 64) { snprintf(c.err, sizeof c.err, "implausible body size %llu", total); return M2V_ECUDA; }   // > 24 bits per level
        if (M2V_BODY_WORDS(total) <= c.d_out[slot].n * 4) { *len = (size_t)total; return M2V_OK; }
        if (attempt) { snprintf(c.err, sizeof c.err, "body does not fit after regrowing"); return M2V_ECUDA; }
        CKC(cudaStreamSynchronize(c.st));
        c.reserve_per_mb = (size_t)(total / nmbF) + (size_t)(total / nmbF) / 4 + 16;
        int rc = encode_enqueue(c, g, d_in, F, n0, slot);
        if (rc) return rc;
    }
}

void add_timing(DevCtx &c, int slot) {               // after ev_enc[slot] has completed
    if (!c.timing) return;
    for (int i = 0; i < 4; i++) { float ms = 0; cudaEventElapsedTime(&ms, c.ev_t[slot][i], c.ev_t[slot][i + 1]); c.kms[i] += ms; }
    float ms = 0; cudaEventElapsedTime(&ms, c.ev_t[slot][0], c.ev_t[slot][4]); c.kms[4] += ms;
}

// ---- streaming engine -----------------------------------------------------------------------------
struct Seg {                                     // one ordered piece of the output stream
    uint8_t *p = nullptr; size_t len = 0;
    PinBuf *buf = nullptr;                       // pool buffer behind p (returned when drained); null for header / tail
    bool ready = false;
};
struct Job {
    const uint8_t *src = nullptr; long nframes = 0, n0 = 0;
    int stage = -1;                              // staging buffer to release when the copy is done (-1: caller's memory)
    size_t seg = 0;                              // index of the output segment
    Geom g;
};
struct Worker {
    DevCtx c;
    std::thread th;
    std::deque<Job> q;
    std::condition_variable cv;
    bool quit = false;
};

}  // namespace

struct m2v_encoder {
    int XL, YL, VL, Q;
    int ndev = 0;
    std::vector<Worker *> w;
    char err[256] = {0};
    // sequence state (RTL:1017-1022)
    bool busy = false, ended = false;
    Geom g;
    long frames_dealt = 0;             // absolute index of the next frame to hand to a device
    long batches_dealt = 0;
    long batch_frames = 0;             // flush threshold (whole GOPs)
    long force_batch = 0, force_chunk = 0;   // m2v_set_limits (0 = automatic)
    // staging of pushed-but-not-yet-dealt frames (m2v_push4, frame-by-frame pushes, a trailing partial GOP): pinned ring
    PinBuf stage[2]; bool stage_busy[2] = {false, false}; int cur_stage = 0;
    long staged_frames = 0; size_t px_in_frame = 0;
    // ordered output
    std::mutex mu; std::condition_variable cv_main;
    std::deque<Seg> segs; size_t seg_base = 0;       // segs[i] is segment number seg_base + i
    size_t rd_off = 0;                               // read offset inside segs.front()
    std::vector<PinBuf *> pool;                      // free pinned output buffers
    long copies_pending = 0;                         // jobs whose host->device copy has not finished (their source is still needed)
    long jobs_pending = 0;
    int async_rc = 0;
    uint8_t hdr[34]; uint8_t tail[64];
    size_t out_bytes = 0;                            // bytes of all segments dealt so far that are complete (for the tail rule)
};

namespace {

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            snprintf(e->err, sizeof e->err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return M2V_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

int clamp16(int s, int L) { return s > (1 << L) ? (1 << L) : s < 4 ? 4 : s; }   // RTL:985-991

PinBuf *pool_get(m2v_encoder *e, size_t want) {      // e->mu held
    for (size_t i = 0; i < e->pool.size(); i++)
        if (e->pool[i]->cap >= want) { PinBuf *b = e->pool[i]; e->pool.erase(e->pool.begin() + i); return b; }
    PinBuf *b = nullptr;
    if (!e->pool.empty()) { b = e->pool.back(); e->pool.pop_back(); }      // grow the largest-lived one instead of piling up buffers
    else b = new (std::nothrow) PinBuf();
    if (!b || !b->reserve(want)) { if (b) { b->release(); delete b; } return nullptr; }
    return b;
}

void fail_async(m2v_encoder *e, int rc, const char *msg) {                // e->mu held
    if (!e->async_rc) { e->async_rc = rc; snprintf(e->err, sizeof e->err, "%s", msg); }
}

// Worker thread of one device.  Per batch k (slot = k & 1): copy in on st_in, kernels on st, and - one batch behind, so that
// the next batch's copy and kernels are already queued - wait for the size, take a pinned segment, copy the body out on st_out.
void worker_main(m2v_encoder *e, Worker *w) {
    DevCtx &c = w->c;
    cudaSetDevice(c.dev);
    struct Pending { bool valid = false; Job j; int slot = 0; } pend;
    long k = 0;
    auto finalize = [&](Pending &p) {
        if (!p.valid) return;
        p.valid = false;
        size_t len = 0;
        int rc;
        { Range r("m2v wait size"); rc = encode_size(c, p.slot, &len); }
        TRACE(c.dev, "batch n0=%ld: scans done, body %zu bytes", p.j.n0, len);
        PinBuf *buf = nullptr;
        if (!rc) {
            std::lock_guard<std::mutex> lk(e->mu);
            buf = pool_get(e, std::max(len, (size_t)64));
            if (!buf) rc = M2V_ENOMEM;
        }
        if (!rc) {
            Range r("m2v D2H body");
            cudaError_t ce = cudaStreamWaitEvent(c.st_out, c.ev_enc[p.slot], 0);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(buf->p, c.d_out[p.slot].p, len, cudaMemcpyDeviceToHost, c.st_out);
            if (ce == cudaSuccess) ce = cudaEventRecord(c.ev_out[p.slot], c.st_out);
            if (ce == cudaSuccess) ce = cudaEventSynchronize(c.ev_out[p.slot]);
            if (ce != cudaSuccess) { snprintf(c.err, sizeof c.err, "device->host copy of a body: %s", cudaGetErrorString(ce)); rc = M2V_ECUDA; }
        }
        TRACE(c.dev, "batch n0=%ld: body in host memory", p.j.n0);
        std::lock_guard<std::mutex> lk(e->mu);
        Seg &s = e->segs[p.j.seg - e->seg_base];
        if (rc) { fail_async(e, rc, c.err[0] ? c.err : "worker failed"); if (buf) e->pool.push_back(buf); s.len = 0; }
        else { s.p = buf->p; s.len = len; s.buf = buf; }
        s.ready = true;
        e->jobs_pending--;
        e->cv_main.notify_all();
    };
    // the host->device copy of a batch is over: its source (the caller's memory or a staging buffer) is free again
    struct CopyWait { bool valid = false; Job j; int slot = 0; int rc = 0; } cw;
    auto complete_copy = [&](CopyWait &x) {
        if (!x.valid) return;
        x.valid = false;
        cudaError_t ce = cudaEventSynchronize(c.ev_in[x.slot]);
        TRACE(c.dev, "batch n0=%ld: host->device copy done", x.j.n0);
        if (ce != cudaSuccess && !x.rc) { snprintf(c.err, sizeof c.err, "host->device copy of a batch: %s", cudaGetErrorString(ce)); x.rc = M2V_ECUDA; }
        std::lock_guard<std::mutex> lk(e->mu);
        if (x.j.stage >= 0) e->stage_busy[x.j.stage] = false;
        e->copies_pending--;
        if (x.rc) {
            fail_async(e, x.rc, c.err[0] ? c.err : "worker failed");
            Seg &sg = e->segs[x.j.seg - e->seg_base]; sg.len = 0; sg.ready = true; e->jobs_pending--;
        }
        e->cv_main.notify_all();
    };
    for (;;) {
        Job j; bool have = false;
        {
            std::lock_guard<std::mutex> lk(e->mu);
            if (!w->q.empty()) { j = w->q.front(); w->q.pop_front(); have = true; }
        }
        if (!have) {                                              // nothing to overlap with: finish what is in flight, then sleep
            complete_copy(cw);
            finalize(pend);
            std::unique_lock<std::mutex> lk(e->mu);
            w->cv.wait(lk, [&] { return w->quit || !w->q.empty(); });
            if (w->q.empty()) break;
            continue;
        }
        const int slot = (int)(k++ & 1);
        const size_t fsz = (size_t)j.g.mbw * j.g.mbh * 768, bytes = fsz * j.nframes;
        TRACE(c.dev, "batch n0=%ld: %ld frames, %zu bytes, slot %d: queueing copy and kernels", j.n0, j.nframes, bytes, slot);
        int rc = M2V_OK;
        {
            // d_in[slot] and d_out[slot] were last used by batch k-2, which finalize() saw through to its device->host copy
            Range r("m2v H2D frames");
            cudaError_t ce = c.d_in[slot].reserve(bytes);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(c.d_in[slot].p, j.src, bytes, cudaMemcpyHostToDevice, c.st_in);
            if (ce == cudaSuccess) ce = cudaEventRecord(c.ev_in[slot], c.st_in);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(c.st, c.ev_in[slot], 0);
            if (ce != cudaSuccess) { snprintf(c.err, sizeof c.err, "host->device copy of a batch: %s", cudaGetErrorString(ce)); rc = M2V_ECUDA; }
        }
        if (!rc) rc = encode_enqueue(c, j.g, c.d_in[slot].p, j.nframes, j.n0, slot);
        TRACE(c.dev, "batch n0=%ld: queued", j.n0);
        // batch k's copy and kernels are queued: now the previous batch's copy-done signal and its way out
        complete_copy(cw);
        finalize(pend);
        cw.valid = true; cw.j = j; cw.slot = slot; cw.rc = rc;
        if (!rc) { pend.valid = true; pend.j = j; pend.slot = slot; }
    }
    complete_copy(cw);
    finalize(pend);
    cudaStreamSynchronize(c.st_in); cudaStreamSynchronize(c.st); cudaStreamSynchronize(c.st_out);
}

int create_on(const int *devices, int ndev, int XL, int YL, int VL, int Q, m2v_encoder **out) {
    if (!out) return M2V_EINVAL;
    *out = nullptr;
    if (XL < 4 || XL > 7 || YL < 4 || YL > 7 || VL < 1 || VL > 3 || Q < 1 || Q > 4 || ndev < 1 || ndev > 8) return M2V_EINVAL;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return M2V_ENODEV;
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return M2V_ENODEV;
    m2v_encoder *e = new (std::nothrow) m2v_encoder();
    if (!e) return M2V_ENOMEM;
    e->XL = XL; e->YL = YL; e->VL = VL; e->Q = Q;
    int rc = M2V_OK;
    for (int i = 0; i < ndev && !rc; i++) {
        const int dev = devices ? devices[i] : cur;
        if (dev < 0 || dev >= have) { rc = M2V_ENODEV; break; }
        Worker *w = new (std::nothrow) Worker();
        if (!w) { rc = M2V_ENOMEM; break; }
        e->w.push_back(w); e->ndev++;
        rc = ctx_init(w->c, dev, VL, Q);
    }
    cudaSetDevice(cur);
    if (rc) { m2v_destroy(e); return rc; }
    for (Worker *w : e->w) w->th = std::thread(worker_main, e, w);
    *out = e;
    return M2V_OK;
}

}  // namespace

extern "C" int m2v_create(int XL, int YL, int VL, int Q, m2v_encoder **out) { return create_on(nullptr, 1, XL, YL, VL, Q, out); }

extern "C" int m2v_create_multi(int ndev, int XL, int YL, int VL, int Q, m2v_encoder **out) {
    if (ndev < 1 || ndev > 8) return M2V_EINVAL;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return M2V_ENODEV;
    if (ndev > have) return M2V_ENODEV;
    int devs[8];
    for (int i = 0; i < ndev; i++) devs[i] = i;
    return create_on(devs, ndev, XL, YL, VL, Q, out);
}

extern "C" int m2v_device_count(const m2v_encoder *e) { return e ? e->ndev : 0; }

extern "C" void m2v_destroy(m2v_encoder *e) {
    if (!e) return;
    int cur = 0; cudaGetDevice(&cur);
    {
        std::lock_guard<std::mutex> lk(e->mu);
        for (Worker *w : e->w) { w->quit = true; w->cv.notify_all(); }
    }
    for (Worker *w : e->w) if (w->th.joinable()) w->th.join();
    for (Worker *w : e->w) { ctx_destroy(w->c); delete w; }
    for (Seg &s : e->segs) if (s.buf) e->pool.push_back(s.buf);
    for (PinBuf *b : e->pool) { b->release(); delete b; }
    e->stage[0].release(); e->stage[1].release();
    cudaSetDevice(cur);
    delete e;
}

extern "C" const char *m2v_last_error(const m2v_encoder *e) { return e ? e->err : "null handle"; }
extern "C" long m2v_launch_count(const m2v_encoder *e) {
    long n = 0;
    if (e) for (Worker *w : e->w) n += w->c.launches;
    return n;
}
extern "C" int m2v_set_limits(m2v_encoder *e, long batch_frames, long chunk_frames_) {
    if (!e || batch_frames < 0 || chunk_frames_ < 0) return M2V_EINVAL;
    e->force_batch = batch_frames; e->force_chunk = chunk_frames_;
    return M2V_OK;
}
extern "C" int m2v_set_body_reserve(m2v_encoder *e, long bytes_per_macroblock) {
    if (!e || bytes_per_macroblock < 1) return M2V_EINVAL;
    for (Worker *w : e->w) w->c.reserve_per_mb = (size_t)bytes_per_macroblock;
    return M2V_OK;
}
extern "C" int m2v_set_timing(m2v_encoder *e, int en) {
    if (!e) return M2V_EINVAL;
    for (Worker *w : e->w) w->c.timing = en != 0;
    return M2V_OK;
}
extern "C" int m2v_kernel_ms(const m2v_encoder *e, float ms[5]) {
    if (!e || !ms) return M2V_EINVAL;
    memcpy(ms, e->w[0]->c.kms, sizeof e->w[0]->c.kms);
    return M2V_OK;
}
// A refused request is reported through the return value and must not linger as the runtime's "last error": the launch checks of
// the encoder read that state.
extern "C" void *m2v_alloc_host(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return p;
    cudaGetLastError();
    return nullptr;
}
extern "C" void m2v_free_host(void *p) { if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError(); }
extern "C" int m2v_register_host(void *p, size_t bytes) {
    if (p && bytes && cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess) return M2V_OK;
    cudaGetLastError();
    return p && bytes ? M2V_ECUDA : M2V_EINVAL;
}
extern "C" int m2v_unregister_host(void *p) {
    if (p && cudaHostUnregister(p) == cudaSuccess) return M2V_OK;
    cudaGetLastError();
    return p ? M2V_ECUDA : M2V_EINVAL;
}

// ---- framing helpers ---------------------------------------------------------------------------
namespace {
struct BitW {
    uint8_t *b; size_t pos = 0; uint64_t acc = 0; int n = 0;
    void put(uint32_t c, int len) { acc = (acc << len) | (c & ((len >= 32) ? 0xFFFFFFFFu : ((1u << len) - 1))); n += len; while (n >= 8) { b[pos++] = (uint8_t)(acc >> (n - 8)); n -= 8; } }
    void align() { if (n) put(0, 8 - n); }
};
}

extern "C" int m2v_sequence_header(int mbw, int mbh, uint8_t out[34]) {      // RTL:2596-2617
    if (!out || mbw < 4 || mbw > 128 || mbh < 4 || mbh > 128) return M2V_EINVAL;
    const uint32_t W = mbw * 16, H = mbh * 16;
    BitW w{out};
    w.put(0x000001, 24); w.put(0xB3, 8); w.put(W, 12); w.put(H, 12);
    w.put(0x1209c4, 24); w.put(0x200000, 24); w.put(0x0001B5, 24); w.put(0x144200, 24);
    w.put(0x010000, 24); w.put(0x000001, 24); w.put(0xB52305, 24); w.put(0x0505, 16);
    w.put(W, 14); w.put(1, 1); w.put(H, 14);
    w.align();
    return w.pos == 34 ? M2V_OK : M2V_EINVAL;
}

extern "C" int m2v_finish_stream(uint8_t *buf, size_t len, size_t cap, size_t *total) {   // RTL:2621-2628, 2932-2937
    if (!buf || !total) return M2V_EINVAL;
    const size_t n = len + 4, t = 32 * (n / 32 + 1);
    if (t > cap) return M2V_ESPACE;
    buf[len] = 0; buf[len + 1] = 0; buf[len + 2] = 1; buf[len + 3] = 0xB7;
    memset(buf + n, 0, t - n);
    *total = t;
    return M2V_OK;
}

// ---- device-resident bulk path (synchronous, on the handle's first device, on the calling thread) ------------------
namespace {
long chunk_frames(const m2v_encoder *e, int mbw, int mbh, int P) {
    // bound the level buffer (768 B per macroblock) to ~6 GiB per chunk, whole GOPs
    const size_t per_frame = (size_t)mbw * mbh * 768;
    long f = (long)((6ull << 30) / per_frame);
    if (e->force_chunk > 0) f = std::min(f, e->force_chunk);
    const long gop = P + 1;
    f = std::max(gop, f / gop * gop);
    // and stay below the index range of one K1 launch
    const long maxg = (M2V_K1_MAX_MBS - 1) / ((long)mbw * mbh);
    f = std::min(f, std::max(1l, maxg) * gop);
    return f;
}
int streaming_active(m2v_encoder *e) {
    std::lock_guard<std::mutex> lk(e->mu);
    return e->jobs_pending > 0;
}
}

extern "C" int m2v_encode_gops_device(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0,
                                      const uint8_t **d_body, size_t *body_len) {
    if (!e) return M2V_EINVAL;
    if (!d_in || F <= 0 || mbw < 4 || mbw > (1 << e->XL) || mbh < 4 || mbh > (1 << e->YL) || P < 0 || P > 255 || n0 < 0 || n0 % (P + 1) != 0 ||
        !d_body || !body_len) { snprintf(e->err, sizeof e->err, "encode_gops: bad argument"); return M2V_EINVAL; }
    if (((uintptr_t)d_in & 15) != 0) { snprintf(e->err, sizeof e->err, "encode_gops: the frames must be 16-byte aligned (TMA)"); return M2V_EINVAL; }
    if (streaming_active(e)) { snprintf(e->err, sizeof e->err, "encode_gops while streamed batches are in flight"); return M2V_ESTATE; }
    DevCtx &c = e->w[0]->c;
    CK(cudaSetDevice(c.dev));
    if (c.timing) memset(c.kms, 0, sizeof c.kms);
    Geom g; g.mbw = mbw; g.mbh = mbh; g.P = P;
    const long cf = chunk_frames(e, mbw, mbh, P);
    const size_t fsz = (size_t)mbw * mbh * 256 * 3;
    auto fail = [&](int rc) { snprintf(e->err, sizeof e->err, "%s", c.err); return rc; };
    if (F <= cf) {                                                 // common case: one chunk, no extra copy
        size_t len = 0;
        int rc = encode_enqueue(c, g, d_in, F, n0, 0);
        if (!rc) rc = encode_size(c, 0, &len);
        if (rc) return fail(rc);
        CK(cudaEventSynchronize(c.ev_enc[0]));
        add_timing(c, 0);
        *d_body = (const uint8_t *)c.d_out[0].p; *body_len = len;
        return M2V_OK;
    }
    size_t tot = 0;
    for (long f0 = 0; f0 < F; f0 += cf) {
        const long fc = std::min(cf, F - f0);
        size_t len = 0;
        int rc = encode_enqueue(c, g, d_in + (size_t)f0 * fsz, fc, n0 + f0, 0);
        if (!rc) rc = encode_size(c, 0, &len);
        if (rc) return fail(rc);
        if (tot + len > c.d_body.n) {                              // grow, keeping what is there
            DevBuf<uint8_t> nb;
            CK(nb.reserve(std::max((tot + len) * 2, (size_t)1 << 20)));
            if (tot) { CK(cudaMemcpyAsync(nb.p, c.d_body.p, tot, cudaMemcpyDeviceToDevice, c.st)); CK(cudaStreamSynchronize(c.st)); }
            c.d_body.release(); c.d_body = nb;
        }
        CK(cudaMemcpyAsync(c.d_body.p + tot, c.d_out[0].p, len, cudaMemcpyDeviceToDevice, c.st));
        CK(cudaStreamSynchronize(c.st));
        add_timing(c, 0);
        tot += len;
    }
    *d_body = c.d_body.p; *body_len = tot;
    return M2V_OK;
}

extern "C" int m2v_encode_gops_host(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0,
                                    uint8_t *h_body, size_t cap, size_t *body_len) {
    const uint8_t *d = nullptr; size_t len = 0;
    int rc = m2v_encode_gops_device(e, mbw, mbh, P, d_in, F, n0, &d, &len);
    if (rc) return rc;
    if (body_len) *body_len = len;
    if (!h_body || len > cap) { snprintf(e->err, sizeof e->err, "encode_gops_host: need %zu bytes, have %zu", len, cap); return M2V_ESPACE; }
    DevCtx &c = e->w[0]->c;
    CK(cudaMemcpyAsync(h_body, d, len, cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    return M2V_OK;
}

// ---- asynchronous form of the same (GOP-sharded callers that overlap the body's way to the host with the next chunk) ----
extern "C" int m2v_gops_submit(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0, int slot) {
    if (!e) return M2V_EINVAL;
    if (!d_in || F <= 0 || mbw < 4 || mbw > (1 << e->XL) || mbh < 4 || mbh > (1 << e->YL) || P < 0 || P > 255 || n0 < 0 || n0 % (P + 1) != 0 ||
        slot < 0 || slot > 1 || ((uintptr_t)d_in & 15) != 0 || F > chunk_frames(e, mbw, mbh, P)) { snprintf(e->err, sizeof e->err, "gops_submit: bad argument"); return M2V_EINVAL; }
    if (streaming_active(e)) { snprintf(e->err, sizeof e->err, "gops_submit while streamed batches are in flight"); return M2V_ESTATE; }
    DevCtx &c = e->w[0]->c;
    CK(cudaSetDevice(c.dev));
    CK(cudaStreamWaitEvent(c.st, c.ev_out[slot], 0));              // an earlier m2v_gops_fetch from this slot has finished reading it
    Geom g; g.mbw = mbw; g.mbh = mbh; g.P = P;
    int rc = encode_enqueue(c, g, d_in, F, n0, slot);
    if (rc) snprintf(e->err, sizeof e->err, "%s", c.err);
    return rc;
}
extern "C" int m2v_gops_size(m2v_encoder *e, int slot, size_t *body_len) {
    if (!e || !body_len || slot < 0 || slot > 1) return M2V_EINVAL;
    DevCtx &c = e->w[0]->c;
    CK(cudaSetDevice(c.dev));
    int rc = encode_size(c, slot, body_len);
    if (rc) snprintf(e->err, sizeof e->err, "%s", c.err);
    return rc;
}
extern "C" int m2v_gops_fetch(m2v_encoder *e, int slot, uint8_t *h_dst, size_t len) {
    if (!e || !h_dst || slot < 0 || slot > 1) return M2V_EINVAL;
    DevCtx &c = e->w[0]->c;
    CK(cudaSetDevice(c.dev));
    CK(cudaStreamWaitEvent(c.st_out, c.ev_enc[slot], 0));
    if (len) CK(cudaMemcpyAsync(h_dst, c.d_out[slot].p, len, cudaMemcpyDeviceToHost, c.st_out));
    CK(cudaEventRecord(c.ev_out[slot], c.st_out));
    return M2V_OK;
}
extern "C" int m2v_gops_wait(m2v_encoder *e, int slot) {
    if (!e || slot < 0 || slot > 1) return M2V_EINVAL;
    DevCtx &c = e->w[0]->c;
    CK(cudaSetDevice(c.dev));
    CK(cudaEventSynchronize(c.ev_enc[slot]));
    CK(cudaEventSynchronize(c.ev_out[slot]));
    add_timing(c, slot);
    return M2V_OK;
}
extern "C" int m2v_gops_body(m2v_encoder *e, int slot, const uint8_t **d_body) {
    if (!e || !d_body || slot < 0 || slot > 1) return M2V_EINVAL;
    *d_body = (const uint8_t *)e->w[0]->c.d_out[slot].p;
    return M2V_OK;
}

extern "C" int m2v_debug_copy(m2v_encoder *e, uint32_t *mbinfo, int16_t *coefs, long count) {
    if (!e) return M2V_EINVAL;
    DevCtx &c = e->w[0]->c;
    if (count > c.last_F * c.last_nmb) return M2V_EINVAL;
    CK(cudaSetDevice(c.dev));
    if (mbinfo) CK(cudaMemcpyAsync(mbinfo, c.d_mbinfo.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c.st));
    if (coefs) CK(cudaMemcpyAsync(coefs, c.d_coefs.p, (size_t)count * 768, cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    return M2V_OK;
}

// ---- streaming contract ------------------------------------------------------------------------
namespace {

int check_async(m2v_encoder *e) {
    std::lock_guard<std::mutex> lk(e->mu);
    return e->async_rc;
}

void recycle_output(m2v_encoder *e) {                  // e->mu held, no job in flight
    for (Seg &s : e->segs) if (s.buf) e->pool.push_back(s.buf);
    e->seg_base += e->segs.size();
    e->segs.clear(); e->rd_off = 0;
}

// hand frames [src, src + nframes) to the next device; the caller guarantees src stays valid until copies_pending drops
void deal(m2v_encoder *e, const uint8_t *src, long nframes, int stage) {
    std::lock_guard<std::mutex> lk(e->mu);
    Job j; j.src = src; j.nframes = nframes; j.n0 = e->frames_dealt; j.stage = stage; j.g = e->g;
    j.seg = e->seg_base + e->segs.size();
    e->segs.emplace_back();
    Worker *w = e->w[e->batches_dealt % e->ndev];
    e->batches_dealt++; e->frames_dealt += nframes;
    e->copies_pending++; e->jobs_pending++;
    if (stage >= 0) e->stage_busy[stage] = true;
    w->q.push_back(j);
    w->cv.notify_one();
}

int start_if_idle(m2v_encoder *e) {
    if (e->busy) return M2V_OK;
    if (e->g.mbw == 0 || e->ended) { snprintf(e->err, sizeof e->err, "push before begin"); return M2V_ESTATE; }
    e->busy = true;
    int rc = m2v_sequence_header(e->g.mbw, e->g.mbh, e->hdr);
    std::lock_guard<std::mutex> lk(e->mu);
    Seg s; s.p = e->hdr; s.len = 34; s.ready = true;
    e->segs.push_back(s);
    return rc;
}

// the staged frames become a batch; the other staging buffer takes over (waiting for its previous copy if need be)
int flush_staged(m2v_encoder *e) {
    if (e->staged_frames == 0) return M2V_OK;
    deal(e, e->stage[e->cur_stage].p, e->staged_frames, e->cur_stage);
    e->staged_frames = 0;
    e->cur_stage ^= 1;
    std::unique_lock<std::mutex> lk(e->mu);
    e->cv_main.wait(lk, [&] { return !e->stage_busy[e->cur_stage] || e->async_rc; });
    return e->async_rc;
}

uint8_t *stage_room(m2v_encoder *e, size_t bytes_needed) {
    PinBuf &s = e->stage[e->cur_stage];
    const size_t fsz = (size_t)e->g.mbw * e->g.mbh * 768;
    const size_t used = (size_t)e->staged_frames * fsz + (e->px_in_frame ? fsz : 0);
    // one allocation for the whole batch: pinning memory is slow (tens of ms per 100 MB), growing in steps would pay it repeatedly
    if (!s.reserve(std::max(bytes_needed, (size_t)e->batch_frames * fsz), true, used)) return nullptr;
    return s.p;
}

}  // namespace

extern "C" int m2v_begin(m2v_encoder *e, int xs, int ys, int P, int *mbw, int *mbh) {
    if (!e) return M2V_EINVAL;
    if (e->busy) { snprintf(e->err, sizeof e->err, "begin while o_sequence_busy"); return M2V_ESTATE; }
    if (P < 0 || P > 255) return M2V_EINVAL;
    e->g.mbw = clamp16(xs, e->XL); e->g.mbh = clamp16(ys, e->YL); e->g.P = P;
    if (mbw) *mbw = e->g.mbw; if (mbh) *mbh = e->g.mbh;
    e->frames_dealt = 0; e->batches_dealt = 0; e->staged_frames = 0; e->px_in_frame = 0; e->ended = false;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        recycle_output(e);
        e->async_rc = 0; e->out_bytes = 0;
    }
    // flush threshold: enough GOPs that one K1 step has >= 16k macroblocks (a smaller batch shortens the
    // un-overlapped first H2D copy and last encode of the pipeline), staging <= 1 GiB
    const long gop = P + 1, nmb = (long)e->g.mbw * e->g.mbh;
    long g = (16384 + nmb - 1) / nmb;
    const size_t fsz = (size_t)nmb * 768;
    while ((size_t)g * gop * fsz < ((size_t)64 << 20)) g++;       // and >= 64 MiB of input, so per-batch overheads stay small
    while (g > 1 && (size_t)g * gop * fsz > ((size_t)1 << 30)) g--;
    e->batch_frames = std::min(g * gop, chunk_frames(e, e->g.mbw, e->g.mbh, P));
    if (e->force_batch > 0) e->batch_frames = std::max(gop, std::min(e->batch_frames, e->force_batch / gop * gop));
    // the RTL arms on the first i_en (RTL:1060-1065); the header is emitted then
    return M2V_OK;
}

extern "C" int m2v_push4(m2v_encoder *e, const uint8_t Y[4], const uint8_t U[4], const uint8_t V[4]) {
    if (!e || !Y || !U || !V) return M2V_EINVAL;
    if (e->ended) { snprintf(e->err, sizeof e->err, "push after stop"); return M2V_ESTATE; }
    int rc = start_if_idle(e); if (rc) return rc;
    const size_t ysz = (size_t)e->g.mbw * e->g.mbh * 256, fsz = ysz * 3;
    const size_t base = (size_t)e->staged_frames * fsz;
    uint8_t *s = e->px_in_frame ? e->stage[e->cur_stage].p : stage_room(e, base + fsz);
    if (!s) { snprintf(e->err, sizeof e->err, "cudaHostAlloc of the staging buffer failed"); return M2V_ENOMEM; }
    uint8_t *f = s + base;
    memcpy(f + e->px_in_frame, Y, 4); memcpy(f + ysz + e->px_in_frame, U, 4); memcpy(f + 2 * ysz + e->px_in_frame, V, 4);
    e->px_in_frame += 4;
    if (e->px_in_frame == ysz) {
        e->px_in_frame = 0; e->staged_frames++;
        if (e->staged_frames >= e->batch_frames) return flush_staged(e);
    }
    return M2V_OK;
}

extern "C" int m2v_push_frames(m2v_encoder *e, const uint8_t *yuv, long nframes) {
    if (!e || !yuv || nframes < 0) return M2V_EINVAL;
    if (e->ended) { snprintf(e->err, sizeof e->err, "push after stop"); return M2V_ESTATE; }
    if (nframes == 0) return M2V_OK;                              // no pixel, no i_en: the sequence is not armed (RTL:1060-1065)
    int rc = check_async(e); if (rc) return rc;
    rc = start_if_idle(e); if (rc) return rc;
    if (e->px_in_frame) { snprintf(e->err, sizeof e->err, "push_frames inside a frame"); return M2V_ESTATE; }
    const size_t fsz = (size_t)e->g.mbw * e->g.mbh * 768;
    const long gopf = e->g.P + 1;
    // whole batches go straight from the caller's buffer to the devices (no host staging copy); only a trailing partial GOP
    // is staged.  The call returns when every copy from the caller's memory has finished; the kernels may still be running.
    if (e->staged_frames == 0 && nframes >= gopf) {
        const long direct = nframes / gopf * gopf;
        // several devices: cut the push into at least one batch per device
        long bf = e->batch_frames;
        if (e->ndev > 1) bf = std::max(gopf, std::min(bf, (direct / gopf + e->ndev - 1) / e->ndev * gopf));
        // the kernels of the LAST batch of a push overlap nothing (the caller comes back for more, or stops): it is cut in two
        // halves so that what is left exposed is half as long
        bool halved = false;
        for (long f0 = 0; f0 < direct;) {
            long n = std::min(bf, direct - f0);
            if (!halved && n == direct - f0 && n >= 2 * gopf) { n = (n / gopf + 1) / 2 * gopf; halved = true; }
            deal(e, yuv + (size_t)f0 * fsz, n, -1);
            f0 += n;
        }
        {
            Range r("m2v push_frames: wait for the copies");
            std::unique_lock<std::mutex> lk(e->mu);
            e->cv_main.wait(lk, [&] { return e->copies_pending == 0; });
            if (e->async_rc) return e->async_rc;
        }
        yuv += (size_t)direct * fsz; nframes -= direct;
    }
    while (nframes > 0) {
        const long take = std::min(nframes, e->batch_frames - e->staged_frames);
        const size_t base = (size_t)e->staged_frames * fsz;
        uint8_t *s = stage_room(e, base + take * fsz);
        if (!s) { snprintf(e->err, sizeof e->err, "cudaHostAlloc of the staging buffer failed"); return M2V_ENOMEM; }
        memcpy(s + base, yuv, take * fsz);
        e->staged_frames += take; yuv += take * fsz; nframes -= take;
        if (e->staged_frames >= e->batch_frames) { rc = flush_staged(e); if (rc) return rc; }
    }
    return M2V_OK;
}

extern "C" int m2v_stop(m2v_encoder *e) {
    if (!e) return M2V_EINVAL;
    if (!e->busy || e->ended) return M2V_OK;                       // stop while idle is ignored (RTL:1090)
    if (e->px_in_frame) {                                          // pad the unfinished frame (RTL:1036-1037, 1049-1056)
        const size_t ysz = (size_t)e->g.mbw * e->g.mbh * 256, fsz = ysz * 3;
        uint8_t *f = e->stage[e->cur_stage].p + (size_t)e->staged_frames * fsz;
        memset(f + e->px_in_frame, 0, ysz - e->px_in_frame);
        memset(f + ysz + e->px_in_frame, 0x80, ysz - e->px_in_frame);
        memset(f + 2 * ysz + e->px_in_frame, 0x80, ysz - e->px_in_frame);
        e->px_in_frame = 0; e->staged_frames++;
    }
    if (e->staged_frames) { deal(e, e->stage[e->cur_stage].p, e->staged_frames, e->cur_stage); e->staged_frames = 0; e->cur_stage ^= 1; }
    std::unique_lock<std::mutex> lk(e->mu);
    { Range r("m2v stop: wait for the batches"); e->cv_main.wait(lk, [&] { return e->jobs_pending == 0; }); }
    if (e->async_rc) return e->async_rc;
    size_t len = 0;
    for (const Seg &s : e->segs) len += s.len;
    len += e->out_bytes;                                           // bytes already drained (segments popped)
    const size_t n = len + 4, t = 32 * (n / 32 + 1);               // RTL:2621-2628, 2932-2937: end code, "always one more word"
    memset(e->tail, 0, sizeof e->tail);
    e->tail[2] = 1; e->tail[3] = 0xB7;
    Seg s; s.p = e->tail; s.len = t - len; s.ready = true;
    e->segs.push_back(s);
    e->ended = true;
    return M2V_OK;
}

extern "C" int m2v_busy(const m2v_encoder *e) { return e && e->busy; }

namespace {
// segments -> caller's buffer.  A long stretch (tens of MB after a big push) is copied by a few threads: one core moves
// ~10 GB/s, and this copy sits on the critical path of the end-to-end time after the last kernel.
void copy_out(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t kMin = (size_t)1 << 20;
    if (n < 2 * kMin) { memcpy(dst, src, n); return; }
    const int parts = (int)std::min<size_t>(8, n / kMin);
    const size_t per = (n / parts + 63) & ~(size_t)63;
    std::thread th[7];
    for (int i = 1; i < parts; i++) {
        const size_t o = (size_t)i * per, len = std::min(per, n - o);
        th[i - 1] = std::thread([=] { memcpy(dst + o, src + o, len); });
    }
    memcpy(dst, src, std::min(per, n));
    for (int i = 1; i < parts; i++) th[i - 1].join();
}
}

extern "C" int m2v_drain(m2v_encoder *e, uint8_t *dst, size_t cap, size_t *n, int *last) {
    if (!e || !dst || !n) return M2V_EINVAL;
    *n = 0; if (last) *last = 0;
    std::unique_lock<std::mutex> lk(e->mu);
    if (e->async_rc) return e->async_rc;
    // bytes available in order: ready segments from the front
    size_t avail = 0;
    for (const Seg &s : e->segs) { if (!s.ready) break; avail += s.len; }
    avail -= std::min(avail, e->rd_off);
    size_t take = std::min(avail, cap) / 32 * 32;                  // whole 32-byte words only (RTL:2961-2994)
    size_t done = 0;
    while (done < take) {
        Seg &s = e->segs.front();
        const size_t k = std::min(take - done, s.len - e->rd_off);
        if (k) { lk.unlock(); copy_out(dst + done, s.p + e->rd_off, k); lk.lock(); }
        done += k; e->rd_off += k;
        if (e->rd_off == s.len) {
            e->out_bytes += s.len;
            if (s.buf) e->pool.push_back(s.buf);
            e->segs.pop_front(); e->seg_base++; e->rd_off = 0;
        }
    }
    while (!e->segs.empty() && e->segs.front().ready && e->rd_off == e->segs.front().len) {   // empty segments
        e->out_bytes += e->segs.front().len;
        if (e->segs.front().buf) e->pool.push_back(e->segs.front().buf);
        e->segs.pop_front(); e->seg_base++; e->rd_off = 0;
    }
    *n = take;
    const bool fin = e->ended && e->segs.empty();
    if (last) *last = fin && take > 0;
    if (fin) { e->busy = false; e->ended = false; e->g.mbw = 0; e->out_bytes = 0; }   // back to IDLE (RTL:1045-1047)
    return M2V_OK;
}

extern "C" int m2v_pull(m2v_encoder *e, uint8_t out[32], int *last) {
    size_t n = 0; int l = 0;
    int rc = m2v_drain(e, out, 32, &n, &l);
    if (rc) return rc;
    if (last) *last = l;
    return n == 32 ? 1 : 0;
}
