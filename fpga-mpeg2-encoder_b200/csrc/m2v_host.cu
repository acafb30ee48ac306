// m2v_host.cu - C-ABI (include/m2venc.h) and host state machine of the B200 MPEG-2 encoder.
//
// Mirrors the sequence FSM of the reference (RTL/mpeg2encoder.v:1027-1095: IDLE -> DURING -> ENDING
// -> ENDED -> IDLE) and its word-oriented output port (RTL:2924-2994), but batches whole closed GOPs
// (RTL:2645-2656: closed_gop=1; RTL:1820-1825: I-frames ignore the reference frame) so that one K1
// launch covers frame t of every GOP in the batch.  Product code: there is no CPU fallback and
// nothing here touches oracle/.
#include "../../include/m2venc.h"
#include "m2v_kernels.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <new>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            snprintf(e->err, sizeof e->err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return M2V_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

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

// Output word queue in PINNED host memory: the device->host copies of the bodies run at PCIe rate instead of
// going through the driver's bounce buffer for pageable memory.
struct PinnedQ {
    uint8_t *p = nullptr; size_t n = 0, cap = 0;
    size_t size() const { return n; }
    uint8_t *data() { return p; }
    void clear() { n = 0; }
    bool resize(size_t want) {
        if (want > cap) {
            size_t nc = std::max(want + want / 2, (size_t)1 << 20);
            uint8_t *q = nullptr;
            if (cudaHostAlloc((void **)&q, nc, cudaHostAllocDefault) != cudaSuccess) return false;
            if (n) memcpy(q, p, n);
            if (p) cudaFreeHost(p);
            p = q; cap = nc;
        }
        n = want;
        return true;
    }
    void erase_front(size_t k) { memmove(p, p + k, n - k); n -= k; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = cap = 0; }
};
#define QRESIZE(q, want)                                                                  \
    do {                                                                                  \
        if (!(q).resize(want)) { snprintf(e->err, sizeof e->err, "cudaHostAlloc of the output queue failed"); return M2V_ENOMEM; } \
    } while (0)

struct m2v_encoder {
    int XL, YL, VL, Q;
    int dev;
    cudaStream_t st = nullptr;
    char err[256] = {0};
    // sequence state (RTL:1017-1022)
    bool busy = false, ended = false;
    int mbw = 0, mbh = 0, P = 0;
    long frames_encoded = 0;           // absolute index of the next frame to encode
    std::vector<uint8_t> stage;        // pushed but not yet encoded frames, planar yuv444p
    long staged_frames = 0;
    size_t px_in_frame = 0;            // pixels of the partially pushed frame (push4)
    long batch_frames = 0;             // flush threshold (whole GOPs)
    long force_batch = 0, force_chunk = 0;   // m2v_set_limits (0 = automatic)
    PinnedQ outq; size_t out_rd = 0;
    // device buffers
    DevBuf<uint8_t> d_in, d_in2, d_recon0, d_recon1, d_body;
    cudaStream_t st_copy = nullptr; cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    DevBuf<int16_t> d_coefs;
    DevBuf<uint32_t> d_mbinfo, d_mb_bits, d_mb_off, d_slice_off, d_frame_bytes, d_out;
    DevBuf<unsigned long long> d_frame_off;
    DevBuf<unsigned> d_k1ctr; unsigned k1_seq = 0;      // K1 work counters (see M2VBatch::k1_ctr)
    // last encode_gops chunk (debug taps) and statistics
    long last_F = 0; int last_nmb = 0;
    long launches = 0;
    bool timing = false; float kms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

static int clamp16(int s, int L) { return s > (1 << L) ? (1 << L) : s < 4 ? 4 : s; }   // RTL:985-991

extern "C" int m2v_create(int XL, int YL, int VL, int Q, m2v_encoder **out) {
    if (!out) return M2V_EINVAL;
    *out = nullptr;
    if (XL < 4 || XL > 7 || YL < 4 || YL > 7 || VL < 1 || VL > 3 || Q < 1 || Q > 4) return M2V_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return M2V_ENODEV;
    m2v_encoder *e = new (std::nothrow) m2v_encoder();
    if (!e) return M2V_ENOMEM;
    e->XL = XL; e->YL = YL; e->VL = VL; e->Q = Q;
    cudaDeviceProp prop;
    if (cudaGetDevice(&e->dev) != cudaSuccess || cudaGetDeviceProperties(&prop, e->dev) != cudaSuccess || prop.major != 10) {
        delete e; return M2V_ENODEV;                          // sm_100a image only; no fallback
    }
    if (cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking) != cudaSuccess || m2v_upload_tables(Q) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {           // table uploads ride the legacy stream: finish them before any launch on e->st
        delete e; return M2V_ECUDA;
    }
    if (e->d_k1ctr.reserve(2) != cudaSuccess || cudaMemset(e->d_k1ctr.p, 0, 2 * sizeof(unsigned)) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) { m2v_destroy(e); return M2V_ECUDA; }
    for (int i = 0; i < 5; i++) cudaEventCreate(&e->ev[i]);
    cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking);
    for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&e->ev_copy[i], cudaEventDisableTiming);
    *out = e;
    return M2V_OK;
}

extern "C" void m2v_destroy(m2v_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->dev);
    if (e->st) { cudaStreamSynchronize(e->st); cudaStreamDestroy(e->st); }
    for (int i = 0; i < 5; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->st_copy) cudaStreamDestroy(e->st_copy);
    for (int i = 0; i < 2; i++) if (e->ev_copy[i]) cudaEventDestroy(e->ev_copy[i]);
    e->d_in.release(); e->d_in2.release(); e->d_recon0.release(); e->d_recon1.release(); e->d_body.release(); e->d_coefs.release();
    e->d_mbinfo.release(); e->d_mb_bits.release(); e->d_mb_off.release(); e->d_slice_off.release();
    e->d_frame_bytes.release(); e->d_out.release(); e->d_frame_off.release(); e->d_k1ctr.release(); e->outq.release();
    delete e;
}

extern "C" const char *m2v_last_error(const m2v_encoder *e) { return e ? e->err : "null handle"; }
extern "C" long m2v_launch_count(const m2v_encoder *e) { return e ? e->launches : 0; }
extern "C" int m2v_set_limits(m2v_encoder *e, long batch_frames, long chunk_frames_) {
    if (!e || batch_frames < 0 || chunk_frames_ < 0) return M2V_EINVAL;
    e->force_batch = batch_frames; e->force_chunk = chunk_frames_;
    return M2V_OK;
}
extern "C" int m2v_set_timing(m2v_encoder *e, int en) { if (!e) return M2V_EINVAL; e->timing = en != 0; return M2V_OK; }
extern "C" int m2v_kernel_ms(const m2v_encoder *e, float ms[5]) { if (!e) return M2V_EINVAL; memcpy(ms, e->kms, sizeof e->kms); return M2V_OK; }

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

// ---- the batch pipeline ------------------------------------------------------------------------
// Encodes frames [n0, n0+F) (whole GOPs, n0 on a GOP boundary) from device memory into e->d_out;
// returns the body length.  One chunk; the caller splits long jobs.
static int encode_chunk(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0, size_t *len) {
    M2VBatch b;
    b.g.mbw = mbw; b.g.mbh = mbh; b.g.W = mbw * 16; b.g.H = mbh * 16; b.g.nmb = mbw * mbh; b.g.P = P; b.g.VL = e->VL; b.g.Q = e->Q;
    b.F = F; b.n0 = n0; b.in = d_in;
    const long gop = P + 1, G = (F + gop - 1) / gop;
    b.CWp = ((b.g.W / 2) + 15) & ~15; b.fsz420 = (size_t)b.g.W * b.g.H + (size_t)2 * b.CWp * (b.g.H / 2);
    const size_t fsz420 = b.fsz420, nmbF = (size_t)F * b.g.nmb;
    CK(e->d_recon0.reserve(G * fsz420)); CK(e->d_recon1.reserve(P ? G * fsz420 : 16));
    CK(e->d_coefs.reserve(nmbF * 384)); CK(e->d_mbinfo.reserve(nmbF)); CK(e->d_mb_bits.reserve(nmbF)); CK(e->d_mb_off.reserve(nmbF));
    CK(e->d_slice_off.reserve((size_t)F * mbh)); CK(e->d_frame_bytes.reserve(F)); CK(e->d_frame_off.reserve(F + 1));
    b.recon[0] = e->d_recon0.p; b.recon[1] = P ? e->d_recon1.p : e->d_recon0.p;
    b.coefs = e->d_coefs.p; b.mbinfo = e->d_mbinfo.p; b.mb_bits = e->d_mb_bits.p; b.mb_off = e->d_mb_off.p;
    b.slice_off = e->d_slice_off.p; b.frame_bytes = e->d_frame_bytes.p; b.frame_off = e->d_frame_off.p; b.out_words = nullptr;
    b.k1_ctr = e->d_k1ctr.p;
    if (G * b.g.nmb >= M2V_K1_MAX_MBS) { snprintf(e->err, sizeof e->err, "chunk too large for one K1 launch"); return M2V_EINVAL; }
    e->last_F = F; e->last_nmb = b.g.nmb;
    if (!m2v_make_tmaps(b)) { snprintf(e->err, sizeof e->err, "cuTensorMapEncodeTiled failed"); return M2V_ECUDA; }

    if (e->timing) CK(cudaEventRecord(e->ev[0], e->st));
    for (int t = 0; t <= P && t < F; t++) {                       // frame t of every GOP that has one
        const long ng = (F - t + gop - 1) / gop;
        m2v_launch_k1(b, t, ng, e->k1_seq++, e->st); e->launches++;
    }
    if (e->timing) CK(cudaEventRecord(e->ev[1], e->st));
    m2v_launch_k2(b, false, e->st); e->launches++;
    if (e->timing) CK(cudaEventRecord(e->ev[2], e->st));
    m2v_launch_k3_scan(b, e->st); e->launches += 2;
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, b.frame_off + F, sizeof total, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    const size_t words = (size_t)(total / 4) + 4;
    CK(e->d_out.reserve(words));
    b.out_words = e->d_out.p;
    CK(cudaMemsetAsync(b.out_words, 0, words * 4, e->st));
    m2v_launch_headers(b, e->st); e->launches++;
    if (e->timing) CK(cudaEventRecord(e->ev[3], e->st));
    m2v_launch_k2(b, true, e->st); e->launches++;
    if (e->timing) CK(cudaEventRecord(e->ev[4], e->st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->st));
    if (e->timing) {
        for (int i = 0; i < 4; i++) { float ms = 0; cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]); e->kms[i] += ms; }
        float ms = 0; cudaEventElapsedTime(&ms, e->ev[0], e->ev[4]); e->kms[4] += ms;
    }
    *len = (size_t)total;
    return M2V_OK;
}

static long chunk_frames(const m2v_encoder *e, int mbw, int mbh, int P) {
    // bound the level buffer (768 B per macroblock) to ~6 GiB per chunk, whole GOPs
    const size_t per_frame = (size_t)mbw * mbh * 768;
    long f = (long)((6ull << 30) / per_frame);
    if (e->force_chunk > 0) f = std::min(f, e->force_chunk);
    const long gop = P + 1;
    f = std::max(gop, f / gop * gop);
    return f;
}

extern "C" int m2v_encode_gops_device(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0,
                                      const uint8_t **d_body, size_t *body_len) {
    if (!e) return M2V_EINVAL;
    if (!d_in || F <= 0 || mbw < 4 || mbw > (1 << e->XL) || mbh < 4 || mbh > (1 << e->YL) || P < 0 || P > 255 || n0 < 0 || n0 % (P + 1) != 0 ||
        !d_body || !body_len) { snprintf(e->err, sizeof e->err, "encode_gops: bad argument"); return M2V_EINVAL; }
    CK(cudaSetDevice(e->dev));
    if (e->timing) memset(e->kms, 0, sizeof e->kms);
    const long cf = chunk_frames(e, mbw, mbh, P);
    const size_t fsz = (size_t)mbw * mbh * 256 * 3;
    if (F <= cf) {                                                 // common case: one chunk, no extra copy
        size_t len = 0;
        int rc = encode_chunk(e, mbw, mbh, P, d_in, F, n0, &len);
        if (rc) return rc;
        *d_body = (const uint8_t *)e->d_out.p; *body_len = len;
        return M2V_OK;
    }
    size_t tot = 0;
    for (long f0 = 0; f0 < F; f0 += cf) {
        const long fc = std::min(cf, F - f0);
        size_t len = 0;
        int rc = encode_chunk(e, mbw, mbh, P, d_in + (size_t)f0 * fsz, fc, n0 + f0, &len);
        if (rc) return rc;
        if (tot + len > e->d_body.n) {                             // grow, keeping what is there
            DevBuf<uint8_t> nb;
            CK(nb.reserve(std::max((tot + len) * 2, (size_t)1 << 20)));
            if (tot) { CK(cudaMemcpyAsync(nb.p, e->d_body.p, tot, cudaMemcpyDeviceToDevice, e->st)); CK(cudaStreamSynchronize(e->st)); }
            e->d_body.release(); e->d_body = nb;
        }
        CK(cudaMemcpyAsync(e->d_body.p + tot, e->d_out.p, len, cudaMemcpyDeviceToDevice, e->st));
        CK(cudaStreamSynchronize(e->st));
        tot += len;
    }
    *d_body = e->d_body.p; *body_len = tot;
    return M2V_OK;
}

extern "C" int m2v_encode_gops_host(m2v_encoder *e, int mbw, int mbh, int P, const uint8_t *d_in, long F, long n0,
                                    uint8_t *h_body, size_t cap, size_t *body_len) {
    const uint8_t *d = nullptr; size_t len = 0;
    int rc = m2v_encode_gops_device(e, mbw, mbh, P, d_in, F, n0, &d, &len);
    if (rc) return rc;
    if (body_len) *body_len = len;
    if (!h_body || len > cap) { snprintf(e->err, sizeof e->err, "encode_gops_host: need %zu bytes, have %zu", len, cap); return M2V_ESPACE; }
    CK(cudaMemcpyAsync(h_body, d, len, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    return M2V_OK;
}

extern "C" int m2v_debug_copy(m2v_encoder *e, uint32_t *mbinfo, int16_t *coefs, long count) {
    if (!e || count > e->last_F * e->last_nmb) return M2V_EINVAL;
    CK(cudaSetDevice(e->dev));
    if (mbinfo) CK(cudaMemcpyAsync(mbinfo, e->d_mbinfo.p, (size_t)count * 4, cudaMemcpyDeviceToHost, e->st));
    if (coefs) CK(cudaMemcpyAsync(coefs, e->d_coefs.p, (size_t)count * 768, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    return M2V_OK;
}

// ---- streaming contract ------------------------------------------------------------------------
extern "C" int m2v_begin(m2v_encoder *e, int xs, int ys, int P, int *mbw, int *mbh) {
    if (!e) return M2V_EINVAL;
    if (e->busy) { snprintf(e->err, sizeof e->err, "begin while o_sequence_busy"); return M2V_ESTATE; }
    if (P < 0 || P > 255) return M2V_EINVAL;
    e->mbw = clamp16(xs, e->XL); e->mbh = clamp16(ys, e->YL); e->P = P;
    if (mbw) *mbw = e->mbw; if (mbh) *mbh = e->mbh;
    e->frames_encoded = 0; e->staged_frames = 0; e->px_in_frame = 0; e->ended = false;
    e->stage.clear(); e->outq.clear(); e->out_rd = 0;
    // flush threshold: enough GOPs that one K1 step has >= 16k macroblocks (a smaller batch shortens the
    // un-overlapped first H2D copy and last encode of the pipeline), staging <= 1 GiB
    const long gop = P + 1, nmb = (long)e->mbw * e->mbh;
    long g = (16384 + nmb - 1) / nmb;
    const size_t fsz = (size_t)nmb * 768;
    while ((size_t)g * gop * fsz < ((size_t)64 << 20)) g++;       // and >= 64 MiB of input, so per-batch overheads stay small
    while (g > 1 && (size_t)g * gop * fsz > ((size_t)1 << 30)) g--;
    e->batch_frames = g * gop;
    if (e->force_batch > 0) e->batch_frames = std::max(gop, std::min(e->batch_frames, e->force_batch / gop * gop));
    // the RTL arms on the first i_en (RTL:1060-1065); the header is emitted then
    return M2V_OK;
}

static int start_if_idle(m2v_encoder *e) {
    if (e->busy) return M2V_OK;
    if (e->mbw == 0 || e->ended) { snprintf(e->err, sizeof e->err, "push before begin"); return M2V_ESTATE; }
    e->busy = true;
    QRESIZE(e->outq, 34);
    return m2v_sequence_header(e->mbw, e->mbh, e->outq.data());
}

static int flush_staged(m2v_encoder *e) {
    if (e->staged_frames == 0) return M2V_OK;
    CK(cudaSetDevice(e->dev));
    const size_t fsz = (size_t)e->mbw * e->mbh * 768, bytes = fsz * e->staged_frames;
    CK(e->d_in.reserve(bytes));
    // on the encoder's own stream: a non-blocking stream does not synchronise with the legacy default stream,
    // and a pageable H2D cudaMemcpy may return before its DMA has finished
    CK(cudaMemcpyAsync(e->d_in.p, e->stage.data(), bytes, cudaMemcpyHostToDevice, e->st));
    const uint8_t *d = nullptr; size_t len = 0;
    int rc = m2v_encode_gops_device(e, e->mbw, e->mbh, e->P, e->d_in.p, e->staged_frames, e->frames_encoded, &d, &len);
    if (rc) return rc;
    const size_t at = e->outq.size();
    QRESIZE(e->outq, at + len);
    CK(cudaMemcpyAsync(e->outq.data() + at, d, len, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    e->frames_encoded += e->staged_frames;
    // keep a partially pushed frame at the front of the staging area
    if (e->px_in_frame) memmove(e->stage.data(), e->stage.data() + bytes, fsz);
    e->staged_frames = 0;
    return M2V_OK;
}

extern "C" int m2v_push4(m2v_encoder *e, const uint8_t Y[4], const uint8_t U[4], const uint8_t V[4]) {
    if (!e || !Y || !U || !V) return M2V_EINVAL;
    int rc = start_if_idle(e); if (rc) return rc;
    if (e->ended) { snprintf(e->err, sizeof e->err, "push after stop"); return M2V_ESTATE; }
    const size_t ysz = (size_t)e->mbw * e->mbh * 256, fsz = ysz * 3;
    const size_t base = (size_t)e->staged_frames * fsz;
    if (e->stage.size() < base + fsz) e->stage.resize(base + fsz);
    uint8_t *f = e->stage.data() + base;
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
    int rc = start_if_idle(e); if (rc) return rc;
    if (e->ended) { snprintf(e->err, sizeof e->err, "push after stop"); return M2V_ESTATE; }
    if (e->px_in_frame) { snprintf(e->err, sizeof e->err, "push_frames inside a frame"); return M2V_ESTATE; }
    const size_t fsz = (size_t)e->mbw * e->mbh * 768;
    // whole batches go straight from the caller's buffer to HBM (no host staging copy), double
    // buffered: the H2D copy of batch i+1 (copy stream) overlaps the kernels of batch i.
    const long gopf = e->P + 1;
    if (e->staged_frames == 0 && nframes >= gopf) {
        CK(cudaSetDevice(e->dev));
        // every whole GOP goes directly; only a trailing partial GOP is staged on the host
        const long direct = nframes / gopf * gopf;
        const long bf = std::min(e->batch_frames, direct);
        const size_t bytes = fsz * bf;
        CK(e->d_in.reserve(bytes)); CK(e->d_in2.reserve(bytes));
        uint8_t *dbuf[2] = {e->d_in.p, e->d_in2.p};
        const long nb = (direct + bf - 1) / bf;
        auto cnt = [&](long i) { return std::min(bf, direct - i * bf); };
        CK(cudaMemcpyAsync(dbuf[0], yuv, fsz * cnt(0), cudaMemcpyHostToDevice, e->st_copy));
        CK(cudaEventRecord(e->ev_copy[0], e->st_copy));
        for (long i = 0; i < nb; i++) {
            if (i + 1 < nb) {
                CK(cudaMemcpyAsync(dbuf[(i + 1) & 1], yuv + (size_t)(i + 1) * bytes, fsz * cnt(i + 1), cudaMemcpyHostToDevice, e->st_copy));
                CK(cudaEventRecord(e->ev_copy[(i + 1) & 1], e->st_copy));
            }
            CK(cudaStreamWaitEvent(e->st, e->ev_copy[i & 1], 0));
            const uint8_t *d = nullptr; size_t len = 0;
            rc = m2v_encode_gops_device(e, e->mbw, e->mbh, e->P, dbuf[i & 1], cnt(i), e->frames_encoded, &d, &len);
            if (rc) return rc;
            const size_t at = e->outq.size();
            QRESIZE(e->outq, at + len);
            CK(cudaMemcpyAsync(e->outq.data() + at, d, len, cudaMemcpyDeviceToHost, e->st));
            CK(cudaStreamSynchronize(e->st));
            e->frames_encoded += cnt(i);
        }
        yuv += (size_t)direct * fsz; nframes -= direct;
    }
    while (nframes > 0) {
        const long take = std::min(nframes, e->batch_frames - e->staged_frames);
        const size_t base = (size_t)e->staged_frames * fsz;
        if (e->stage.size() < base + take * fsz) e->stage.resize(base + take * fsz);
        memcpy(e->stage.data() + base, yuv, take * fsz);
        e->staged_frames += take; yuv += take * fsz; nframes -= take;
        if (e->staged_frames >= e->batch_frames) { rc = flush_staged(e); if (rc) return rc; }
    }
    return M2V_OK;
}

extern "C" int m2v_stop(m2v_encoder *e) {
    if (!e) return M2V_EINVAL;
    if (!e->busy || e->ended) return M2V_OK;                       // stop while idle is ignored (RTL:1090)
    if (e->px_in_frame) {                                          // pad the unfinished frame (RTL:1036-1037, 1049-1056)
        const size_t ysz = (size_t)e->mbw * e->mbh * 256, fsz = ysz * 3;
        uint8_t *f = e->stage.data() + (size_t)e->staged_frames * fsz;
        memset(f + e->px_in_frame, 0, ysz - e->px_in_frame);
        memset(f + ysz + e->px_in_frame, 0x80, ysz - e->px_in_frame);
        memset(f + 2 * ysz + e->px_in_frame, 0x80, ysz - e->px_in_frame);
        e->px_in_frame = 0; e->staged_frames++;
    }
    int rc = flush_staged(e); if (rc) return rc;
    const size_t len = e->outq.size();
    QRESIZE(e->outq, 32 * ((len + 4) / 32 + 1));
    size_t tot = 0;
    rc = m2v_finish_stream(e->outq.data(), len, e->outq.size(), &tot); if (rc) return rc;
    e->ended = true;
    return M2V_OK;
}

extern "C" int m2v_busy(const m2v_encoder *e) { return e && e->busy; }

// queue -> caller's buffer.  A long stream (tens of MB after a big push) is copied by a few threads: one core moves
// ~10 GB/s, and this copy sits on the critical path of the end-to-end time after the last kernel.
static void copy_out(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t kMin = (size_t)2 << 20;
    if (n < 2 * kMin) { memcpy(dst, src, n); return; }
    const int parts = (int)std::min<size_t>(4, n / kMin);
    const size_t per = (n / parts + 63) & ~(size_t)63;
    std::thread th[3];
    for (int i = 1; i < parts; i++) {
        const size_t o = (size_t)i * per, len = std::min(per, n - o);
        th[i - 1] = std::thread([=] { memcpy(dst + o, src + o, len); });
    }
    memcpy(dst, src, std::min(per, n));
    for (int i = 1; i < parts; i++) th[i - 1].join();
}

extern "C" int m2v_drain(m2v_encoder *e, uint8_t *dst, size_t cap, size_t *n, int *last) {
    if (!e || !dst || !n) return M2V_EINVAL;
    size_t avail = (e->outq.size() - e->out_rd) / 32 * 32;
    size_t take = std::min(avail, cap / 32 * 32);
    if (take) copy_out(dst, e->outq.data() + e->out_rd, take);
    e->out_rd += take; *n = take;
    const bool fin = e->ended && e->out_rd == e->outq.size();
    if (last) *last = fin && take > 0;
    if (fin) { e->busy = false; e->ended = false; e->mbw = 0; e->outq.clear(); e->out_rd = 0; }   // back to IDLE (RTL:1045-1047)
    else if (e->out_rd > (64u << 20)) { e->outq.erase_front(e->out_rd); e->out_rd = 0; }
    return M2V_OK;
}

extern "C" int m2v_pull(m2v_encoder *e, uint8_t out[32], int *last) {
    size_t n = 0; int l = 0;
    int rc = m2v_drain(e, out, 32, &n, &l);
    if (rc) return rc;
    if (last) *last = l;
    return n == 32 ? 1 : 0;
}
