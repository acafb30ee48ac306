// m2venc_tb - C++ host that replays the stimulus of the reference testbench
// (SIM/tb_mpeg2encoder.v:142-274) through the C-ABI of include/m2venc.h:
// for every video: read planar yuv444p frames (TB:210-218), push them 4 pixels per "clock" or in
// bulk (TB:224-235), pulse i_sequence_stop (TB:249-251), and write every o_en word to the .m2v file
// while o_sequence_busy (TB:256-265).  Videos run back to back on ONE encoder instance (TB:150).
//
// File -> file at line rate: the input is read into a ring of PINNED buffers (m2v_alloc_host) by a read-ahead
// thread that cuts every chunk into slices read by several threads (one pread stream moves a few GB/s from the page
// cache; the PCIe link behind it takes ~50 GB/s), whole chunks are pushed straight from those buffers (no staging
// copy), and the output words go to a write-behind thread.
//
// usage: m2venc_tb [-XL n] [-YL n] [-VL n] [-Q n] [-P n] [-gpus n] [-push4] [-frame] [-chunk frames] [-readers n] [-pin] [-dry]
//                  <in.yuv> <width> <height> <out.m2v> [...more quadruples]
// defaults = testbench defaults: XL=7 YL=6 VECTOR_LEVEL=3 Q_LEVEL=2 i_pframes_count=23 (TB:23-24,98-99,106)
//   -gpus n   one module instance spread over n GPUs (m2v_create_multi)
//   -push4    the 4-pixel port, one call per "clock";  -frame   one frame per push (the testbench's frame loop)
//   -chunk n  frames per read-ahead chunk (default: whole GOPs, ~256 MiB);  -readers n  threads per chunk (default: 2/3 of the hardware threads, 4..16)
//   -pin      no copy at all: the file is mapped and each chunk of the mapping is pinned in place (m2v_register_host) by the
//             read-ahead thread, pushed from there and unpinned (falls back to the copying ring where pinning is refused)
//   -dry      read ahead only, nothing is pushed: prints what the input side alone delivers
#include "../../include/m2venc.h"
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Chunk { uint8_t *p = nullptr; long frames = 0; };

// read-ahead: fills pinned chunks from the file, `readers` pread threads per chunk; or (pin mode) pins chunks of the mapped file
struct ReadAhead {
    int fd = -1; size_t fsz = 0; long total_frames = 0, chunk_frames = 0; int readers = 8;
    bool pin = false; uint8_t *map = nullptr; size_t map_len = 0;
    static const int NBUF = 3;
    uint8_t *buf[NBUF] = {nullptr, nullptr, nullptr}; size_t buf_bytes = 0;      // the copying ring lives as long as the process
    std::mutex mu; std::condition_variable cv;
    std::deque<Chunk> full; std::deque<uint8_t *> empty; bool done = false; int inflight = 0;
    std::thread th;
    bool alloc() {
        if (pin) return true;
        const size_t need = (size_t)chunk_frames * fsz;
        if (need > buf_bytes) {
            for (int i = 0; i < NBUF; i++) { if (buf[i]) m2v_free_host(buf[i]); buf[i] = (uint8_t *)m2v_alloc_host(need); if (!buf[i]) return false; }
            buf_bytes = need;
        }
        empty.clear(); full.clear(); done = false;
        for (int i = 0; i < NBUF; i++) empty.push_back(buf[i]);
        return true;
    }
    void start() { full.clear(); done = false; inflight = 0; th = std::thread([this] { pin ? run_pin() : run(); }); }
    void run() {
        for (long f0 = 0; f0 < total_frames; f0 += chunk_frames) {
            uint8_t *b;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !empty.empty(); }); b = empty.front(); empty.pop_front(); }
            const long nf = std::min(chunk_frames, total_frames - f0);
            const size_t bytes = (size_t)nf * fsz, off0 = (size_t)f0 * fsz;
            const int nt = (int)std::max<size_t>(1, std::min<size_t>(readers, bytes >> 22));
            const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
            std::vector<std::thread> ts;
            for (int t = 0; t < nt; t++) {
                const size_t o = (size_t)t * per;
                if (o >= bytes) break;
                const size_t len = std::min(per, bytes - o);
                ts.emplace_back([=] {
                    size_t got = 0;
                    while (got < len) {
                        ssize_t k = pread(fd, b + o + got, len - got, (off_t)(off0 + o + got));
                        if (k <= 0) break;
                        got += (size_t)k;
                    }
                });
            }
            for (auto &t : ts) t.join();
            { std::lock_guard<std::mutex> lk(mu); Chunk c; c.p = b; c.frames = nf; full.push_back(c); }
            cv.notify_all();
        }
        { std::lock_guard<std::mutex> lk(mu); done = true; }
        cv.notify_all();
    }
    // pin mode: at most two chunks pinned ahead of the consumer; chunk boundaries are page aligned (chunk_frames is a multiple of 16
    // and a frame is a multiple of 256 bytes), the tail of the file is rounded up to the page the mapping ends in
    void run_pin() {
        for (long f0 = 0; f0 < total_frames; f0 += chunk_frames) {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return inflight < 2; }); inflight++; }
            const long nf = std::min(chunk_frames, total_frames - f0);
            uint8_t *ptr = map + (size_t)f0 * fsz;
            size_t bytes = ((size_t)nf * fsz + 4095) & ~(size_t)4095;
            if ((size_t)f0 * fsz + bytes > map_len) bytes = map_len - (size_t)f0 * fsz;
            Chunk c; c.frames = nf;
            c.p = m2v_register_host(ptr, bytes) == M2V_OK ? ptr : nullptr;
            { std::lock_guard<std::mutex> lk(mu); full.push_back(c); }
            cv.notify_all();
            if (!c.p) break;
        }
        { std::lock_guard<std::mutex> lk(mu); done = true; }
        cv.notify_all();
    }
    bool next(Chunk *c) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !full.empty() || done; });
        if (full.empty()) return false;
        *c = full.front(); full.pop_front();
        return true;
    }
    void give_back(uint8_t *b) {
        if (pin) m2v_unregister_host(b);
        { std::lock_guard<std::mutex> lk(mu); if (pin) inflight--; else empty.push_back(b); }
        cv.notify_all();
    }
    void join() { if (th.joinable()) th.join(); }
    void release() { for (int i = 0; i < NBUF; i++) if (buf[i]) { m2v_free_host(buf[i]); buf[i] = nullptr; } buf_bytes = 0; }
};

// write-behind: the drained words are written by their own thread
struct WriteBehind {
    FILE *fo = nullptr;
    std::mutex mu; std::condition_variable cv;
    std::deque<std::vector<uint8_t>> q; bool done = false;
    std::thread th;
    void start() { th = std::thread([this] {
        for (;;) {
            std::vector<uint8_t> v;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !q.empty() || done; }); if (q.empty()) return; v.swap(q.front()); q.pop_front(); }
            fwrite(v.data(), 1, v.size(), fo);
        } }); }
    void put(std::vector<uint8_t> &&v) { { std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(v)); } cv.notify_all(); }
    void stop() { { std::lock_guard<std::mutex> lk(mu); done = true; } cv.notify_all(); if (th.joinable()) th.join(); }
};
}  // namespace

int main(int argc, char **argv) {
    int XL = 7, YL = 6, VL = 3, Q = 2, P = 23, push4 = 0, per_frame = 0, gpus = 1, pin = 0, dry = 0, i = 1;
    // reading a file is a CPU copy out of the page cache: ~3 GB/s per thread; two thirds of the hardware threads, 4..16
    int readers = (int)std::min(16u, std::max(4u, std::thread::hardware_concurrency() * 2 / 3));
    long chunk_opt = 0;
    for (; i < argc && argv[i][0] == '-'; i++) {
        std::string a = argv[i];
        if (a == "-push4") { push4 = 1; continue; }
        if (a == "-frame") { per_frame = 1; continue; }
        if (a == "-pin") { pin = 1; continue; }
        if (a == "-dry") { dry = 1; continue; }
        if (i + 1 >= argc) break;
        int v = atoi(argv[++i]);
        if (a == "-XL") XL = v; else if (a == "-YL") YL = v; else if (a == "-VL") VL = v; else if (a == "-Q") Q = v; else if (a == "-P") P = v;
        else if (a == "-gpus") gpus = v; else if (a == "-chunk") chunk_opt = v; else if (a == "-readers") readers = v;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if ((argc - i) < 4 || (argc - i) % 4) { fprintf(stderr, "usage: %s [opts] in.yuv W H out.m2v [...]\n", argv[0]); return 2; }
    m2v_encoder *e = nullptr;
    int rc = gpus > 1 ? m2v_create_multi(gpus, XL, YL, VL, Q, &e) : m2v_create(XL, YL, VL, Q, &e);   // reset (TB:144-148)
    if (rc) { fprintf(stderr, "*** m2v_create failed (%d): %d B200(s) required\n", rc, gpus); return 1; }
    ReadAhead ra;
    for (int nv = 1; i < argc; i += 4, nv++) {
        const char *fin = argv[i], *fout = argv[i + 3];
        const int xs = atoi(argv[i + 1]), ys = atoi(argv[i + 2]);
        printf("start to encode video %d (%4dx%4d)\n", nv, xs, ys);                // TB:173
        int fd = pin ? open(fin, O_RDWR) : -1;                                       // a shared writable mapping can be pinned in place
        bool use_pin = pin && fd >= 0 && !push4 && !per_frame;
        if (fd < 0) fd = open(fin, O_RDONLY);
        FILE *fo = fopen(fout, "wb");
        if (fd < 0) { printf("*** couldn't open input file\n"); return 1; }        // TB:175-180
        if (!fo) { printf("*** couldn't open output file\n"); return 1; }
        if (xs < 64 || xs > (16 << XL) || xs % 16) { printf("*** xsize=%4d is invalid\n", xs); return 1; }   // TB:189-194
        if (ys < 64 || ys > (16 << YL) || ys % 16) { printf("*** ysize=%4d is invalid\n", ys); return 1; }   // TB:196-201
        int mbw, mbh;
        if ((rc = m2v_begin(e, xs / 16, ys / 16, P, &mbw, &mbh))) { fprintf(stderr, "begin: %s\n", m2v_last_error(e)); return 1; }
        const size_t fsz = (size_t)xs * ys * 3;
        struct stat sb;
        fstat(fd, &sb);
        ra.fd = fd; ra.fsz = fsz; ra.total_frames = (long)((size_t)sb.st_size / fsz);   // whole frames only (TB:220, $feof)
        ra.readers = readers;
        const long gop = P + 1;
        long cf = chunk_opt > 0 ? chunk_opt : std::max(gop, (long)(((size_t)256 << 20) / fsz) / gop * gop);
        if (push4 || per_frame) cf = 1;
        if (use_pin) {
            cf = std::max(16l, cf / 16 * 16);
            ra.map_len = ((size_t)sb.st_size + 4095) & ~(size_t)4095;
            ra.map = (uint8_t *)mmap(nullptr, ra.map_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (ra.map == MAP_FAILED || m2v_register_host(ra.map, 4096) != M2V_OK) {   // pinning refused (file system, limits): copy instead
                if (ra.map != MAP_FAILED) munmap(ra.map, ra.map_len);
                ra.map = nullptr; use_pin = false;
            } else m2v_unregister_host(ra.map);
        }
        ra.pin = use_pin;
        ra.chunk_frames = std::max(1l, use_pin ? cf : std::min(cf, std::max(1l, ra.total_frames)));
        if (!ra.alloc()) { fprintf(stderr, "*** pinned host memory for the read-ahead ring could not be allocated\n"); return 1; }
        const auto t0 = std::chrono::steady_clock::now();          // the clock starts before the first byte is read
        ra.start();
        WriteBehind wb; wb.fo = fo; wb.start();
        auto sink = [&]() {                                                       // TB:259-264
            for (;;) {
                std::vector<uint8_t> word(1 << 22);
                size_t n = 0; int last = 0;
                if (m2v_drain(e, word.data(), word.size(), &n, &last)) return;
                const bool more = n == word.size() / 32 * 32;
                if (n) { word.resize(n); wb.put(std::move(word)); }
                if (!more) return;
            }
        };
        long f = 0;
        Chunk c;
        while (ra.next(&c)) {
            if (!c.p) { fprintf(stderr, "*** pinning a chunk of the mapped file failed\n"); return 1; }
            if (dry) { ra.give_back(c.p); f += c.frames; continue; }
            if (push4) {
                const uint8_t *Y = c.p, *U = Y + (size_t)xs * ys, *V = U + (size_t)xs * ys;
                for (size_t p = 0; p < (size_t)xs * ys; p += 4) rc |= m2v_push4(e, Y + p, U + p, V + p);
            } else rc = m2v_push_frames(e, c.p, c.frames);
            if (rc) { fprintf(stderr, "push: %s\n", m2v_last_error(e)); return 1; }
            ra.give_back(c.p);
            sink();
            f += c.frames;
        }
        if ((rc = m2v_stop(e))) { fprintf(stderr, "stop: %s\n", m2v_last_error(e)); return 1; }   // TB:249-251
        sink();
        wb.stop(); ra.join();
        fclose(fo);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();   // last word written
        if (ra.map) { munmap(ra.map, ra.map_len); ra.map = nullptr; }
        close(fd);
        printf("end of video %d (%ld frames), busy=%d, %.3f s, %.1f Mpixel/s file to file\n", nv, f, m2v_busy(e), dt,
               dt > 0 ? (double)f * xs * ys / dt / 1e6 : 0.0);                       // TB:270
    }
    ra.release();
    m2v_destroy(e);
    return 0;
}
