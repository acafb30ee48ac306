// oracle/rtl_tb.cpp - TEST INFRASTRUCTURE.
// C++ restatement of the stimulus / sink of the reference testbench (SIM/tb_mpeg2encoder.v:142-274)
// driving the C++ model that oracle/vl2c.py generates from the reference RTL (oracle/_ref/rtl_model_*.hpp,
// never committed).  Mirrors the testbench step by step:
//   reset 4 clocks low once (TB:144-148); per video: i_xsize16 / i_ysize16 / i_pframes_count held
//   (TB:203-204, 106); one clock per 4 pixels with i_en=1, optional idle clocks between them (the bubbles of
//   TB:233); i_sequence_stop for one clock after the last pixel (TB:249-251); every o_en word appended,
//   byte i = o_data[8i+:8] (TB:259-264), until o_sequence_busy falls.  Several videos can run back to back
//   on one instance (TB:150).
// Built by oracle/Makefile into oracle/_ref/librtl_ref_XL<>_YL<>_VL<>_Q<>.so (MODEL_HPP selects the model).
#include MODEL_HPP
#include <stddef.h>

struct RtlRef { Sim sim; long clocks = 0; };

extern "C" void *rtl_ref_create() {
    RtlRef *r = new RtlRef();
    Sim &s = r->sim;
    s.init();
    s.v_rstn = 1; s.v_i_en = 0; s.v_i_sequence_stop = 0;
    for (int i = 0; i < 4; i++) { s.clock(); r->clocks++; }
    s.v_rstn = 0;
    for (int i = 0; i < 4; i++) { s.clock(); r->clocks++; }
    s.v_rstn = 1;
    s.clock(); r->clocks++;
    return r;
}

extern "C" void rtl_ref_destroy(void *h) { delete (RtlRef *)h; }
extern "C" long rtl_ref_clocks(void *h) { return ((RtlRef *)h)->clocks; }

// one video sequence.  bubble_seed != 0 inserts pseudo-random idle clocks (i_en=0) between pixel groups.
extern "C" int rtl_ref_sequence(void *h, const uint8_t *yuv, int xsize16, int ysize16, int pframes, long nframes, long partial_px4,
                                unsigned bubble_seed, uint8_t *out, size_t cap, size_t *outlen) {
    RtlRef *r = (RtlRef *)h;
    Sim &s = r->sim;
    size_t n = 0;
    bool ovf = false, saw_last = false;
    auto tick = [&]() {
        s.clock(); r->clocks++;
        if (s.v_o_en) {
            for (int i = 0; i < 32; i++) {
                if (n < cap) out[n] = (uint8_t)s.v_o_data.slice(8 * i, 8); else ovf = true;
                n++;
            }
            if (s.v_o_last) saw_last = true;
        }
    };
    if (s.v_o_sequence_busy) return -2;
    s.v_i_xsize16 = (u64)xsize16; s.v_i_ysize16 = (u64)ysize16; s.v_i_pframes_count = (u64)pframes;
    s.comb();
    const int W = ((int)s.v_i_max_x16 + 1) * 16, H = ((int)s.v_i_max_y16 + 1) * 16;   // clamped geometry (RTL:985-991)
    const size_t ysz = (size_t)W * H;
    const long total = nframes + (partial_px4 > 0 ? 1 : 0);
    unsigned rng = bubble_seed;
    for (long f = 0; f < total; f++) {
        const uint8_t *Y = yuv + (size_t)f * 3 * ysz, *U = Y + ysz, *V = U + ysz;
        const size_t groups = (f < nframes) ? ysz / 4 : (size_t)partial_px4;
        for (size_t g = 0; g < groups; g++) {
            const size_t p = g * 4;
            s.v_i_en = 1;
            s.v_i_Y0 = Y[p]; s.v_i_Y1 = Y[p + 1]; s.v_i_Y2 = Y[p + 2]; s.v_i_Y3 = Y[p + 3];
            s.v_i_U0 = U[p]; s.v_i_U1 = U[p + 1]; s.v_i_U2 = U[p + 2]; s.v_i_U3 = U[p + 3];
            s.v_i_V0 = V[p]; s.v_i_V1 = V[p + 1]; s.v_i_V2 = V[p + 2]; s.v_i_V3 = V[p + 3];
            tick();
            if (bubble_seed) {
                s.v_i_en = 0;
                for (;;) { rng = rng * 1664525u + 1013904223u; if ((rng >> 16) % 3) break; tick(); }
            }
        }
    }
    s.v_i_en = 0;
    s.v_i_sequence_stop = 1; tick();
    s.v_i_sequence_stop = 0; tick();
    const long limit = r->clocks + (long)(ysz / 4) * 4 + 100000;
    while (s.v_o_sequence_busy && r->clocks < limit) tick();
    if (outlen) *outlen = n;
    return (ovf || s.v_o_sequence_busy || !saw_last) ? -1 : 0;
}
