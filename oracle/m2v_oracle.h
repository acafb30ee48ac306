/* oracle/m2v_oracle.h
 *
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 * CPU restatement (plain C) of the behaviour of the reference encoder core
 * /root/reference/RTL/mpeg2encoder.v ("RTL" below), used ONLY by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
 * The product (fpga-mpeg2-encoder_b200/) never includes, links or calls anything in oracle/.
 *
 * PARITY STATUS: PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no golden bitstreams and no
 * Verilog simulator exists in the build image or on the GPU box, so oracle/vl2c.py translates the
 * reference RTL (read where it lies under /root/reference, never copied) into a cycle-based C++ model
 * under oracle/_ref/ which oracle/rtl_tb.cpp drives exactly like SIM/tb_mpeg2encoder.v.  That model
 * (a) reproduces the only number the reference publishes - 775 456 bytes for SIM/data.zip:1440x704.yuv
 * with the testbench defaults (README.md:748) - and (b) is byte-identical to this oracle on all three
 * testbench clips run back to back, on every VECTOR_LEVEL x Q_LEVEL, on the S1-S4 clip classes, with
 * mid-frame stops, input bubbles, size clamps and GOP lengths up to 255 (tests/test_rtl_pin.py; committed
 * RTL-written fixtures in tests/golden/).  Caveat, stated once: the simulator that executes the RTL is this
 * repository's own translator, not iverilog.  See DESIGN.md "Oracle".
 */
#ifndef M2V_ORACLE_H
#define M2V_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Optional per-macroblock dumps for localising mismatches (any pointer may be NULL).
 * Index f = frame number in the encoded sequence, mb = by*mbw+bx. */
typedef struct m2v_oracle_dbg {
    int8_t  *mb_inter;   /* [f][mb]          1 = inter (RTL:1805-1814,1820-1825) */
    int8_t  *mb_mvx;     /* [f][mb]          half-pel units (RTL:1827-1828) */
    int8_t  *mb_mvy;     /* [f][mb] */
    uint8_t *mb_cbp;     /* [f][mb]          nzflags, Y00 in bit 5 (RTL:2461-2467) */
    int16_t *coefs;      /* [f][mb][6][64]   quantised levels in zig-zag order */
    uint8_t *recon;      /* [f][W*H*3/2]     reconstructed 4:2:0 planes Y,U,V (RTL:2350-2356) */
} m2v_oracle_dbg;

/* Clamp of i_xsize16 / i_ysize16 (RTL:985-991).  Returns the value in macroblocks. */
int m2v_oracle_clamp16(int size16, int L);

/* Encode one whole sequence exactly as the RTL would emit it (all o_en words concatenated).
 *   frames      : planar yuv444p, frame after frame (Y plane, U plane, V plane; TB:210-218),
 *                 geometry = the CLAMPED size.
 *   nframes     : number of complete frames pushed before i_sequence_stop.
 *   partial_px4 : if >0, one more frame follows of which only the first partial_px4 groups of
 *                 4 pixels (raster order) were pushed before the stop; the rest is padded with
 *                 Y=0,U=V=0x80 (RTL:1036-1037,1049-1056).  frames must then hold nframes+1 frames.
 * Returns 0, or -1 on bad arguments / output overflow.  *outlen is always a multiple of 32. */
int m2v_oracle_encode(int XL, int YL, int VECTOR_LEVEL, int Q_LEVEL,
                      int xsize16, int ysize16, int pframes_count,
                      const uint8_t *frames, long nframes, long partial_px4,
                      uint8_t *out, size_t cap, size_t *outlen, m2v_oracle_dbg *dbg);

/* Body bytes (GOP/picture/slice layers only, byte aligned) of frames [n0, n1) where n0 is the
 * absolute index of an I-frame (n0 % (pframes_count+1) == 0).  frames points at frame n0.
 * Used to run closed GOPs on several host threads for the CPU baseline and to check GOP sharding. */
int m2v_oracle_encode_range(int VECTOR_LEVEL, int Q_LEVEL, int mbw, int mbh, int pframes_count,
                            const uint8_t *frames, long n0, long n1,
                            uint8_t *out, size_t cap, size_t *outlen, m2v_oracle_dbg *dbg);

/* Sequence header block (34 bytes, RTL:2596-2617) and tail (end code + zero padding, RTL:2621-2628,
 * 2932-2937) helpers: seq_header writes exactly 34 bytes; tail_len returns the final file length
 * for a stream of body_end bytes (header + bodies) once `00 00 01 B7` and the padding are added. */
void   m2v_oracle_seq_header(int mbw, int mbh, uint8_t out34[34]);
size_t m2v_oracle_tail_len(size_t body_end);

/* Pure functions exported for known-answer tests. */
int  m2v_oracle_mean2(int a, int b);                                  /* RTL:750-757 */
int  m2v_oracle_mean4(int a, int b, int c, int d);                    /* RTL:760-767 */
int  m2v_oracle_find_min10(const int v[10]);                          /* RTL:804-840 */
void m2v_oracle_fdct_quant(const int16_t res[64], int inter, int Q_LEVEL, int16_t q[64]);  /* RTL:2029-2077 */
void m2v_oracle_dequant_idct(const int16_t q[64], int inter, int Q_LEVEL, int16_t out[64]); /* RTL:2128-2150,844-972 */
int  m2v_oracle_put_ac(int v, int run, uint32_t *code);               /* RTL:2525-2547, returns len */
void m2v_oracle_subsample420(const uint8_t *c444, int W, int H, uint8_t *c420); /* RTL:1086-1089,1167-1170 */

#ifdef __cplusplus
}
#endif
#endif
