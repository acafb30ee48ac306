/* oracle/m2v_oracle.c
 *
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE (see m2v_oracle.h for the usage rule and parity status).
 *
 * Frame-level functional restatement of /root/reference/RTL/mpeg2encoder.v ("RTL").  The RTL is a
 * 64-clock-per-macroblock pipeline; its output is a pure function of the parameters and of the
 * pixel sequence (no back-pressure, every per-macroblock state re-initialised, the P-frame search
 * only ever sees the complete reconstruction of the previous frame), so each stage is restated
 * here as a function of whole macroblocks / frames.  Every block cites the RTL lines it follows.
 */
#include "m2v_oracle.h"
#include "m2v_tables.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* pure helpers                                                                                */
/* ------------------------------------------------------------------------------------------ */
int m2v_oracle_mean2(int a, int b) { return (a + b + 1) >> 1; }                 /* RTL:750-757 */
int m2v_oracle_mean4(int a, int b, int c, int d) { return (a + b + c + d + 1) >> 2; } /* RTL:760-767: +1, not +2 */
#define mean2 m2v_oracle_mean2
#define mean4 m2v_oracle_mean4
static inline int iabs(int x) { return x < 0 ? -x : x; }

int m2v_oracle_clamp16(int s, int L) {                                           /* RTL:985-991 */
    if (s > (1 << L)) return 1 << L;
    if (s < 4) return 4;
    return s;
}

/* 10-way argmin with the RTL's fixed tie order (RTL:804-840).  v[] are 13-bit keys. */
int m2v_oracle_find_min10(const int v[10]) {
    int i01 = v[1] < v[0], m01 = i01 ? v[1] : v[0];
    int i23 = v[3] < v[2], m23 = i23 ? v[3] : v[2];
    int i45 = v[5] < v[4], m45 = i45 ? v[5] : v[4];
    int i67 = v[7] < v[6], m67 = i67 ? v[7] : v[6];
    int i89 = v[9] < v[8], m89 = i89 ? v[9] : v[8];
    int hi03 = m23 < m01, m03 = hi03 ? m23 : m01;
    int hi47 = m67 < m45, m47 = hi47 ? m67 : m45;
    if (m89 <= m03 && m89 <= m47) return 8 + i89;
    if (m03 < m47) return hi03 ? 2 + i23 : 0 + i01;
    return hi47 ? 6 + i67 : 4 + i45;
}

/* 4:4:4 -> 4:2:0: horizontal mean2 of pixel pairs (RTL:1086-1089), then vertical mean2 of the
 * two horizontally subsampled rows (RTL:1167-1170).  Two cascaded roundings. */
void m2v_oracle_subsample420(const uint8_t *c, int W, int H, uint8_t *o) {
    for (int j = 0; j < H / 2; j++)
        for (int i = 0; i < W / 2; i++) {
            int h0 = mean2(c[(2 * j) * W + 2 * i], c[(2 * j) * W + 2 * i + 1]);
            int h1 = mean2(c[(2 * j + 1) * W + 2 * i], c[(2 * j + 1) * W + 2 * i + 1]);
            o[j * (W / 2) + i] = (uint8_t)mean2(h1, h0);
        }
}

/* Forward transform + quantiser of one 8x8 tile (RTL:2029-2077).  res = cur - pred (9-bit). */
void m2v_oracle_fdct_quant(const int16_t res[64], int inter, int Q, int16_t q[64]) {
    int32_t A[64], B;
    for (int r = 0; r < 8; r++)                      /* phase 1: right-multiply DCTM^T (RTL:2029-2036) */
        for (int j = 0; j < 8; j++) {
            int32_t s = 0;
            for (int k = 0; k < 8; k++) s += res[r * 8 + k] * M2V_DCTM[j * 8 + k];
            A[r * 8 + j] = s;
        }
    for (int i = 0; i < 8; i++)                      /* phase 2: left-multiply DCTM (RTL:2054-2062) */
        for (int c = 0; c < 8; c++) {
            B = 0;
            for (int k = 0; k < 8; k++) B += M2V_DCTM[i * 8 + k] * A[k * 8 + c];
            /* (B>>12) + B[11]  ==  floor((B+2048)/4096); low 17 bits, signed (RTL:2058-2059) */
            int32_t C = (B + 2048) >> 12;
            C = (int32_t)((uint32_t)C << 15) >> 15;
            int a = iabs(C) & 0xFFFF, y;             /* g_t3 is 16 bit (RTL:2068) */
            if (inter)
                y = (a + 2) >> (4 + Q);                                              /* RTL:2070 */
            else if (i || c) {
                int w = M2V_INTRA_Q[i * 8 + c];
                y = ((a + ((w * ((3 << Q) + 2)) >> 3)) >> Q) / w;                    /* RTL:2072 */
            } else
                y = (a >> 4) + ((a >> 3) & 1);                                       /* RTL:2074 */
            if (y > 2047) y = 2047;                                                  /* RTL:2075 */
            q[i * 8 + c] = (int16_t)(C < 0 ? -y : y);                                /* RTL:2076 */
        }
}

#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565
static inline int32_t sx18(int32_t v) { return (int32_t)((uint32_t)v << 14) >> 14; }

/* Inverse quantiser (RTL:2128-2150) + Chen-Wang IDCT rows (RTL:844-907) then columns
 * (RTL:911-972).  out = 9-bit residual clipped to +-255.  32-bit wrap arithmetic as in the RTL
 * (done in uint32_t to stay defined). */
void m2v_oracle_dequant_idct(const int16_t q[64], int inter, int Q, int16_t out[64]) {
    int32_t iq[64], r1[64];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            int32_t x = q[i * 8 + j];
            if (inter) {
                x = x * 2;
                x += (x < 0) ? -1 : (x > 0) ? 1 : 0;
                x = x * (1 << Q);
                if (x < -2047) x = -2047; else if (x > 2047) x = 2047;
            } else if (i || j) {
                x = x * M2V_INTRA_Q[i * 8 + j];
                x = (int32_t)((uint32_t)x << 15) >> 15;            /* h_t1 is 17 bit signed */
                if (Q >= 3) x = x * (1 << (Q - 3)); else x = x >> (3 - Q);   /* >>> = floor */
                x = (int32_t)((uint32_t)x << 15) >> 15;
                if (x < -2047) x = -2047; else if (x > 2047) x = 2047;
            } else
                x = x * 2;
            iq[i * 8 + j] = (int32_t)((uint32_t)x << 19) >> 19;    /* h_iquant is 13 bit signed */
        }
    for (int i = 0; i < 8; i++) {                                   /* rows (RTL:849-904) */
        const int32_t *a = iq + i * 8;
        uint32_t x0 = ((uint32_t)a[0] << 11) | 128u, x1 = (uint32_t)a[4] << 11, x2 = a[6], x3 = a[2],
                 x4 = a[1], x5 = a[7], x6 = a[5], x7 = a[3], x8;
        x8 = W7 * (x4 + x5);
        x4 = x8 + (W1 - W7) * x4;
        x5 = x8 - (W1 + W7) * x5;
        x8 = W3 * (x6 + x7);
        x6 = x8 - (W3 - W5) * x6;
        x7 = x8 - (W3 + W5) * x7;
        x8 = x0 + x1;
        x0 = x0 - x1;
        x1 = W6 * (x3 + x2);
        x2 = x1 - (W2 + W6) * x2;
        x3 = x1 + (W2 - W6) * x3;
        x1 = x4 + x6;
        x4 = x4 - x6;
        x6 = x5 + x7;
        x5 = x5 - x7;
        x7 = x8 + x3;
        x8 = x8 - x3;
        x3 = x0 + x2;
        x0 = x0 - x2;
        x2 = (uint32_t)(((int32_t)(181u * (x4 + x5) + 128u)) >> 8);
        x4 = (uint32_t)(((int32_t)(181u * (x4 - x5) + 128u)) >> 8);
        int32_t *r = r1 + i * 8;
        r[0] = sx18((int32_t)(x7 + x1) >> 8);
        r[1] = sx18((int32_t)(x3 + x2) >> 8);
        r[2] = sx18((int32_t)(x0 + x4) >> 8);
        r[3] = sx18((int32_t)(x8 + x6) >> 8);
        r[4] = sx18((int32_t)(x8 - x6) >> 8);
        r[5] = sx18((int32_t)(x0 - x4) >> 8);
        r[6] = sx18((int32_t)(x3 - x2) >> 8);
        r[7] = sx18((int32_t)(x7 - x1) >> 8);
    }
    for (int c = 0; c < 8; c++) {                                   /* columns (RTL:916-970) */
        uint32_t x0 = ((uint32_t)r1[0 * 8 + c] << 8) + 8192u, x1 = (uint32_t)r1[4 * 8 + c] << 8,
                 x2 = r1[6 * 8 + c], x3 = r1[2 * 8 + c], x4 = r1[1 * 8 + c], x5 = r1[7 * 8 + c],
                 x6 = r1[5 * 8 + c], x7 = r1[3 * 8 + c], x8;
        x8 = W7 * (x4 + x5) + 4u;
        x4 = (uint32_t)((int32_t)(x8 + (W1 - W7) * x4) >> 3);
        x5 = (uint32_t)((int32_t)(x8 - (W1 + W7) * x5) >> 3);
        x8 = W3 * (x6 + x7) + 4u;
        x6 = (uint32_t)((int32_t)(x8 - (W3 - W5) * x6) >> 3);
        x7 = (uint32_t)((int32_t)(x8 - (W3 + W5) * x7) >> 3);
        x8 = x0 + x1;
        x0 = x0 - x1;
        x1 = W6 * (x3 + x2) + 4u;
        x2 = (uint32_t)((int32_t)(x1 - (W2 + W6) * x2) >> 3);
        x3 = (uint32_t)((int32_t)(x1 + (W2 - W6) * x3) >> 3);
        x1 = x4 + x6;
        x4 = x4 - x6;
        x6 = x5 + x7;
        x5 = x5 - x7;
        x7 = x8 + x3;
        x8 = x8 - x3;
        x3 = x0 + x2;
        x0 = x0 - x2;
        x2 = (uint32_t)(((int32_t)(181u * (x4 + x5) + 128u)) >> 8);
        x4 = (uint32_t)(((int32_t)(181u * (x4 - x5) + 128u)) >> 8);
        int32_t o[8];
        o[0] = (int32_t)(x7 + x1) >> 14;
        o[1] = (int32_t)(x3 + x2) >> 14;
        o[2] = (int32_t)(x0 + x4) >> 14;
        o[3] = (int32_t)(x8 + x6) >> 14;
        o[4] = (int32_t)(x8 - x6) >> 14;
        o[5] = (int32_t)(x0 - x4) >> 14;
        o[6] = (int32_t)(x3 - x2) >> 14;
        o[7] = (int32_t)(x7 - x1) >> 14;
        for (int i = 0; i < 8; i++) {
            /* clip_neg255_pos255 takes a 28-bit signed argument (RTL:778-783) */
            int32_t v = (int32_t)((uint32_t)o[i] << 4) >> 4;
            out[i * 8 + c] = (int16_t)(v < -255 ? -255 : v > 255 ? 255 : v);
        }
    }
}

/* run/level -> code (RTL:2525-2547).  Returns length, *code holds the bits (sign included). */
int m2v_oracle_put_ac(int v, int run, uint32_t *code) {
    int m = iabs(v) - 1;                      /* absv-1 */
    int s = v < 0;
    int in03 = (run == 0 && m < 40) || (run == 1 && m < 18) || (run == 2 && m < 5) || (run == 3 && m < 4);
    int in431 = (run <= 6 && m < 3) || (run <= 16 && m < 2) || (run <= 31 && m < 1);
    if (in03 || in431) {
        uint32_t e = M2V_VLC_AC[run * M2V_AC_LEVELS + m];
        *code = ((e & 0xFFFF) << 1) | (uint32_t)s;
        return (int)(e >> 16) + 1;
    }
    *code = (1u << 18) | ((uint32_t)(run & 63) << 12) | ((uint32_t)v & 0xFFF);   /* escape */
    return 24;
}

/* ------------------------------------------------------------------------------------------ */
/* bit writer: MSB first; "align" = zero bits up to the next byte boundary (RTL:2940-2943)     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t *buf; size_t cap, pos; uint64_t acc; int nacc; int ovf; } bw_t;

static void bw_put(bw_t *w, uint32_t code, int len) {
    if (len <= 0) return;
    w->acc = (w->acc << len) | (code & ((len >= 32) ? 0xFFFFFFFFu : ((1u << len) - 1)));
    w->nacc += len;
    while (w->nacc >= 8) {
        if (w->pos < w->cap) w->buf[w->pos] = (uint8_t)(w->acc >> (w->nacc - 8)); else w->ovf = 1;
        w->pos++;
        w->nacc -= 8;
    }
}
static void bw_align(bw_t *w) { if (w->nacc) bw_put(w, 0, 8 - w->nacc); }

/* ------------------------------------------------------------------------------------------ */
/* encoder context                                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int VL, Q, mbw, mbh, W, H, P;
    uint8_t *cur[3], *ref[3], *rec[3];        /* 4:2:0 planes */
    bw_t bw;
    m2v_oracle_dbg *dbg;
    long dbg_frame;                            /* frame slot in the dbg arrays */
} enc_t;

static void put_seq_header(bw_t *w, int W, int H) {                  /* RTL:2596-2617 */
    bw_align(w);
    bw_put(w, 0x000001, 24); bw_put(w, 0xB3, 8);
    bw_put(w, (uint32_t)W, 12); bw_put(w, (uint32_t)H, 12);
    bw_put(w, 0x1209c4, 24); bw_put(w, 0x200000, 24); bw_put(w, 0x0001B5, 24); bw_put(w, 0x144200, 24);
    bw_put(w, 0x010000, 24); bw_put(w, 0x000001, 24); bw_put(w, 0xB52305, 24); bw_put(w, 0x0505, 16);
    bw_put(w, (uint32_t)W, 14); bw_put(w, 1, 1); bw_put(w, (uint32_t)H, 14);
}

void m2v_oracle_seq_header(int mbw, int mbh, uint8_t out34[34]) {
    bw_t w; memset(&w, 0, sizeof w); w.buf = out34; w.cap = 34;
    put_seq_header(&w, mbw * 16, mbh * 16);
    bw_align(&w);
}

size_t m2v_oracle_tail_len(size_t body_end) {                        /* RTL:2621-2628, 2932-2937 */
    size_t n = body_end + 4;
    return 32 * (n / 32 + 1);
}

/* SAD of the current luma block against ref at integer offset (dy,dx) */
static int sad16(const uint8_t *c, const uint8_t *r, int stride) {
    int s = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) s += iabs((int)c[y * stride + x] - (int)r[y * stride + x]);
    return s;
}

/* one macroblock: decision, prediction, transform, reconstruction, entropy */
typedef struct { int dcp[3]; int pmvx, pmvy; } slice_state;

static void encode_mb(enc_t *e, int k, int by, int bx, slice_state *ss) {
    const int W = e->W, CW = W / 2, R = 2 * e->VL, Q = e->Q;
    const int Y0 = by * 16, X0 = bx * 16;
    const uint8_t *cY = e->cur[0] + Y0 * W + X0;
    int inter = 0, mvx = 0, mvy = 0;               /* half-pel units */
    uint8_t pred[6][64];
    int fmvy = 0, fmvx = 0, hy = 0, hx = 0;

    if (k != 0) {
        /* ---- full-pel search (RTL:1634-1715) ------------------------------------------------ */
        int have = 0, best = 0;
        for (int dy = -R; dy <= R; dy++)
            for (int dx = -R; dx <= R; dx++) {
                if ((bx == 0 && dx < 0) || (bx == e->mbw - 1 && dx > 0) ||
                    (by == 0 && dy < 0) || (by == e->mbh - 1 && dy > 0)) continue;   /* RTL:1642-1645 */
                int s = sad16(cY, e->ref[0] + (Y0 + dy) * W + X0 + dx, W);
                if (s >= 4096) continue;                     /* 13th bit sticky (RTL:1669-1670) */
                /* survivors = minimal SAD; mvy = largest dy, then largest dx (RTL:1696-1710) */
                if (!have || s <= best) { have = 1; best = s; fmvy = dy; fmvx = dx; }
            }
        if (!have) { fmvy = 0; fmvx = 0; }                   /* RTL:1695,1707 */

        /* ---- half-pel refinement + intra/inter decision (RTL:1743-1816) --------------------- */
        int t[18][18];                                       /* t[y+1][x+1], y,x in -1..16 */
        for (int y = -1; y <= 16; y++)
            for (int x = -1; x <= 16; x++) {
                int yy = Y0 + fmvy + y, xx = X0 + fmvx + x;
                t[y + 1][x + 1] = (yy >= 0 && yy < e->H && xx >= 0 && xx < W) ? e->ref[0][yy * W + xx] : 0;
            }
        int key[10];
        int S = 0;
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) S += cY[y * W + x];
        for (int cy = -1; cy <= 1; cy++)
            for (int cx = -1; cx <= 1; cx++) {
                int dis = ((bx == 0 || fmvx == -R) && cx < 0) || ((bx == e->mbw - 1 || fmvx == R) && cx > 0) ||
                          ((by == 0 || fmvy == -R) && cy < 0) || ((by == e->mbh - 1 || fmvy == R) && cy > 0); /* RTL:1757-1760 */
                int s = 0;
                if (!dis) {
                    for (int y = 0; y < 16; y++)
                        for (int x = 0; x < 16; x++) {
                            /* f_Y_hlf[2y+cy][2x+cx] (RTL:1746-1752) */
                            int gy = 2 * y + cy, gx = 2 * x + cx;
                            int ty = (gy + 2) / 2 - 1, tx = (gx + 2) / 2 - 1;     /* floor(g/2) for g>=-1 */
                            int oy = gy & 1, ox = gx & 1, p;
                            const int (*T)[18] = (const int (*)[18])t;
                            if (!oy && !ox) p = T[ty + 1][tx + 1];
                            else if (!oy) p = mean2(T[ty + 1][tx + 1], T[ty + 1][tx + 2]);
                            else if (!ox) p = mean2(T[ty + 1][tx + 1], T[ty + 2][tx + 1]);
                            else p = mean4(T[ty + 1][tx + 1], T[ty + 1][tx + 2], T[ty + 2][tx + 1], T[ty + 2][tx + 2]);
                            s += iabs((int)cY[y * W + x] - p);
                        }
                }
                key[3 * (cy + 1) + (cx + 1)] = (dis || s >= 4096) ? 8191 : s;    /* {f_over,f_diff} */
            }
        {   /* intra key: f_Y_sum keeps the pixel sum and adds sum|cur-mean| on top (RTL:1600,1662,1744,1776-1777,1791) */
            int m = (S >> 8) & 0xFF, D = 0;
            for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) D += iabs((int)cY[y * W + x] - m);
            int T = (S + D) & 0xFFFF;
            key[9] = (T < 4096) ? T : 4095;
        }
        int w = m2v_oracle_find_min10(key);
        if (w == 9) { inter = 0; hy = hx = 0; } else { inter = 1; hy = w / 3 - 1; hx = w % 3 - 1; }
        mvy = 2 * fmvy + hy;                                  /* RTL:1827-1828 (kept for intra too) */
        mvx = 2 * fmvx + hx;

        /* ---- prediction (RTL:1847-1917) ----------------------------------------------------- */
        if (inter) {
            const int (*T)[18] = (const int (*)[18])t;
            for (int y = 0; y < 16; y++)
                for (int x = 0; x < 16; x++) {
                    int gy = 2 * y + hy, gx = 2 * x + hx;
                    int ty = (gy + 2) / 2 - 1, tx = (gx + 2) / 2 - 1, oy = gy & 1, ox = gx & 1, p;
                    if (!oy && !ox) p = T[ty + 1][tx + 1];
                    else if (!oy) p = mean2(T[ty + 1][tx + 1], T[ty + 1][tx + 2]);
                    else if (!ox) p = mean2(T[ty + 1][tx + 1], T[ty + 2][tx + 1]);
                    else p = mean4(T[ty + 1][tx + 1], T[ty + 1][tx + 2], T[ty + 2][tx + 1], T[ty + 2][tx + 2]);
                    pred[(y >> 3) * 2 + (x >> 3)][(y & 7) * 8 + (x & 7)] = (uint8_t)p;
                }
            /* chroma vector = luma vector >> 1 with FLOOR (RTL:1854,1860,1876,1882,1904-1910) */
            int cyv = mvy >> 1, cxv = mvx >> 1;
            int fy = cyv >> 1, fx = cxv >> 1, oy = cyv & 1, ox = cxv & 1;
            for (int c = 1; c <= 2; c++) {
                const uint8_t *rc = e->ref[c];
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) {
                        int py = by * 8 + y + fy, px = bx * 8 + x + fx;
                        int a = rc[py * CW + px], p;
                        if (oy && ox) p = mean4(a, rc[py * CW + px + 1], rc[(py + 1) * CW + px], rc[(py + 1) * CW + px + 1]);
                        else if (ox) p = mean2(a, rc[py * CW + px + 1]);
                        else if (oy) p = mean2(a, rc[(py + 1) * CW + px]);
                        else p = a;
                        pred[3 + c][y * 8 + x] = (uint8_t)p;
                    }
            }
        }
    }
    if (!inter) memset(pred, 0x80, sizeof pred);              /* RTL:1894-1895,1901-1903; I-frame RTL:1820-1825 */

    /* ---- residual, transform, quantise, scan, reconstruct (RTL:1980-2077, 2128-2356, 2452-2467) */
    int16_t zz[6][64];
    int cbp = 0;
    for (int tl = 0; tl < 6; tl++) {
        int16_t res[64], q[64], rr[64];
        const uint8_t *src; int stride; uint8_t *dst;
        if (tl < 4) {
            int oy = (tl >> 1) * 8, ox = (tl & 1) * 8;
            src = e->cur[0] + (Y0 + oy) * W + X0 + ox; dst = e->rec[0] + (Y0 + oy) * W + X0 + ox; stride = W;
        } else {
            src = e->cur[tl - 3] + (by * 8) * CW + bx * 8; dst = e->rec[tl - 3] + (by * 8) * CW + bx * 8; stride = CW;
        }
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) res[y * 8 + x] = (int16_t)((int)src[y * stride + x] - (int)pred[tl][y * 8 + x]);
        m2v_oracle_fdct_quant(res, inter, Q, q);
        int nz = !inter;
        for (int i = 0; i < 64; i++) { zz[tl][M2V_ZIGZAG[i]] = q[i]; nz |= (q[i] != 0); }      /* RTL:2461-2466 */
        cbp = (cbp << 1) | nz;                                                                 /* RTL:2467 */
        m2v_oracle_dequant_idct(q, inter, Q, rr);
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                int v = (int)pred[tl][y * 8 + x] + rr[y * 8 + x];                               /* RTL:786-795,2352 */
                dst[y * stride + x] = (uint8_t)(v > 255 ? 255 : v < 0 ? 0 : v);
            }
    }

    if (e->dbg) {
        long mb = e->dbg_frame * e->mbw * e->mbh + by * e->mbw + bx;
        if (e->dbg->mb_inter) e->dbg->mb_inter[mb] = (int8_t)inter;
        if (e->dbg->mb_mvx) e->dbg->mb_mvx[mb] = (int8_t)mvx;
        if (e->dbg->mb_mvy) e->dbg->mb_mvy[mb] = (int8_t)mvy;
        if (e->dbg->mb_cbp) e->dbg->mb_cbp[mb] = (uint8_t)cbp;
        if (e->dbg->coefs) memcpy(e->dbg->coefs + mb * 384, zz, sizeof zz);
    }

    /* ---- entropy layer (RTL:2718-2847) ------------------------------------------------------- */
    bw_t *w = &e->bw;
    if (!inter && k != 0) bw_put(w, 0x23, 6);                 /* intra in P: 1 00011 (RTL:2722-2724) */
    else if (inter && cbp == 0) bw_put(w, 0x09, 4);           /* MC not coded: 1 001 (RTL:2725-2727) */
    else bw_put(w, 0x03, 2);                                  /* 1 1 (RTL:2728-2730) */
    if (inter) {
        for (int comp = 0; comp < 2; comp++) {                /* x then y (RTL:2736-2763) */
            int d = comp ? mvy - ss->pmvy : mvx - ss->pmvx;
            if (d > 15) d -= 32; else if (d < -16) d += 32;
            uint32_t en = M2V_VLC_MOTION[iabs(d)];
            bw_put(w, en & 0xFFFF, (int)(en >> 16));
            if (d != 0) bw_put(w, d < 0, 1);
        }
        bw_put(w, M2V_VLC_CBP[cbp] & 0xFFFF, (int)(M2V_VLC_CBP[cbp] >> 16));       /* RTL:2766-2767 */
        ss->pmvx = mvx; ss->pmvy = mvy;                                            /* RTL:2769-2770 */
    } else { ss->pmvx = 0; ss->pmvy = 0; }                                         /* RTL:2772-2773 */

    for (int tl = 0; tl < 6; tl++) {
        int nz = (cbp >> (5 - tl)) & 1;
        int comp = tl < 4 ? 0 : tl - 3;
        int val = zz[tl][0];
        int diff = val - ss->dcp[comp];
        ss->dcp[comp] = inter ? 0 : val;                                           /* RTL:2784-2793 */
        int run = 0;
        if (inter) {                                                               /* RTL:2795-2806 */
            if (val == 0) run = 1;
            else if (val == 1 || val == -1) { if (nz) bw_put(w, 2u | (val < 0), 2); }
            else if (nz) { uint32_t c; int l = m2v_oracle_put_ac(val, 0, &c); bw_put(w, c, l); }
        } else {                                                                   /* RTL:2808-2821 */
            int a = iabs(diff), size = 0;
            while (a >> size) size++;
            uint32_t bits = (uint32_t)(diff < 0 ? diff + (1 << size) - 1 : diff) & 0xFFF;
            uint32_t en = tl < 4 ? M2V_VLC_DC_Y[size] : M2V_VLC_DC_C[size];
            if (nz) { bw_put(w, en & 0xFFFF, (int)(en >> 16)); bw_put(w, bits, size); }
        }
        for (int i = 1; i < 64; i++) {                                             /* RTL:2824-2833 */
            int v = zz[tl][i];
            if (v != 0) {
                if (nz) { uint32_t c; int l = m2v_oracle_put_ac(v, run, &c); bw_put(w, c, l); }
                run = 0;
            } else run++;
        }
        if (nz) bw_put(w, 2, 2);                                                   /* EOB (RTL:2835,2897-2900) */
    }
}

static void encode_frame(enc_t *e, const uint8_t *f444, long n) {
    const int W = e->W, H = e->H, Q = e->Q;
    const int k = (int)(n % (e->P + 1));                        /* a_i_frame (RTL:1078) */
    memcpy(e->cur[0], f444, (size_t)W * H);                     /* Y passes through (RTL:1085) */
    m2v_oracle_subsample420(f444 + (size_t)W * H, W, H, e->cur[1]);
    m2v_oracle_subsample420(f444 + (size_t)2 * W * H, W, H, e->cur[2]);
    bw_t *w = &e->bw;
    if (k == 0) {                                               /* GOP header (RTL:2645-2656), time code RTL:2685-2698 */
        long hh = n / 86400; if (hh > 63) hh = 63;
        bw_align(w);
        bw_put(w, 0x000001, 24); bw_put(w, 0xB8, 8);
        bw_put(w, (uint32_t)hh, 6); bw_put(w, (uint32_t)((n / 1440) % 60), 6);
        bw_put(w, 0x40u | (uint32_t)((n / 24) % 60), 7); bw_put(w, (uint32_t)(n % 24), 6);
        bw_put(w, 2, 2);
    }
    bw_align(w);                                                /* picture header + coding ext (RTL:2666-2682) */
    bw_put(w, 0x000001, 24); bw_put(w, (uint32_t)k, 18);
    if (k == 0) { bw_put(w, 0x10000, 19); bw_put(w, 0, 3); } else { bw_put(w, 0x20000, 19); bw_put(w, 0x380, 11); }
    bw_put(w, 0x000001, 24); bw_put(w, 0xB58111, 24); bw_put(w, 0x1BC000, 24);
    for (int by = 0; by < e->mbh; by++) {
        bw_align(w);                                            /* slice header (RTL:2704-2710) */
        bw_put(w, 0x000001, 24); bw_put(w, (uint32_t)(by + 1), 8); bw_put(w, (uint32_t)(2 << Q), 6);
        slice_state ss; memset(&ss, 0, sizeof ss);              /* RTL:2713-2715 */
        for (int bx = 0; bx < e->mbw; bx++) encode_mb(e, k, by, bx, &ss);
    }
    if (e->dbg && e->dbg->recon) {
        uint8_t *d = e->dbg->recon + (size_t)e->dbg_frame * (W * H * 3 / 2);
        memcpy(d, e->rec[0], (size_t)W * H);
        memcpy(d + W * H, e->rec[1], (size_t)W * H / 4);
        memcpy(d + W * H * 5 / 4, e->rec[2], (size_t)W * H / 4);
    }
    for (int c = 0; c < 3; c++) { uint8_t *tmp = e->ref[c]; e->ref[c] = e->rec[c]; e->rec[c] = tmp; }
    e->dbg_frame++;
}

static int enc_init(enc_t *e, int VL, int Q, int mbw, int mbh, int P, uint8_t *out, size_t cap, m2v_oracle_dbg *dbg) {
    memset(e, 0, sizeof *e);
    if (VL < 1 || VL > 3 || Q < 1 || Q > 4 || mbw < 4 || mbh < 4 || P < 0 || P > 255) return -1;
    e->VL = VL; e->Q = Q; e->mbw = mbw; e->mbh = mbh; e->W = mbw * 16; e->H = mbh * 16; e->P = P;
    size_t ysz = (size_t)e->W * e->H;
    for (int c = 0; c < 3; c++) {
        size_t sz = c ? ysz / 4 : ysz;
        e->cur[c] = (uint8_t *)malloc(sz); e->ref[c] = (uint8_t *)calloc(sz, 1); e->rec[c] = (uint8_t *)calloc(sz, 1);
        if (!e->cur[c] || !e->ref[c] || !e->rec[c]) return -1;
    }
    e->bw.buf = out; e->bw.cap = cap; e->dbg = dbg;
    return 0;
}
static void enc_free(enc_t *e) {
    for (int c = 0; c < 3; c++) { free(e->cur[c]); free(e->ref[c]); free(e->rec[c]); }
}

int m2v_oracle_encode_range(int VL, int Q, int mbw, int mbh, int P, const uint8_t *frames, long n0, long n1,
                            uint8_t *out, size_t cap, size_t *outlen, m2v_oracle_dbg *dbg) {
    enc_t e;
    if (enc_init(&e, VL, Q, mbw, mbh, P, out, cap, dbg) || n0 % (P + 1) != 0) { enc_free(&e); return -1; }
    size_t fsz = (size_t)e.W * e.H * 3;
    for (long n = n0; n < n1; n++) encode_frame(&e, frames + (size_t)(n - n0) * fsz, n);
    bw_align(&e.bw);
    *outlen = e.bw.pos;
    int ovf = e.bw.ovf;
    enc_free(&e);
    return ovf ? -1 : 0;
}

int m2v_oracle_encode(int XL, int YL, int VL, int Q, int xsize16, int ysize16, int P,
                      const uint8_t *frames, long nframes, long partial_px4,
                      uint8_t *out, size_t cap, size_t *outlen, m2v_oracle_dbg *dbg) {
    if (XL < 4 || XL > 7 || YL < 4 || YL > 7) return -1;
    int mbw = m2v_oracle_clamp16(xsize16, XL), mbh = m2v_oracle_clamp16(ysize16, YL);
    enc_t e;
    if (enc_init(&e, VL, Q, mbw, mbh, P, out, cap, dbg)) { enc_free(&e); return -1; }
    if (nframes + (partial_px4 > 0) <= 0) { enc_free(&e); return -1; }     /* never started: no output */
    size_t ysz = (size_t)e.W * e.H, fsz = ysz * 3;
    put_seq_header(&e.bw, e.W, e.H);
    for (long n = 0; n < nframes; n++) encode_frame(&e, frames + (size_t)n * fsz, n);
    if (partial_px4 > 0) {                                                   /* RTL:1036-1037,1049-1056 */
        uint8_t *pad = (uint8_t *)malloc(fsz);
        size_t npx = (size_t)partial_px4 * 4; if (npx > ysz) npx = ysz;
        memset(pad, 0, ysz); memset(pad + ysz, 0x80, 2 * ysz);
        const uint8_t *src = frames + (size_t)nframes * fsz;
        memcpy(pad, src, npx); memcpy(pad + ysz, src + ysz, npx); memcpy(pad + 2 * ysz, src + 2 * ysz, npx);
        encode_frame(&e, pad, nframes);
        free(pad);
    }
    bw_align(&e.bw);                                                          /* sequence end (RTL:2621-2628) */
    bw_put(&e.bw, 0x000001, 24); bw_put(&e.bw, 0xB7, 8);
    /* final flush: zero-pad to 32 bytes; an all-zero word if nothing remains (RTL:2932-2937) */
    size_t total = 32 * (e.bw.pos / 32 + 1);
    while (e.bw.pos < total) bw_put(&e.bw, 0, 8);
    *outlen = e.bw.pos;
    int ovf = e.bw.ovf;
    enc_free(&e);
    return ovf ? -1 : 0;
}
