// m2v_kernels.cu - sm_100a kernels of the MPEG-2 I/P macroblock path.
//
//   K1 mb_encode : one warp per macroblock.  4:4:4->4:2:0 chroma subsample, full-search SAD motion
//                  estimation, half-pel refinement, intra/inter decision, prediction, 6x 8x8 integer
//                  DCT, quantise, zig-zag, dequantise, Chen-Wang IDCT, reconstruction.
//   K1 is launched as persistent warps fed by TMA (cp.async.bulk.tensor + mbarrier, double buffered) that draw
//   their macroblocks from an atomic work counter.
//   K2 vlc       : one thread per macroblock (lanes = consecutive macroblocks of a slice).  run/level +
//                  VLC over the sparse levels; count pass (bit length per macroblock) and write pass.
//   K3 scans     : warp inclusive scans over the per-macroblock bit lengths -> offsets in slice, slice
//                  bytes -> offsets in frame, frame bytes -> offsets in the body; K4 writes the headers.
//
// Behaviour follows /root/reference/RTL/mpeg2encoder.v ("RTL") stage by stage; the citations say
// which lines each block reproduces.  All arithmetic is integer and must be bit-exact.
#include "m2v_kernels.cuh"
#include "m2v_tables.cuh"
#include <cuda.h>
#include <stdio.h>

#define FULL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// device tables
// ------------------------------------------------------------------------------------------------
__constant__ uint32_t c_vlc_motion[17];
__constant__ uint32_t c_vlc_cbp[64];
__constant__ uint32_t c_vlc_dcy[12];
__constant__ uint32_t c_vlc_dcc[12];
__device__ uint32_t d_vlc_ac[32 * M2V_AC_LEVELS];
// (Measured and rejected: the entry packed into 8 bytes (recip; W | off << 8 | zz << 24) - the I-frame kernel is bound by shared-memory
//  wavefronts and a 64-bit load takes half the wavefronts of a 128-bit one.  One I-frame launch under ncu 351 -> 339 us, but 1088 ->
//  1146 instructions per macroblock for the unpacking, and config 2 (I-only) in the bench 175.7 -> 173.7 Gpixel/s:
//  profiles/r02/k1_i_qentry8_experiment.txt.)
struct QEntry { uint32_t W, recip, off, zz; };          // per coefficient position i*8+j
__device__ QEntry d_qtab[4][64];                        // [Q_LEVEL-1]

cudaError_t m2v_upload_tables(int) {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_vlc_motion, M2V_VLC_MOTION, sizeof(M2V_VLC_MOTION)))) return e;
    if ((e = cudaMemcpyToSymbol(c_vlc_cbp, M2V_VLC_CBP, sizeof(M2V_VLC_CBP)))) return e;
    if ((e = cudaMemcpyToSymbol(c_vlc_dcy, M2V_VLC_DC_Y, sizeof(M2V_VLC_DC_Y)))) return e;
    if ((e = cudaMemcpyToSymbol(c_vlc_dcc, M2V_VLC_DC_C, sizeof(M2V_VLC_DC_C)))) return e;
    if ((e = cudaMemcpyToSymbol(d_vlc_ac, M2V_VLC_AC, sizeof(M2V_VLC_AC)))) return e;
    QEntry h[4][64];                                        // per call: several devices / threads may upload concurrently
    for (int q = 1; q <= 4; q++)
        for (int i = 0; i < 64; i++) {
            uint32_t w = M2V_INTRA_Q[i];
            h[q - 1][i].W = w;
            // floor(n / w) == umulhi(n, ceil(2^32 / w)) for every n < 2^15 (checked in tests/test_host_logic.py)
            h[q - 1][i].recip = (uint32_t)((0x100000000ull + w - 1) / w);
            h[q - 1][i].off = (w * ((3u << q) + 2u)) >> 3;            // RTL:2072
            h[q - 1][i].zz = M2V_ZIGZAG[i];
        }
    return cudaMemcpyToSymbol(d_qtab, h, sizeof(h));
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// sum of 4 byte-wise |a-b| plus c in one instruction (VABSDIFF4.U8.ACC)
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
__device__ __forceinline__ uint32_t exlo(uint32_t v) { return __byte_perm(v, 0, 0x4140); }   // bytes 0,1 -> halfwords
__device__ __forceinline__ uint32_t exhi(uint32_t v) { return __byte_perm(v, 0, 0x4342); }   // bytes 2,3 -> halfwords
__device__ __forceinline__ uint32_t pack4(uint32_t h0, uint32_t h1) { return __byte_perm(h0, h1, 0x6420); }
__device__ __forceinline__ uint32_t pack4h(uint32_t h0, uint32_t h1) { return __byte_perm(h0, h1, 0x7531); }   // the high byte of every halfword
// Residuals are staged as halfwords biased by +256 (always positive: the packed subtraction needs no per-halfword borrow
// handling - one IADD3 per two pixels).  The forward transform is linear and every row of the RTL's matrix except the first sums
// to zero, so the bias only shifts the first output of the row pass, by 64*8*256, which is subtracted there as an immediate.
#define RBIAS2 0x01000100u
#define RBIAS_ROW0 (64 * 8 * 256)
// mean2 of four byte pairs (RTL:750-757): (a + b + 1) >> 1 == (a | b) - ((a ^ b) >> 1) per byte; the subtraction never borrows
// across bytes.  One LOP3 less than the compiler's expansion of __vavgu4.
// The (a ^ b) & mask is one LOP3, spelled in PTX: left to itself the compiler shifts first and masks afterwards, one ALU-pipe
// instruction more per average.
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) {
    uint32_t t;
    asm("lop3.b32 %0, %1, %2, 0xFEFEFEFE, 0x28;" : "=r"(t) : "r"(a), "r"(b));
    return (a | b) - (t >> 1);
}

// forward 8-point transform with the RTL's 8-bit matrix (RTL:102-112), exact integer butterflies
// (Measured and rejected: additions written as a * k1 + b with k1 = 1 from the kernel arguments, to keep them off the ALU pipe - the
// compiler moves other instructions of the region to the ALU pipe instead; 2 ALU instructions less of 315.)
__device__ __forceinline__ void fdct8(const int x[8], int o[8]) {
    int e0 = x[0] + x[7], e1 = x[1] + x[6], e2 = x[2] + x[5], e3 = x[3] + x[4];
    int d0 = x[0] - x[7], d1 = x[1] - x[6], d2 = x[2] - x[5], d3 = x[3] - x[4];
    int ee0 = e0 + e3, ee1 = e1 + e2, eo0 = e0 - e3, eo1 = e1 - e2;
    o[0] = 64 * (ee0 + ee1);
    o[4] = 64 * (ee0 - ee1);
    o[2] = 84 * eo0 + 35 * eo1;
    o[6] = 35 * eo0 - 84 * eo1;
    o[1] = 89 * d0 + 75 * d1 + 50 * d2 + 18 * d3;
    o[3] = 75 * d0 - 18 * d1 - 89 * d2 - 50 * d3;
    o[5] = 50 * d0 - 89 * d1 + 18 * d2 + 75 * d3;
    o[7] = 18 * d0 - 50 * d1 + 75 * d2 - 89 * d3;
}

#define W1 2841u
#define W2 2676u
#define W3 2408u
#define W5 1609u
#define W6 1108u
#define W7 565u
// (Measured and rejected: issuing constant arithmetic shifts as IMAD.HI to move them to the FMA pipe made K1
// 2-4 % slower - IMAD.HI is not full rate.)
template <int S> __device__ __forceinline__ int asr(int v) { return v >> S; }
__device__ __forceinline__ int sx18(int v) { return asr<14>((int)((uint32_t)v << 14)); }

// Chen-Wang rows (RTL:849-904): 13-bit in, 18-bit out.  The RTL computes in 32-bit registers that
// wrap; unsigned arithmetic reproduces that without relying on signed overflow.
typedef uint32_t u32;
#define sra(v, s) asr<s>((int)(v))
__device__ __forceinline__ void idct_row(const int a[8], int r[8]) {
    u32 x0 = ((u32)a[0] << 11) + 128u, x1 = (u32)a[4] << 11, x2 = a[6], x3 = a[2], x4 = a[1], x5 = a[7], x6 = a[5], x7 = a[3], x8;
    x8 = W7 * (x4 + x5); x4 = x8 + (W1 - W7) * x4; x5 = x8 - (W1 + W7) * x5;
    x8 = W3 * (x6 + x7); x6 = x8 - (W3 - W5) * x6; x7 = x8 - (W3 + W5) * x7;
    x8 = x0 + x1; x0 = x0 - x1;
    x1 = W6 * (x3 + x2); x2 = x1 - (W2 + W6) * x2; x3 = x1 + (W2 - W6) * x3;
    x1 = x4 + x6; x4 = x4 - x6; x6 = x5 + x7; x5 = x5 - x7;
    x7 = x8 + x3; x8 = x8 - x3; x3 = x0 + x2; x0 = x0 - x2;
    x2 = (u32)sra(181u * (x4 + x5) + 128u, 8); x4 = (u32)sra(181u * (x4 - x5) + 128u, 8);
    r[0] = sx18(sra(x7 + x1, 8)); r[1] = sx18(sra(x3 + x2, 8)); r[2] = sx18(sra(x0 + x4, 8)); r[3] = sx18(sra(x8 + x6, 8));
    r[4] = sx18(sra(x8 - x6, 8)); r[5] = sx18(sra(x0 - x4, 8)); r[6] = sx18(sra(x3 - x2, 8)); r[7] = sx18(sra(x7 - x1, 8));
}
__device__ __forceinline__ int clip255(int v) { return v < -255 ? -255 : v > 255 ? 255 : v; }
// Chen-Wang columns (RTL:916-970): 18-bit in, 9-bit clipped out
__device__ __forceinline__ void idct_col(const int a[8], int o[8]) {
    u32 x0 = ((u32)a[0] << 8) + 8192u, x1 = (u32)a[4] << 8, x2 = a[6], x3 = a[2], x4 = a[1], x5 = a[7], x6 = a[5], x7 = a[3], x8;
    x8 = W7 * (x4 + x5) + 4u; x4 = (u32)sra(x8 + (W1 - W7) * x4, 3); x5 = (u32)sra(x8 - (W1 + W7) * x5, 3);
    x8 = W3 * (x6 + x7) + 4u; x6 = (u32)sra(x8 - (W3 - W5) * x6, 3); x7 = (u32)sra(x8 - (W3 + W5) * x7, 3);
    x8 = x0 + x1; x0 = x0 - x1;
    x1 = W6 * (x3 + x2) + 4u; x2 = (u32)sra(x1 - (W2 + W6) * x2, 3); x3 = (u32)sra(x1 + (W2 - W6) * x3, 3);
    x1 = x4 + x6; x4 = x4 - x6; x6 = x5 + x7; x5 = x5 - x7;
    x7 = x8 + x3; x8 = x8 - x3; x3 = x0 + x2; x0 = x0 - x2;
    x2 = (u32)sra(181u * (x4 + x5) + 128u, 8); x4 = (u32)sra(181u * (x4 - x5) + 128u, 8);
    o[0] = clip255(sra(x7 + x1, 14)); o[1] = clip255(sra(x3 + x2, 14)); o[2] = clip255(sra(x0 + x4, 14)); o[3] = clip255(sra(x8 + x6, 14));
    o[4] = clip255(sra(x8 - x6, 14)); o[5] = clip255(sra(x0 - x4, 14)); o[6] = clip255(sra(x3 - x2, 14)); o[7] = clip255(sra(x7 - x1, 14));
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
#define K1_WARPS 8
// Shared-memory strides chosen against bank conflicts (the I-frame kernel is bound by shared-memory wavefronts: ncu
// showed l1tex at 97 % with 38 % of the wavefronts caused by conflicts):
#define TROW 12                      // words per row of a transform scratch tile: 128-bit row accesses of 8 lanes hit 8 distinct bank quads
#define TSTR 104                     // words per scratch tile (8 rows x 12 + 8): the 4 tile groups' column accesses fall in distinct bank octets
#define RSTR 72                      // halfwords per level tile (64 + 8): the zig-zag scatter of the 4 tile groups is spread over the banks
#define PSTR 72                      // bytes per prediction tile (64 + 8): byte accesses of the 4 tile groups do not collide
// One pipeline stage = everything TMA brings in for one macroblock.  Every member is a dense TMA box
// and starts on a 128-byte boundary.  The I-frame instantiation needs no windows: its warps take 4.3 KB instead
// of 8.3 KB of shared memory, which (with 64 registers) lets a fourth CTA live on every SM - the I-frame kernel is
// bound by the latency of its shared-memory transposes, not by a pipe, so it is the resident warps that count.
template <bool PF> struct StageSmemT;
template <> struct __align__(128) StageSmemT<true> {
    uint32_t curY[16][4];            // current luma block, 16 rows x 16 B      } one box 16x16x3 over the planes
    uint32_t curU[16][4];            // current 4:4:4 U block                   } Y,U,V of the input frame
    uint32_t curV[16][4];            //                 V                        }
    // TMA needs the inner coordinate of a box to be a multiple of 16 bytes (probed: tools/tma_probe.cu),
    // so the windows start at the 16-byte boundary at or below the first byte that is needed.
    uint32_t winC[2][16][8];         // chroma windows: rows 8by-4..8by+11, 32 bytes from (8bx-8)&~15     (2 boxes 32x16)
    uint32_t winY[32][12];           // luma window: rows Y0-(R+1)..Y0+16+R, bytes X0-16..X0+31           (box 48 x (18+2R))
};
template <> struct __align__(128) StageSmemT<false> {
    uint32_t curY[16][4], curU[16][4], curV[16][4];
};
template <bool PF> struct WarpSmemT;
template <> struct __align__(128) WarpSmemT<true> {
    StageSmemT<true> st[2];          // double buffer: TMA fills st[k^1] while st[k] is being encoded
    uint32_t curC[2][8][2];          // current 4:2:0 chroma blocks
    int16_t res[8][RSTR];            // residual, later the zig-zag levels; tiles 6,7 are all-zero dummies that keep lanes
    uint8_t pred[8][PSTR];           // prediction, later the reconstruction          16..31 busy in the chroma round
};                                   // 8576 bytes; the two mbarriers of each warp live after the warps' areas
template <> struct __align__(128) WarpSmemT<false> {
    StageSmemT<false> st[2];
    uint32_t scratch[4 * TSTR];      // transform scratch (the P-frame kernel aliases it onto the dead windows)
    uint32_t curC[2][8][2];
    int16_t res[8][RSTR];
    uint8_t pred[8][PSTR];
};                                   // 5120 bytes
// the transform scratch (4 tile slots x TSTR words) aliases winC+winY of the stage being encoded:
// the windows are dead once the prediction has been formed.
static_assert(4 * TSTR * 4 <= sizeof(uint32_t) * (2 * 16 * 8 + 32 * 12), "scratch must fit in the window area");

struct K1Args {
    uint8_t *rec;                                           // reconstruction out: [G][fsz420]
    int16_t *coefs; uint32_t *mbinfo;
    int W, H, mbw, mbh, nmb, P, Q, t, CWp;                  // CWp = chroma row stride of the recon buffers (16-byte multiple)
    unsigned fsz420;                                        // bytes per reconstructed frame = W*H + 2*CWp*H/2
    unsigned total;                                         // ngops_t * nmb macroblocks in this launch
    int write_rec;                                          // 0 for the last frame of a GOP: its reconstruction is never read
    unsigned *ctr, *ctr_next;                               // work counter of this launch (zero on entry) / of the next one (zeroed here)
    uint32_t mw, mh;                                        // ceil(2^32/mbw), ceil(2^32/mbh): index -> (GOP, row, column) without divisions
    uint32_t k1024, k64, k16;                               // = 1024, 64, 16: opaque to the compiler (keep multiply-adds on the FMA pipe)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// A TMA that faults never completes its mbarrier: the wait is bounded (each try_wait already blocks for a hardware time slice) and
// traps, so a bad tensor map surfaces as a launch error (M2V_ECUDA) instead of a hung stream.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
// one lane of the (converged) warp, chosen by the hardware: the form the compiler turns into a single ELECT and a uniform branch
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// Persistent warps with DYNAMIC work distribution: a warp's first macroblock is its own global index, every
// further one comes from an atomic counter, fetched one macroblock ahead so that the TMA loads of the next
// macroblock are in flight while the current one is encoded and the atomic's round trip hides behind the encode.
// (With a static stride the warps of an SM sub-partition finish far apart - the scheduler favours some of them -
// and the tail runs at a fraction of the occupancy: ncu showed 4.5 of 6 resident warps active on average.)
template <int VL, bool PFRAME>
__global__ void __launch_bounds__(K1_WARPS * 32, PFRAME ? 3 : 4) k1_mb_encode(K1Args p, const __grid_constant__ CUtensorMap tm_in,
                                                               const __grid_constant__ CUtensorMap tm_refY,
                                                               const __grid_constant__ CUtensorMap tm_refC) {
    constexpr int R = 2 * VL;
    constexpr int WROWS = 18 + 2 * R;
    constexpr uint32_t TX_BYTES = 3 * 256 + (PFRAME ? 2 * 512 + 48 * WROWS : 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];     // keeps the .shared address space: LDS/STS, not generic LD/ST
    typedef WarpSmemT<PFRAME> WarpSmem;
    typedef StageSmemT<PFRAME> StageSmem;
    QEntry *const qt = reinterpret_cast<QEntry *>(smem_raw + sizeof(WarpSmem) * K1_WARPS);   // 64 entries after the warps' areas
    unsigned long long *const bars0 = reinterpret_cast<unsigned long long *>(qt + 64);                          // one mbarrier per stage and warp
    // the shuffle from lane 0 tells the compiler that the warp index is warp-uniform: the per-warp shared-memory addresses, the
    // mbarrier and the TMA issue then live in the uniform datapath instead of vector registers and waterfall loops
    const int lane = threadIdx.x & 31, warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x < 64) qt[threadIdx.x] = d_qtab[p.Q - 1][threadIdx.x];
    __syncthreads();
    WarpSmem &s = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
    unsigned long long *const bars = bars0 + 2 * warp;
    const unsigned gwarp = blockIdx.x * K1_WARPS + warp, nwarps = gridDim.x * K1_WARPS;
    if (gwarp >= p.total) return;
    const int W = p.W, CWp = p.CWp;
    const size_t ysz = (size_t)W * p.H;

    // macroblock index -> (GOP, block row, block column).  floor(n/d) == umulhi(n, ceil(2^32/d)) for n*d < 2^32:
    // d <= 128 and a launch has fewer than 2^25 macroblocks (m2v_launch_k1 checks; tests/test_host_logic.py)
    struct MbPos { int g, by, bx; };
    auto decode = [&](unsigned idx) {
        MbPos m;
        const unsigned rowg = __umulhi(idx, p.mw), g = __umulhi(rowg, p.mh);
        m.bx = (int)(idx - rowg * (unsigned)p.mbw); m.by = (int)(rowg - g * (unsigned)p.mbh); m.g = (int)g;
        return m;
    };
    // one elected lane arms the stage's mbarrier with the byte count and issues the TMA boxes
    auto issue = [&](const MbPos &m, int stg) {
        const int n = m.g * (p.P + 1) + p.t;
        StageSmem &S = s.st[stg];
        const uint32_t bar = smem_u32(&bars[stg]);
        mbar_expect_tx(bar, TX_BYTES);
        tma_load_4d(smem_u32(S.curY), &tm_in, m.bx * 16, m.by * 16, 0, n, bar);       // one 16x16x3 box: curY, curU, curV
        if constexpr (PFRAME) {
            // out-of-frame parts of a box are zero-filled by TMA; they only ever feed candidates the border
            // rule disables (RTL:1642-1645, 1757-1760).  (RTL:1350-1425, 1613-1629 fetch the same windows.)
            const int cx0 = (m.bx * 8 - 8) & ~15;
            tma_load_3d(smem_u32(S.winY), &tm_refY, m.bx * 16 - 16, m.by * 16 - (R + 1), m.g, bar);
            tma_load_4d(smem_u32(S.winC[0]), &tm_refC, cx0, m.by * 8 - 4, 0, m.g, bar);   // one 32x16x2 box: U and V windows
        }
    };
    for (int i = lane; i < 2 * RSTR / 2; i += 32) reinterpret_cast<uint32_t *>(&s.res[6][0])[i] = 0;    // dummy tiles: zero residual,
    for (int i = lane; i < 2 * PSTR / 4; i += 32) reinterpret_cast<uint32_t *>(&s.pred[6][0])[i] = 0;   // zero prediction
    // lane 0 draws the next index; the value is only looked at one macroblock later
    auto draw = [&]() { unsigned v = 0; if (lane == 0) v = nwarps + atomicAdd(p.ctr, 1u); return v; };
    MbPos cur = decode(gwarp);
    // Everything above is private to this launch (shared memory, arguments, a constant table).  From here on it reads what the launch
    // before it wrote (the reconstruction, the work counter it zeroed for us) and writes what that launch may still be using (the other
    // counter): wait for it to complete.  The next launch may start setting itself up behind this one at once.
    // (The mbarriers are initialised after the wait as well: with them before it compute-sanitizer's racecheck reports hazards between
    // mbarrier.init and the first expect_tx of the same lane - two grids resident at once - and the two instructions are not worth it.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.ctr_next = 0;                  // nobody touches the other counter during this launch
    if (lane == 0) {
        mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        issue(cur, 0);
    }
    unsigned nidx = __shfl_sync(FULL, draw(), 0);                              // index of the macroblock after this one
    uint32_t it = 0;                                                           // macroblocks done by this warp: stage it & 1, mbarrier parity (it >> 1) & 1
    int stg = 0;
#pragma unroll 1
    for (;; stg ^= 1) {
    const bool more = nidx < p.total;
    const MbPos nxt = decode(nidx);
    unsigned drawn = 0;
    if (more) {
        if (elect_one()) { fence_proxy_async(); issue(nxt, stg ^ 1); }
        drawn = draw();
    }
    const int g = cur.g, by = cur.by, bx = cur.bx;
    const unsigned mb = (unsigned)(by * p.mbw + bx);
    const unsigned n = (unsigned)(g * (p.P + 1) + p.t);     // frame index inside the batch
    const int Y0 = by * 16, X0 = bx * 16;
    StageSmem &S = s.st[stg];
    int32_t *tmp;
    if constexpr (PFRAME) tmp = reinterpret_cast<int32_t *>(&S.winC[0][0][0]); else tmp = reinterpret_cast<int32_t *>(s.scratch);
    mbar_wait(smem_u32(&bars[stg]), (it >> 1) & 1u);
    it++;

    // ---- current block: Y as is; U,V 4:4:4 -> 4:2:0 = mean2 of pixel pairs, then mean2 of the two
    //      horizontally subsampled rows (RTL:1086-1089, 1167-1170) -------------------------------
    {
        const int comp = lane >> 4, r = lane & 15;
        uint4 v = reinterpret_cast<const uint4 *>(S.curU)[lane];      // curV follows curU: row r of plane comp = 16-byte row `lane`
        uint32_t h0 = avg4(__byte_perm(v.x, v.y, 0x6420), __byte_perm(v.x, v.y, 0x7531));
        uint32_t h1 = avg4(__byte_perm(v.z, v.w, 0x6420), __byte_perm(v.z, v.w, 0x7531));
        uint32_t g0 = __shfl_xor_sync(FULL, h0, 1), g1 = __shfl_xor_sync(FULL, h1, 1);
        // mean2 is symmetric: the two lanes of a row pair hold the same result and both store it (same address, same value) - no branch
        *(uint2 *)s.curC[comp][r >> 1] = make_uint2(avg4(g0, h0), avg4(g1, h1));
    }
    __syncwarp();

    int inter = 0, mvx = 0, mvy = 0;
    if constexpr (PFRAME) {
        // ---- full-pel search (RTL:1634-1715).  lane = dxi + 16*half: candidate column dx = dxi-R,
        //      half = which 8 bytes of every 16-byte row.  Each lane keeps 2R+1 accumulators (one
        //      per dy) and walks the window rows once.
        int fmvy = 0, fmvx = 0;
        {
            const int half = lane >> 4;
            int dxi = lane & 15; if (dxi > 2 * R) dxi = 2 * R;
            uint32_t cur[16][2];
#pragma unroll
            for (int y = 0; y < 16; y++) { cur[y][0] = S.curY[y][half * 2]; cur[y][1] = S.curY[y][half * 2 + 1]; }
            uint32_t acc[2 * R + 1];
#pragma unroll
            for (int i = 0; i <= 2 * R; i++) acc[i] = 0;
            const int o = 16 + (dxi - R) + 8 * half, wi = o >> 2, sh = (o & 3) * 8;
#pragma unroll
            for (int wr = 0; wr < 16 + 2 * R; wr++) {       // reference row Y0 - R + wr  = window row wr+1
                uint32_t w0 = S.winY[wr + 1][wi], w1 = S.winY[wr + 1][wi + 1], w2 = S.winY[wr + 1][wi + 2];
                uint32_t a = fsr(w0, w1, sh), b = fsr(w1, w2, sh);
#pragma unroll
                for (int dyi = 0; dyi <= 2 * R; dyi++) {
                    const int cy = wr - dyi;
                    if (cy >= 0 && cy < 16) {
                        acc[dyi] = sad4(a, cur[cy][0], acc[dyi]);
                        acc[dyi] = sad4(b, cur[cy][1], acc[dyi]);
                    }
                }
            }
            // key = SAD<<10 | (R-dy)<<5 | (R-dx): minimum = smallest SAD, ties -> largest dy, then largest dx
            // (RTL:1696-1710).  SAD >= 4096 disqualifies (RTL:1669-1670)  <=>  key >= 1<<22, checked once at
            // the end.  The border rule (RTL:1642-1645) is warp-uniform in dy and per-lane in dx.
            // The key is linear in the two half-row sums, so each half forms acc*1024 + half of the dy term (one IMAD
            // with an immediate) BEFORE the exchange and the sum and the running minimum are one VIADDMNMX; the dx term
            // is the same for every dy of a lane and is added once after the minimum.
            uint32_t best = 0xFFFFFFFFu;
            auto fold = [&](int dyi) {
                const uint32_t k = acc[dyi] * p.k1024 + (uint32_t)((2 * R - dyi) << 4);   // IMAD (FMA pipe); a literal 1024 becomes an ALU-pipe LEA
                best = min(best, k + __shfl_xor_sync(FULL, k, 16));
            };
            if (by != 0) {                                   // dy < 0 is forbidden in the top block row (two warp-uniform branches)
#pragma unroll
                for (int dyi = 0; dyi < R; dyi++) fold(dyi);
            }
            fold(R);
            if (by != p.mbh - 1) {                           // dy > 0 is forbidden in the bottom block row
#pragma unroll
                for (int dyi = R + 1; dyi <= 2 * R; dyi++) fold(dyi);
            }
            best += (uint32_t)(2 * R - lane);                // R - dx, dx = lane - R; only lanes 0..2R (half 0, one per dx) take part below
            // lanes lo..hi are allowed: dx < 0 is forbidden in the first block column, dx > 0 in the last (warp-uniform bounds)
            const unsigned lo = bx == 0 ? R : 0, hi = bx == p.mbw - 1 ? R : 2 * R;
            if ((unsigned)lane - lo > hi - lo) best = 0xFFFFFFFFu;
            best = __reduce_min_sync(FULL, best);
            if (best < (1u << 22)) { fmvy = R - (int)((best >> 5) & 31); fmvx = R - (int)(best & 31); }
        }

        // ---- half-pel refinement + intra/inter decision (RTL:1743-1816).  lane = q + 4*j: the lane owns the 4-byte quarter q
        //      of the two rows a = 2j and b = 2j+1 and builds the shifted words and pair sums of the FOUR window rows
        //      a-1, a, b, b+1 once.  The vertical and diagonal half-pel planes are shared between neighbouring rows
        //      (candidate "down" of row a IS candidate "up" of row b), so three row pairs serve both rows: 4 row builds and
        //      3 pair builds per 2 rows, where a lane that owns one row needs 3 and 2 per row.
        const int q = lane & 3, j = lane >> 2;
        const uint32_t ca = S.curY[2 * j][q], cb = S.curY[2 * j + 1][q];
        uint32_t zz[4], V[3], DL[3], DR[3], c3a, c5a, c3b, c5b;       // full-pel rows; mean2 up/down; mean4 left/right diagonals; mean2 left/right
        {
            const int wr0 = R + fmvy + 2 * j;                          // window row of picture row a-1
            const int o = 15 + fmvx + 4 * q, wi = o >> 2, sh = (o & 3) * 8;   // byte of pixel x-1, x = 4q
            // The six bytes b0..b5 = pixels x-1..x+4 of a row give five horizontal pair sums S_i = b_i + b_(i+1); the left diagonals
            // use S0..S3, the right ones S1..S4.  Kept as packed halfwords by parity: P02 = (S0,S2), P13 = (S1,S3), P24 = (S2,S4) -
            // four unpacks and three adds per row.
            uint32_t P02[4], P13[4], P24[4];
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                const uint32_t *row = S.winY[wr0 + rr];
                const uint32_t w0 = row[wi], w1 = row[wi + 1], w2 = row[wi + 2];
                const uint32_t v0 = fsr(w0, w1, sh), v1 = fsr(w1, w2, sh);    // bytes x-1..x+2, x+3..x+6
                const uint32_t z = fsr(v0, v1, 8), zp = fsr(v0, v1, 16);      // bytes x..x+3, x+1..x+4
                zz[rr] = z;
                const uint32_t e0 = v0 & 0x00FF00FFu, o0 = __byte_perm(v0, 0, 0x4341);     // (b0,b2), (b1,b3)
                const uint32_t e1 = zp & 0x00FF00FFu, o1 = __byte_perm(zp, 0, 0x4341);     // (b2,b4), (b3,b5)
                P02[rr] = e0 + o0; P13[rr] = o0 + e1; P24[rr] = e1 + o1;
                if (rr == 1) { c3a = avg4(v0, z); c5a = avg4(z, zp); }            // mean2 left / right (RTL:1749)
                if (rr == 2) { c3b = avg4(v0, z); c5b = avg4(z, zp); }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                V[k] = avg4(zz[k], zz[k + 1]);                                  // mean2 up / down (RTL:1750)
                // mean4 (RTL:1751, 764) = (pair sum of the upper row + pair sum of the lower row + 1) >> 2.  (sum + 1) >> 2 of a packed
                // halfword = byte 1 of (sum + 1) * 64: one IMAD on the FMA pipe (the multiplier comes from the kernel arguments, a
                // literal 64 becomes an ALU-pipe shift) and the byte permute that interleaves the two parities picks bytes 1 and 3
                const uint32_t q02 = (P02[k] + P02[k + 1]) * p.k64 + 0x00400040u, q13 = (P13[k] + P13[k + 1]) * p.k64 + 0x00400040u;
                const uint32_t q24 = (P24[k] + P24[k + 1]) * p.k64 + 0x00400040u;
                DL[k] = __byte_perm(q02, q13, 0x7351);
                DR[k] = __byte_perm(q13, q24, 0x7351);
            }
        }
        int key[10];
        {
            // a disabled candidate (RTL:1757-1760) gets 0x800 per lane = 0x10000 per macroblock on top of its SAD through the
            // initial accumulator.  {f_over, f_diff}: any key >= 4096 loses against the intra key (<= 4095) and against every
            // valid candidate, and ties among keys >= 4096 never decide anything (RTL:1784-1785, 1795-1804).
            const uint32_t kxn = (bx == 0 || fmvx == -R) ? 0x800u : 0u, kxp = (bx == p.mbw - 1 || fmvx == R) ? 0x800u : 0u;
            const uint32_t kyn = (by == 0 || fmvy == -R) ? 0x800u : 0u, kyp = (by == p.mbh - 1 || fmvy == R) ? 0x800u : 0u;
            auto sadk = [&](uint32_t xa, uint32_t xb, uint32_t k0) { return (int)__reduce_add_sync(FULL, sad4(xa, ca, sad4(xb, cb, k0))); };
            key[0] = sadk(DL[0], DL[1], kxn + kyn); key[1] = sadk(V[0], V[1], kyn); key[2] = sadk(DR[0], DR[1], kxp + kyn);
            key[3] = sadk(c3a, c3b, kxn);           key[4] = sadk(zz[1], zz[2], 0); key[5] = sadk(c5a, c5b, kxp);
            key[6] = sadk(DL[1], DL[2], kxn + kyp); key[7] = sadk(V[1], V[2], kyp); key[8] = sadk(DR[1], DR[2], kxp + kyp);
            // intra key: pixel sum + sum|pixel-mean|, 16 bit, saturated to 4095 (RTL:1600,1662,1744,1776-1777,1791)
            uint32_t S = __reduce_add_sync(FULL, sad4(ca, 0, sad4(cb, 0, 0)));
            uint32_t m = (S >> 8) & 0xFF; m |= m << 8; m |= m << 16;
            uint32_t D = __reduce_add_sync(FULL, sad4(ca, m, sad4(cb, m, 0)));
            uint32_t T = (S + D) & 0xFFFF;
            key[9] = T < 4096u ? (int)T : 4095;
        }
        // find_min_in_10_values (RTL:804-840) is a tree whose comparisons resolve ties in the fixed order 8, 9, 4, 5, 6, 7, 0, 1, 2, 3:
        // the smallest of key*16 + rank is the same choice (tests/test_host_logic.py checks the equivalence
        // against the oracle's transcription of the tree, exhaustively on small values and on random ones)
        int w;
        {
            const uint32_t k16 = p.k16;
            uint32_t m = min(min((uint32_t)key[8] * k16 + 0u, (uint32_t)key[9] * k16 + 1u), (uint32_t)key[4] * k16 + 2u);
            m = min(min(m, (uint32_t)key[5] * k16 + 3u), (uint32_t)key[6] * k16 + 4u);
            m = min(min(m, (uint32_t)key[7] * k16 + 5u), (uint32_t)key[0] * k16 + 6u);
            m = min(min(m, (uint32_t)key[1] * k16 + 7u), (uint32_t)key[2] * k16 + 8u);
            m = min(m, (uint32_t)key[3] * k16 + 9u);
            w = (int)((0x3210765498ull >> (4 * (m & 15u))) & 15u);
        }
        inter = (w != 9);

        // ---- luma prediction + residual (RTL:1891-1897, 1980-2002).  w is warp-uniform: one branch instead of nine selects
        uint32_t pa = 0x80808080u, pb = 0x80808080u;
        int hy = 0, hx = 0;
        switch (w) {
            case 0: pa = DL[0]; pb = DL[1]; hy = -1; hx = -1; break;
            case 1: pa = V[0];  pb = V[1];  hy = -1; break;
            case 2: pa = DR[0]; pb = DR[1]; hy = -1; hx = 1; break;
            case 3: pa = c3a;   pb = c3b;   hx = -1; break;
            case 4: pa = zz[1]; pb = zz[2]; break;
            case 5: pa = c5a;   pb = c5b;   hx = 1; break;
            case 6: pa = DL[1]; pb = DL[2]; hy = 1; hx = -1; break;
            case 7: pa = V[1];  pb = V[2];  hy = 1; break;
            case 8: pa = DR[1]; pb = DR[2]; hy = 1; hx = 1; break;
            default: break;                                            // intra: predictor 128, vector 0
        }
        mvy = 2 * fmvy + hy; mvx = 2 * fmvx + hx;                      // RTL:1827-1828
        {
            const int tile = (j >> 2) * 2 + (q >> 1), e = ((2 * j) & 7) * 8 + 4 * (q & 1);
            *(uint32_t *)&s.pred[tile][e] = pa; *(uint32_t *)&s.pred[tile][e + 8] = pb;
            *(uint2 *)&s.res[tile][e] = make_uint2(exlo(ca) + RBIAS2 - exlo(pa), exhi(ca) + RBIAS2 - exhi(pa));
            *(uint2 *)&s.res[tile][e + 8] = make_uint2(exlo(cb) + RBIAS2 - exlo(pb), exhi(cb) + RBIAS2 - exhi(pb));
        }
        // ---- chroma prediction + residual (RTL:1847-1888, 1899-1916).  lane = comp*16 + y*2 + half.
        {
            const int comp = lane >> 4, cyy = (lane >> 1) & 7, ch = lane & 1;
            const uint32_t cc = s.curC[comp][cyy][ch];
            uint32_t pc = 0x80808080u;
            if (inter) {
                const int cyv = mvy >> 1, cxv = mvx >> 1;          // floor (RTL:1904-1910)
                const int fy = cyv >> 1, fx = cxv >> 1, oy = cyv & 1, ox = cxv & 1;
                const int row = 4 + cyy + fy, o = ((bx & 1) ? 8 : 16) + 4 * ch + fx, wi = o >> 2, sh = (o & 3) * 8;
                // bytes s..s+3 and s+1..s+4 of a word pair, s = o & 3 (the clamping funnel shift returns the high word at 32)
                const uint32_t u0 = S.winC[comp][row][wi], u1 = S.winC[comp][row][wi + 1], d0 = S.winC[comp][row + 1][wi], d1 = S.winC[comp][row + 1][wi + 1];
                const uint32_t a0 = fsr(u0, u1, sh), a1 = __funnelshift_rc(u0, u1, sh + 8), b0 = fsr(d0, d1, sh), b1 = __funnelshift_rc(d0, d1, sh + 8);
                if (oy && ox) {
                    uint32_t lo = exlo(a0) + exlo(a1) + exlo(b0) + exlo(b1), hi = exhi(a0) + exhi(a1) + exhi(b0) + exhi(b1);
                    pc = pack4h(lo * p.k64 + 0x00400040u, hi * p.k64 + 0x00400040u);
                } else if (ox) pc = avg4(a0, a1);
                else if (oy) pc = avg4(a0, b0);
                else pc = a0;
            }
            *(uint32_t *)&s.pred[4 + comp][cyy * 8 + ch * 4] = pc;
            *(uint2 *)&s.res[4 + comp][cyy * 8 + ch * 4] = make_uint2(exlo(cc) + RBIAS2 - exlo(pc), exhi(cc) + RBIAS2 - exhi(pc));
        }
    } else {
        // I-frame: every macroblock intra, predictor 128, vector 0 (RTL:1820-1825, 1894-1903)
        const int y = lane >> 1, half = lane & 1;
        const uint32_t c0 = S.curY[y][half * 2], c1 = S.curY[y][half * 2 + 1], pz = 0x80808080u;
        const int tile = (y >> 3) * 2 + half, r = y & 7;
        *(uint2 *)&s.pred[tile][r * 8] = make_uint2(pz, pz);
        constexpr uint32_t kz = RBIAS2 - 0x00800080u;             // biased residual against the predictor 128
        *(uint4 *)&s.res[tile][r * 8] = make_uint4(exlo(c0) + kz, exhi(c0) + kz, exlo(c1) + kz, exhi(c1) + kz);
        const int comp = lane >> 4, cyy = (lane >> 1) & 7, ch = lane & 1;
        const uint32_t cc = s.curC[comp][cyy][ch];
        *(uint32_t *)&s.pred[4 + comp][cyy * 8 + ch * 4] = pz;
        *(uint2 *)&s.res[4 + comp][cyy * 8 + ch * 4] = make_uint2(exlo(cc) + kz, exhi(cc) + kz);
    }
    __syncwarp();

    // ---- transform / quantise / scan / reconstruct.  Round 0: luma tiles on 32 lanes (tile = lane>>3,
    //      vector = lane&7); round 1: chroma tiles on lanes 0..15 while lanes 16..31 run the same code on two
    //      all-zero dummy tiles (no predicate, no divergence, warp syncs with the constant full mask - a run-time
    //      mask makes the compiler guard every sync with MATCH/REDUX/VOTE).  (RTL:2029-2077, 2128-2356, 2452-2467)
    // (Measured and rejected: THREE rounds of one pass each - rows of luma; rows of chroma beside columns of luma tiles 0,1; the other
    //  columns - which saves one of four forward passes and the dummy tiles: the per-round overhead (roles, branches, syncs) outweighs
    //  it, 1741 instead of 1665 instructions per macroblock, K1 +2.4 %.)
    // (Measured and rejected: unrolling the two rounds - the kernel no longer fits the instruction cache and slows down;
    //  predicating lanes 16..31 off in the chroma round, or syncing only the lower half-warp with __syncwarp(0xFFFF) held in a
    //  run-time mask - the compiler guards every such sync with MATCH/REDUX/VOTE.)
    int cbp = 0;
    const int Q = p.Q;
#pragma unroll 1
    for (int round = 0; round < 2; round++) {
        const int tile = round * 4 + (lane >> 3), v = lane & 7;
        int32_t *tt = tmp + (lane >> 3) * TSTR;
        int x[8], o[8];
        {                                                        // rows: A = R * DCTM^T (RTL:2029-2036)
            uint4 rv = *(const uint4 *)&s.res[tile][v * 8];
            x[0] = (int)(rv.x & 0xFFFF); x[1] = (int)(rv.x >> 16); x[2] = (int)(rv.y & 0xFFFF); x[3] = (int)(rv.y >> 16);
            x[4] = (int)(rv.z & 0xFFFF); x[5] = (int)(rv.z >> 16); x[6] = (int)(rv.w & 0xFFFF); x[7] = (int)(rv.w >> 16);
            fdct8(x, o);
            o[0] -= (tile < 6) ? RBIAS_ROW0 : 0;                 // the two dummy tiles stay all-zero (no bias to remove)
#pragma unroll
            for (int j = 0; j < 8; j++) tt[v * TROW + j] = o[j];
        }
        __syncwarp();
        bool nzl = false, maybe = true;
        {                                                        // columns: B = DCTM * A (RTL:2054-2057)
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = tt[k * TROW + v];
            fdct8(x, o);
            if (inter) {
                // Inter levels are (|C|+2)>>(4+Q) with C = (B+2048)>>12: zero whenever |B| < 4096*(2^(4+Q)-2) - 2048.
                // Most columns of a well-predicted tile are entirely below that, so the zig-zag tile is
                // zero-filled with one vector store per lane and only columns that may hold a level are
                // quantised coefficient by coefficient.
                const int thr = 4096 * ((1 << (4 + Q)) - 2) - 2048;
                const int mx = max(max(max(o[0], o[1]), max(o[2], o[3])), max(max(o[4], o[5]), max(o[6], o[7])));
                const int mn = min(min(min(o[0], o[1]), min(o[2], o[3])), min(min(o[4], o[5]), min(o[6], o[7])));
                maybe = (mx >= thr) || (mn <= -thr);
                *(uint4 *)&s.res[tile][v * 8] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncwarp();                                        // zero fill complete before any level is scattered
        {                                                        // round, quantise, scan, dequantise
            const int b00 = o[0];
            // |C| <= 255*512*512/4096 = 16320, so every level is < 2047 and the RTL's clip (RTL:2075) never
            // acts; the branch on `inter` is warp-uniform and hoisted out of the coefficient loop.
            if (inter) {
                if (maybe) {
                    // the same bound per coefficient: -thr < B < thr  =>  level 0 (and dequantised value 0), so
                    // only the few coefficients outside it run the quantiser (a real branch, not predication)
                    const int thr = 4096 * ((1 << (4 + Q)) - 2) - 2048;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        int dq = 0;
                        if ((uint32_t)(o[i] + thr - 1) >= (uint32_t)(2 * thr - 1)) {
                            const int C = asr<12>(o[i] + 2048);                      // RTL:2058
                            const int yq = (abs(C) + 2) >> (4 + Q);                  // RTL:2070
                            const int sgn = (C >> 31) | 1;
                            if (yq) { s.res[tile][qt[i * 8 + v].zz] = (int16_t)(yq * sgn); nzl = true; }   // zig-zag (RTL:2464)
                            dq = min(yq ? ((2 * yq + 1) << Q) : 0, 2047) * sgn;      // RTL:2134-2137
                        }
                        o[i] = dq;
                    }
                }                                                // else: no level in this column, o[] is not used again
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const QEntry qe = qt[i * 8 + v];
                    const int C = asr<12>(o[i] + 2048);
                    const int yq = (int)__umulhi((uint32_t)(abs(C) + (int)qe.off) >> Q, qe.recip);   // RTL:2072 (exact division)
                    const int q = yq * ((C >> 31) | 1);
                    s.res[tile][qe.zz] = (int16_t)q;
                    int xq = q * (int)qe.W;                                          // RTL:2139-2144, >>> = floor
                    xq = (Q >= 3) ? (xq << (Q - 3)) : (xq >> (3 - Q));
                    o[i] = max(-2047, min(2047, xq));
                }
                if (v == 0) {                                                        // DC (RTL:2074, 2146)
                    const int C = (b00 + 2048) >> 12, a = abs(C);
                    const int yq = (a >> 4) + ((a >> 3) & 1);
                    const int q = C < 0 ? -yq : yq;
                    s.res[tile][0] = (int16_t)q;
                    o[0] = 2 * q;
                }
                nzl = true;
            }
        }
        // cbp bit of this lane's tile (Y00 = 32 ... V = 1; 0 for the dummy tiles 6 and 7): one REDUX.OR gives the coded tiles
        const uint32_t tb = 0x20u >> tile;
        const uint32_t nzm = __reduce_or_sync(FULL, nzl ? tb : 0u);
        cbp |= (int)nzm;
        // (Measured and rejected: a second REDUX.OR collecting, per tile, which 16-level chunks of the zig-zag scan hold a level, so that
        //  K2 loads only those: K2 -0.03 ms of 0.81, K1 +0.19 ms of 8.25 per step - K2's own all-zero test of a loaded chunk is as good.)
        // A tile whose levels are all zero reconstructs to exactly the prediction (all-zero input gives
        // (128)>>8 = 0 after the row pass and (8192)>>14 = 0 after the column pass), so its inverse
        // transform is skipped; the decision is per 8-lane group, and when no tile of the round holds a level
        // (the usual case in a well-predicted picture) the whole inverse part is one uniform branch.
        if (nzm) {
            const bool inv = (nzm & tb) != 0;
            __syncwarp();                                    // all column reads of A done before overwrite
            if (inv) {
                if (inter && !maybe) {
#pragma unroll
                    for (int i = 0; i < 8; i++) o[i] = 0;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) tt[i * TROW + v] = o[i];
            }
            __syncwarp();
            if (inv) {                                           // inverse rows (in place)
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = tt[v * TROW + j];
                idct_row(x, o);
#pragma unroll
                for (int j = 0; j < 8; j++) tt[v * TROW + j] = o[j];
            }
            __syncwarp();
            if (inv) {                                           // inverse columns, add prediction, clip (RTL:2352)
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = tt[k * TROW + v];
                idct_col(x, o);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    int r = (int)s.pred[tile][i * 8 + v] + o[i];
                    s.pred[tile][i * 8 + v] = (uint8_t)max(0, min(255, r));
                }
            }
        }
        __syncwarp();
    }

    // ---- outputs: reconstruction (next frame's reference), levels, record -----------------------
    {
        if (p.write_rec) {                                       // warp-uniform: the last frame of a GOP is nobody's reference
            uint8_t *oY = p.rec + (size_t)g * p.fsz420;
            const int y = lane >> 1, half = lane & 1, tile = (y >> 3) * 2 + half;
            *(uint2 *)(oY + (unsigned)((Y0 + y) * W + X0 + 8 * half)) = *(const uint2 *)&s.pred[tile][(y & 7) * 8];
            if (lane < 16) {
                const int comp = lane >> 3, cyy = lane & 7;
                *(uint2 *)(oY + (unsigned)(ysz + (comp * (p.H >> 1) + by * 8 + cyy) * CWp + bx * 8)) = *(const uint2 *)&s.pred[4 + comp][cyy * 8];
            }
        }
        const unsigned mbi = n * (unsigned)p.nmb + mb;           // < 2^25 (m2v_launch_k1 bounds the chunk)
        uint2 *dst = (uint2 *)(p.coefs + (size_t)mbi * 384);
        // a tile whose cbp bit is clear holds no level and K2 never reads it (RTL:2799, 2804, 2828): not written
#pragma unroll
        for (int k = 0; k < 3; k++)
            if ((cbp << (2 * k + (lane >> 4))) & 32) dst[lane + 32 * k] = *(const uint2 *)&s.res[2 * k + (lane >> 4)][(lane & 15) * 4];
        const uint32_t info = M2V_INFO(inter, mvx, mvy, cbp);     // warp-uniform: packed in the uniform datapath, stored by one lane
        if (elect_one()) p.mbinfo[mbi] = info;
    }
    __syncwarp();
    if (!more) break;
    cur = nxt;
    nidx = __shfl_sync(FULL, drawn, 0);
    }   // persistent loop
}

// TMA descriptors for one batch: the 4:4:4 input [F][3][H][W], and per reconstruction buffer the luma
// planes [G][H][W] and the chroma planes [G][2][H/2][CWp].  cuTensorMapEncodeTiled is fetched through the
// runtime so that the library does not link against libcuda directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
static bool make_map(CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool m2v_make_tmaps(M2VBatch &b) {
    const cuuint64_t W = b.g.W, H = b.g.H, ysz = W * H, CWp = b.CWp, CH = H / 2, G = (b.F + b.g.P) / (b.g.P + 1);
    const int WROWS = 18 + 4 * b.g.VL;
    {
        const cuuint64_t d[4] = {W, H, 3, (cuuint64_t)b.F}, st[3] = {W, ysz, 3 * ysz};
        const cuuint32_t box[4] = {16, 16, 3, 1};
        if (!make_map(&b.tm_in, (void *)b.in, 4, d, st, box)) return false;
    }
    for (int k = 0; k < 2; k++) {
        const cuuint64_t dy[3] = {W, H, G}, sy[2] = {W, b.fsz420};
        const cuuint32_t boxy[3] = {48, (cuuint32_t)WROWS, 1};
        if (!make_map(&b.tm_refY[k], b.recon[k], 3, dy, sy, boxy)) return false;
        const cuuint64_t dc[4] = {CWp, CH, 2, G}, sc[3] = {CWp, CWp * CH, b.fsz420};
        const cuuint32_t boxc[4] = {32, 16, 2, 1};
        if (!make_map(&b.tm_refC[k], b.recon[k] + ysz, 4, dc, sc, boxc)) return false;
    }
    return true;
}

template <int VL, bool PF> static constexpr size_t k1_smem() { return sizeof(WarpSmemT<PF>) * K1_WARPS + 64 * sizeof(QEntry) + 16 * K1_WARPS; }

// Per-DEVICE launch configuration (the dynamic shared-memory attribute and the occupancy are properties of a device's
// context): called by m2v_create on every device a handle owns, with that device current.  Returns the persistent grid
// sizes = every CTA the device can hold at once.
template <int VL, bool PF> static cudaError_t k1_setup_t(int *grid_cap) {
    cudaError_t e = cudaFuncSetAttribute(k1_mb_encode<VL, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem<VL, PF>());
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess || (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    // P-frame kernel: 3 CTAs of 8 warps per SM (<= 80 registers, 8.4 KB of shared memory per warp; 4x7 warps at 72
    // registers measured slower); I-frame kernel: 4 CTAs (64 registers, 5.1 KB per warp)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_mb_encode<VL, PF>, K1_WARPS * 32, k1_smem<VL, PF>()) != cudaSuccess || per_sm < 1) per_sm = PF ? 3 : 4;
    *grid_cap = sms * per_sm;
    return cudaSuccess;
}
cudaError_t m2v_k1_setup(int VL, int *grid_cap_i, int *grid_cap_p) {
    cudaError_t e = k1_setup_t<1, false>(grid_cap_i);
    if (e != cudaSuccess) return e;
    switch (VL) {
        case 1: return k1_setup_t<1, true>(grid_cap_p);
        case 2: return k1_setup_t<2, true>(grid_cap_p);
        default: return k1_setup_t<3, true>(grid_cap_p);
    }
}

template <int VL, bool PF>
static void launch_k1_t(const K1Args &a, const M2VBatch &b, int refk, cudaStream_t st) {
    unsigned grid = (a.total + K1_WARPS - 1) / K1_WARPS;
    const unsigned cap = (unsigned)(PF ? b.k1_grid_cap_p : b.k1_grid_cap_i);
    if (grid > cap) grid = cap;
    // Programmatic dependent launch: the grid may be set up (CTAs resident, quantiser table and mbarriers in shared memory) behind the
    // tail of the launch before it; it touches global memory only after griddepcontrol.wait, i.e. after that launch has completed.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(K1_WARPS * 32); cfg.dynamicSmemBytes = k1_smem<VL, PF>(); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k1_mb_encode<VL, PF>, a, b.tm_in, b.tm_refY[refk], b.tm_refC[refk]);
}

void m2v_launch_k1(const M2VBatch &b, int t, long ngops_t, unsigned seq, cudaStream_t st) {
    K1Args a;
    a.rec = b.recon[t & 1];
    a.coefs = b.coefs; a.mbinfo = b.mbinfo;
    a.W = b.g.W; a.H = b.g.H; a.mbw = b.g.mbw; a.mbh = b.g.mbh; a.nmb = b.g.nmb; a.P = b.g.P; a.Q = b.g.Q; a.t = t;
    a.CWp = b.CWp; a.fsz420 = (unsigned)b.fsz420;
    a.total = (unsigned)(ngops_t * b.g.nmb);
    a.write_rec = t < b.g.P;
    a.ctr = b.k1_ctr + (seq & 1); a.ctr_next = b.k1_ctr + ((seq & 1) ^ 1);
    a.k1024 = 1024u; a.k64 = 64u; a.k16 = 16u;
    a.mw = (uint32_t)((0x100000000ull + b.g.mbw - 1) / b.g.mbw); a.mh = (uint32_t)((0x100000000ull + b.g.mbh - 1) / b.g.mbh);
    const int refk = (t & 1) ^ 1;
    if (t == 0) { launch_k1_t<1, false>(a, b, refk, st); return; }
    switch (b.g.VL) {
        case 1: launch_k1_t<1, true>(a, b, refk, st); break;
        case 2: launch_k1_t<2, true>(a, b, refk, st); break;
        default: launch_k1_t<3, true>(a, b, refk, st); break;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: entropy layer, one macroblock per thread (RTL:2718-2847)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_bits(uint32_t *words, unsigned long long pos, uint32_t code, int len) {
    if (len <= 0) return;
    const unsigned long long w = pos >> 5;
    const int o = (int)(pos & 31);
    const unsigned long long v = (unsigned long long)code << (64 - len - o);
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    atomicOr(&words[w], __byte_perm(hi, 0, 0x0123));
    if (lo) atomicOr(&words[w + 1], __byte_perm(lo, 0, 0x0123));
}

// code for (level v != 0, run) - RTL:2525-2547.  Returns (len<<24 | code) with the sign included.
__device__ __forceinline__ void ac_code(int v, int run, uint32_t &code, int &len) {
    const int m = abs(v) - 1;
    // The (run, level) pairs the RTL codes from its two tables (RTL:2533-2541: run 0 m<40, run 1 m<18, run 2 m<5, run 3 m<4,
    // run<=6 m<3, run<=16 m<2, run<=31 m<1) are exactly the populated entries of Table B-14 (tests/test_host_logic.py),
    // so one bounds check and a zero test of the entry replace the chain of range comparisons.
    uint32_t e = 0;
    if (m < M2V_AC_LEVELS && run < 32) e = __ldg(&d_vlc_ac[run * M2V_AC_LEVELS + m]);
    if (e) {
        code = ((e & 0xFFFF) << 1) | (v < 0);
        len = (int)(e >> 16) + 1;
    } else {
        code = (1u << 18) | ((uint32_t)(run & 63) << 12) | ((uint32_t)v & 0xFFF);
        len = 24;
    }
}

struct K2Args {
    const int16_t *coefs; const uint32_t *mbinfo;
    uint32_t *mb_bits; uint32_t *mb_code; const uint32_t *mb_off; const uint32_t *slice_off; const unsigned long long *frame_off;
    uint32_t *out;
    int mbw, mbh, nmb, P; long n0; long total;
    long F; unsigned long long cap_words;                     // write pass: frames in the batch, capacity of `out` in 32-bit words
};

// K2: ONE THREAD PER MACROBLOCK, lanes of a warp = consecutive macroblocks (RTL:2718-2847).
// The levels are sparse (typically a handful of non-zero values in the 384 of a macroblock), so the
// work is latency/issue bound, not data bound: a thread reads its record, skips tiles whose cbp bit is
// clear, and for a coded tile walks 16-level chunks by find-first-set over a non-zero mask.  A warp thus
// retires 32 macroblocks in about the time the busiest of them needs.  Codes are accumulated MSB-first in
// a 64-bit window and ORed into the zero-initialised stream one 32-bit word at a time (RED.OR).
struct BitAcc {                       // MSB-first accumulator aligned to 32-bit stream words
    uint32_t *w; unsigned long long acc; int n;
    __device__ __forceinline__ void start(uint32_t *words, unsigned long long bitpos) {
        w = words + (bitpos >> 5); n = (int)(bitpos & 31); acc = 0;      // leading zero bits are harmless: we OR
    }
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc = (acc << len) | code; n += len;
        if (n >= 32) {
            atomicOr(w++, __byte_perm((uint32_t)(acc >> (n - 32)), 0, 0x0123));
            n -= 32; acc &= (1ull << n) - 1;
        }
    }
    __device__ __forceinline__ void flush() { if (n) atomicOr(w, __byte_perm((uint32_t)(acc << (32 - n)), 0, 0x0123)); }
};
// Count pass: besides the length it keeps the macroblock's own bitstring (MSB first, starting at bit 0) in a slot of M2V_MB_SLOT words,
// interleaved over the 32 macroblocks of a warp so that lanes store and load the same word index side by side.  The write pass then only
// shifts those words to the macroblock's place in the stream; a macroblock longer than the slot is walked a second time (BitAcc).
struct BitLocal {
    uint32_t *w; unsigned long long acc; int n, total, nw;
    __device__ __forceinline__ void start(uint32_t *slot) { w = slot; acc = 0; n = 0; total = 0; nw = 0; }
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc = (acc << len) | code; n += len; total += len;
        if (n >= 32) {
            if (nw < M2V_MB_SLOT) w[nw * 32] = (uint32_t)(acc >> (n - 32));
            nw++; n -= 32; acc &= (1ull << n) - 1;
        }
    }
    __device__ __forceinline__ void finish() { if (n && nw < M2V_MB_SLOT) w[nw * 32] = (uint32_t)(acc << (32 - n)); }
};

template <typename Emit>
__device__ __forceinline__ void mb_syntax(Emit &out, const int16_t *__restrict__ zz, int k, uint32_t info, uint32_t li, bool has_left) {
    const int inter = info & 1, mvx = (int8_t)(info >> 8), mvy = (int8_t)(info >> 16), cbp = (info >> 24) & 63;
    // predictors from the left neighbour; reset at slice start (RTL:2713-2715), DC reset by an inter
    // macroblock (RTL:2786-2792), PMV reset by an intra macroblock (RTL:2771-2773)
    int pmvx = 0, pmvy = 0, dcp[3] = {0, 0, 0};
    if (has_left) {
        if (li & 1) { pmvx = (int8_t)(li >> 8); pmvy = (int8_t)(li >> 16); }
        else if (!inter) { dcp[0] = zz[-384 + 3 * 64]; dcp[1] = zz[-384 + 4 * 64]; dcp[2] = zz[-384 + 5 * 64]; }
    }
    // ---- macroblock header (RTL:2722-2767) ----
    if (!inter && k != 0) out.put(0x23u, 6); else if (inter && cbp == 0) out.put(0x09u, 4); else out.put(0x03u, 2);
    if (inter) {
#pragma unroll
        for (int comp = 0; comp < 2; comp++) {
            int d = comp ? mvy - pmvy : mvx - pmvx;
            if (d > 15) d -= 32; else if (d < -16) d += 32;
            const uint32_t e = c_vlc_motion[abs(d)];
            uint32_t c = e & 0xFFFF; int l = (int)(e >> 16);
            if (d != 0) { c = (c << 1) | (d < 0); l++; }
            out.put(c, l);
        }
        const uint32_t e = c_vlc_cbp[cbp];
        out.put(e & 0xFFFF, (int)(e >> 16));
    }
    // ---- tiles (RTL:2777-2847) ----
    int prev_dc = 0;
#pragma unroll 1
    for (int t = 0; t < 6; t++) {
        if (!((cbp >> (5 - t)) & 1)) continue;                    // inter tile without a level: nothing (RTL:2799,2804,2828)
        const uint4 *src = (const uint4 *)(zz + t * 64);
        int prevpos = inter ? -1 : 0;                             // the intra DC slot anchors the first run (RTL:2824)
#pragma unroll 1
        for (int ch = 0; ch < 4; ch++) {
            const uint4 a = __ldg(src + 2 * ch), b = __ldg(src + 2 * ch + 1);
            const uint32_t w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w, w4 = b.x, w5 = b.y, w6 = b.z, w7 = b.w;
            // most 16-level chunks of a coded tile are empty (the levels sit at the low frequencies): skip them before
            // building the mask.  The first chunk of an intra tile always codes its DC (RTL:2808-2821).
            if ((((w0 | w1) | (w2 | w3)) | ((w4 | w5) | (w6 | w7))) == 0 && (inter || ch != 0)) continue;
            uint32_t m = 0;
            m |= ((w0 & 0xFFFFu) ? 1u : 0u) | ((w0 >> 16) ? 2u : 0u);
            m |= ((w1 & 0xFFFFu) ? 4u : 0u) | ((w1 >> 16) ? 8u : 0u);
            m |= ((w2 & 0xFFFFu) ? 16u : 0u) | ((w2 >> 16) ? 32u : 0u);
            m |= ((w3 & 0xFFFFu) ? 64u : 0u) | ((w3 >> 16) ? 128u : 0u);
            m |= ((w4 & 0xFFFFu) ? 256u : 0u) | ((w4 >> 16) ? 512u : 0u);
            m |= ((w5 & 0xFFFFu) ? 1024u : 0u) | ((w5 >> 16) ? 2048u : 0u);
            m |= ((w6 & 0xFFFFu) ? 4096u : 0u) | ((w6 >> 16) ? 8192u : 0u);
            m |= ((w7 & 0xFFFFu) ? 16384u : 0u) | ((w7 >> 16) ? 32768u : 0u);
            if (ch == 0) {
                const int dc = (int16_t)(w0 & 0xFFFF);
                if (!inter) {                                     // intra DC (RTL:2808-2821)
                    const int pred = (t >= 1 && t <= 3) ? prev_dc : dcp[t < 4 ? 0 : t - 3];
                    const int diff = dc - pred, ad = abs(diff), size = 32 - __clz(ad);
                    const uint32_t e = t < 4 ? c_vlc_dcy[size] : c_vlc_dcc[size];
                    const uint32_t db = (uint32_t)(diff < 0 ? diff + (1 << size) - 1 : diff) & ((1u << size) - 1);
                    out.put(((e & 0xFFFF) << size) | db, (int)(e >> 16) + size);
                    m &= ~1u;
                    prev_dc = dc;
                } else if (dc == 1 || dc == -1) {                 // first coefficient +-1 of an inter tile: '1s' (RTL:2798-2802)
                    out.put(2u | (dc < 0), 2);
                    m &= ~1u; prevpos = 0;
                }
            }
            while (m) {
                const int b2 = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t lo = (b2 & 4) ? ((b2 & 2) ? w3 : w2) : ((b2 & 2) ? w1 : w0);
                const uint32_t hi = (b2 & 4) ? ((b2 & 2) ? w7 : w6) : ((b2 & 2) ? w5 : w4);
                const uint32_t ws = (b2 & 8) ? hi : lo;
                const int v = (b2 & 1) ? ((int32_t)ws >> 16) : (int)(int16_t)(ws & 0xFFFF);
                const int pos = 16 * ch + b2;
                uint32_t c; int l;
                ac_code(v, pos - prevpos - 1, c, l);              // RTL:2825-2833
                prevpos = pos;
                out.put(c, l);
            }
        }
        out.put(2u, 2);                                           // end of block (RTL:2835, 2897-2900)
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k2_vlc(K2Args p) {
    const long gw = (long)blockIdx.x * 128 + threadIdx.x;
    if (gw >= p.total) return;
    if (WRITE && M2V_BODY_WORDS(__ldg(&p.frame_off[p.F])) > p.cap_words) return;   // body larger than the buffer: the host re-runs the batch
    const long f = gw / p.nmb;
    const int mb = (int)(gw - f * p.nmb), by = mb / p.mbw, bx = mb - by * p.mbw;
    const int k = (int)((p.n0 + f) % (p.P + 1));
    const uint32_t info = __ldg(&p.mbinfo[gw]);
    const uint32_t li = bx > 0 ? __ldg(&p.mbinfo[gw - 1]) : 0u;
    const int16_t *zz = p.coefs + (size_t)gw * 384;
    uint32_t *slot = p.mb_code + ((size_t)(gw >> 5) * M2V_MB_SLOT) * 32 + (gw & 31);
    if (!WRITE) {
        BitLocal bl; bl.start(slot);
        mb_syntax(bl, zz, k, info, li, bx > 0);
        bl.finish();
        p.mb_bits[gw] = (uint32_t)bl.total;
    } else {
        const int hdr = (k == 0) ? 25 : 18;
        const unsigned long long pos = 8ull * (__ldg(&p.frame_off[f]) + hdr + __ldg(&p.slice_off[f * p.mbh + by])) + __ldg(&p.mb_off[gw]);
        const int nw = (int)((p.mb_bits[gw] + 31) >> 5);
        if (nw <= M2V_MB_SLOT) {                                  // the usual case: shift the cached words into place
            uint32_t *W = p.out + (pos >> 5);
            const int o = (int)(pos & 31);
            uint32_t prev = 0;
            for (int i = 0; i < nw; i++) {
                const uint32_t L = slot[i * 32];
                const uint32_t v = __funnelshift_r(L, prev, o);      // (prev << (32-o)) | (L >> o)
                if (v) atomicOr(W + i, __byte_perm(v, 0, 0x0123));
                prev = L;
            }
            const uint32_t v = __funnelshift_r(0u, prev, o);
            if (v) atomicOr(W + nw, __byte_perm(v, 0, 0x0123));
        } else {
            BitAcc bw; bw.start(p.out, pos);
            mb_syntax(bw, zz, k, info, li, bx > 0);
            bw.flush();
        }
    }
}

void m2v_launch_k2(const M2VBatch &b, bool write, cudaStream_t st) {
    K2Args a;
    a.coefs = b.coefs; a.mbinfo = b.mbinfo; a.mb_bits = b.mb_bits; a.mb_code = b.mb_code; a.mb_off = b.mb_off; a.slice_off = b.slice_off;
    a.frame_off = b.frame_off; a.out = b.out_words;
    a.mbw = b.g.mbw; a.mbh = b.g.mbh; a.nmb = b.g.nmb; a.P = b.g.P; a.n0 = b.n0; a.total = b.F * b.g.nmb;
    a.F = b.F; a.cap_words = b.out_cap_words;
    const unsigned grid = (unsigned)((a.total + 127) / 128);
    if (write) k2_vlc<true><<<grid, 128, 0, st>>>(a); else k2_vlc<false><<<grid, 128, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// K3: scans.  k3_frame: one CTA per frame - per-slice inclusive scans of macroblock bit lengths,
// slice byte lengths (38-bit header + macroblocks, zero padded to a byte: RTL:2704-2710, 2940-2943),
// then the scan over the frame's slices.  k3_batch: one CTA - scan over the frames.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_frame(const uint32_t *mb_bits, uint32_t *mb_off, uint32_t *slice_off,
                                                uint32_t *frame_bytes, int mbw, int mbh) {
    __shared__ uint32_t sbytes[128];
    const long f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sl = warp; sl < mbh; sl += 8) {
        const size_t base = ((size_t)f * mbh + sl) * mbw;
        uint32_t run = 38;                                          // slice header bits
        for (int c0 = 0; c0 < mbw; c0 += 32) {
            const int i = c0 + lane;
            const uint32_t v = i < mbw ? mb_bits[base + i] : 0;
            uint32_t sc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t nb = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc += nb; }
            if (i < mbw) mb_off[base + i] = run + sc - v;
            run += __shfl_sync(FULL, sc, 31);
        }
        if (lane == 0) sbytes[sl] = (run + 7) >> 3;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t run = 0;
        for (int c0 = 0; c0 < mbh; c0 += 32) {
            const int i = c0 + lane;
            const uint32_t v = i < mbh ? sbytes[i] : 0;
            uint32_t sc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t nb = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc += nb; }
            if (i < mbh) slice_off[f * mbh + i] = run + sc - v;
            run += __shfl_sync(FULL, sc, 31);
        }
        if (lane == 0) frame_bytes[f] = run;
    }
}

__global__ void __launch_bounds__(1024) k3_batch(const uint32_t *frame_bytes, unsigned long long *frame_off, long F, long n0, int P) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long c0 = 0; c0 < F; c0 += 1024) {
        const long i = c0 + threadIdx.x;
        unsigned long long v = 0;
        if (i < F) v = (unsigned long long)frame_bytes[i] + (((n0 + i) % (P + 1)) == 0 ? 25 : 18);   // GOP 8 B + picture 17 B | picture 18 B
        unsigned long long sc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long nb = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc += nb; }
        if (lane == 31) wsum[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], ws = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned long long nb = __shfl_up_sync(FULL, ws, d); if (lane >= d) ws += nb; }
            wsum[lane] = ws - w;
        }
        __syncthreads();
        const unsigned long long base = carry + wsum[warp];
        if (i < F) frame_off[i] = base + sc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = base + sc;
        __syncthreads();
    }
    if (threadIdx.x == 0) frame_off[F] = carry;
}

void m2v_launch_k3_scan(const M2VBatch &b, cudaStream_t st) {
    k3_frame<<<(unsigned)b.F, 256, 0, st>>>(b.mb_bits, b.mb_off, b.slice_off, b.frame_bytes, b.g.mbw, b.g.mbh);
    k3_batch<<<1, 1024, 0, st>>>(b.frame_bytes, b.frame_off, b.F, b.n0, b.g.P);
}

// ------------------------------------------------------------------------------------------------
// K4: GOP / picture / slice headers (RTL:2645-2656, 2666-2682, 2704-2710).  One thread per slice.
// ------------------------------------------------------------------------------------------------
__global__ void k4_headers(uint32_t *out, const unsigned long long *frame_off, const uint32_t *slice_off,
                           long F, long n0, int P, int mbh, int Q, unsigned long long cap_words) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * mbh) return;
    if (M2V_BODY_WORDS(frame_off[F]) > cap_words) return;
    const long f = i / mbh; const int sl = (int)(i % mbh);
    const long n = n0 + f; const int k = (int)(n % (P + 1));
    unsigned long long pos = 8ull * frame_off[f];
    if (sl == 0) {
        if (k == 0) {                                               // GOP header, time code = absolute frame index
            long hh = n / 86400; if (hh > 63) hh = 63;
            put_bits(out, pos, 0x000001, 24); put_bits(out, pos + 24, 0xB8, 8);
            put_bits(out, pos + 32, (uint32_t)hh, 6); put_bits(out, pos + 38, (uint32_t)((n / 1440) % 60), 6);
            put_bits(out, pos + 44, 0x40u | (uint32_t)((n / 24) % 60), 7); put_bits(out, pos + 51, (uint32_t)(n % 24), 6);
            put_bits(out, pos + 57, 2, 2);
            pos += 64;
        }
        put_bits(out, pos, 0x000001, 24); put_bits(out, pos + 24, (uint32_t)k, 18);
        if (k == 0) { put_bits(out, pos + 42, 0x10000, 19); pos += 64; }
        else { put_bits(out, pos + 42, 0x20000, 19); put_bits(out, pos + 61, 0x380, 11); pos += 72; }
        put_bits(out, pos, 0x000001, 24); put_bits(out, pos + 24, 0xB58111, 24); put_bits(out, pos + 48, 0x1BC000, 24);
    }
    pos = 8ull * (frame_off[f] + (k == 0 ? 25 : 18) + slice_off[i]);
    put_bits(out, pos, 0x000001, 24); put_bits(out, pos + 24, (uint32_t)(sl + 1), 8); put_bits(out, pos + 32, (uint32_t)(2 << Q), 6);
}

void m2v_launch_headers(const M2VBatch &b, cudaStream_t st) {
    const long n = b.F * b.g.mbh;
    k4_headers<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b.out_words, b.frame_off, b.slice_off, b.F, b.n0, b.g.P, b.g.mbh, b.g.Q, b.out_cap_words);
}

// Zeroes exactly the words the body will occupy (the write pass ORs into them), sized on the DEVICE from the scan's total: no
// host round trip between the scans and the write pass.  A body that does not fit the buffer is left alone (and K4 / K2 write
// skip it): the host sees total > capacity after the batch, grows the buffer and runs the batch again.
__global__ void __launch_bounds__(256) k_zero_body(uint4 *out, const unsigned long long *total, unsigned long long cap_words) {
    const unsigned long long words = M2V_BODY_WORDS(*total);
    if (words > cap_words) return;
    const unsigned long long n4 = (words + 3) / 4;             // the buffer is allocated in multiples of 16 bytes
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = make_uint4(0, 0, 0, 0);
}
void m2v_launch_zero_body(const M2VBatch &b, cudaStream_t st) {
    k_zero_body<<<592, 256, 0, st>>>((uint4 *)b.out_words, b.frame_off + b.F, b.out_cap_words);
}
