// m2v_kernels.cuh - launch interface between the host state machine (m2v_host.cu) and the
// sm_100a kernels (m2v_kernels.cu).  Product code; nothing here touches oracle/.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// per-macroblock record written by K1, read by K2/K3:
//   bit 0 inter | bits 8..15 mvx (s8, half-pel) | bits 16..23 mvy | bits 24..29 cbp (Y00 = bit 29)
#define M2V_INFO(inter, mvx, mvy, cbp) \
    ((uint32_t)((inter) & 1) | ((uint32_t)((mvx) & 0xFF) << 8) | ((uint32_t)((mvy) & 0xFF) << 16) | ((uint32_t)((cbp) & 63) << 24))

struct M2VGeom {
    int mbw, mbh, W, H, nmb;   // macroblocks / pixels per frame
    int P;                     // pframes_count; GOP = P+1 frames
    int VL, Q;                 // VECTOR_LEVEL, Q_LEVEL
};

struct M2VBatch {              // one call of encode_gops: frames [n0, n0+F)
    M2VGeom g;
    long F;                    // frames in the batch
    long n0;                   // absolute index of the first frame (GOP aligned)
    const uint8_t *in;         // [F][3][H][W] planar yuv444p (device)
    uint8_t *recon[2];         // ping-pong reconstruction, each [G][fsz420]: Y W*H, then U and V with row stride CWp
    int CWp; size_t fsz420;    // CWp = W/2 rounded up to 16 (TMA needs 16-byte strides); fsz420 = W*H + 2*CWp*H/2
    CUtensorMap tm_in, tm_refY[2], tm_refC[2];   // TMA descriptors (m2v_make_tmaps)
    int16_t *coefs;            // [F][nmb][6][64] quantised levels, zig-zag order
    uint32_t *mbinfo;          // [F][nmb]
    uint32_t *mb_bits;         // [F][nmb]  bit length of each macroblock's syntax
    uint32_t *mb_code;         // [ceil(F*nmb/32)][M2V_MB_SLOT][32]  the macroblocks' own bitstrings, cached by the count pass for the write pass
    uint32_t *mb_off;          // [F][nmb]  bit offset inside its slice (slice header included)
    uint32_t *slice_off;       // [F][mbh]  byte offset of slice inside the frame's slice area
    uint32_t *frame_bytes;     // [F]       slice-area bytes of the frame
    unsigned long long *frame_off; // [F+1] byte offset of each frame in the body; [F] = total
    uint32_t *out_words;       // body, big-endian bit order packed into bytes
    size_t out_cap_words;      // capacity of out_words (32-bit words, a multiple of 4); a larger body is not written (see k_zero_body)
    int k1_grid_cap_i, k1_grid_cap_p;   // persistent K1 grid sizes on the device this batch runs on (m2v_k1_setup)
    unsigned *k1_ctr;          // [2] work counters of K1's dynamic macroblock distribution; launch `seq` uses [seq&1] and
                               // zeroes the other one for the launch after it (both zero before the first launch)
};
#define M2V_BODY_WORDS(total_bytes) ((total_bytes) / 4 + 4)   // words the body needs: its bytes, rounded up, plus slack for the last RED.OR
#define M2V_MB_SLOT 32                // words of a macroblock's bitstring the count pass caches (1024 bits; longer ones are walked twice)
#define M2V_K1_MAX_MBS (1l << 25)   // macroblocks per K1 launch: bound of the division-free index decode

bool m2v_make_tmaps(M2VBatch &b);
// K1: one warp per macroblock; step t = frame index inside every GOP of the batch; seq = running number of the K1
// launches on this stream (selects the work counter)
void m2v_launch_k1(const M2VBatch &b, int t, long ngops_t, unsigned seq, cudaStream_t st);
// K2: one thread per macroblock; count = bit lengths only, write = emit into out_words
void m2v_launch_k2(const M2VBatch &b, bool write, cudaStream_t st);
// K3: slice/frame/batch scans of the bit lengths; then headers
void m2v_launch_k3_scan(const M2VBatch &b, cudaStream_t st);
void m2v_launch_headers(const M2VBatch &b, cudaStream_t st);
// zero the body's words on the device, sized from frame_off[F] (no host round trip)
void m2v_launch_zero_body(const M2VBatch &b, cudaStream_t st);
// per-device K1 configuration (shared-memory attribute, persistent grid sizes); the device must be current
cudaError_t m2v_k1_setup(int VL, int *grid_cap_i, int *grid_cap_p);
// one-time upload of the VLC / quantiser tables for this Q_LEVEL
cudaError_t m2v_upload_tables(int Q);
