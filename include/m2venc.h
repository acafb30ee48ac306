/* include/m2venc.h - C-ABI of the B200-native MPEG-2 I/P encoder (libm2venc.so).
 *
 * Drop-in boundary for the reference's streaming contract: the module port list of
 * RTL/mpeg2encoder.v:10-38 (semantics README.md:90-234), as driven by SIM/tb_mpeg2encoder.v:95-128
 * (instance) and :206-266 (stimulus / sink).  Plain C, plain pointers and sizes; no C++/torch types.
 * Every function returns M2V_OK (0) or a negative M2V_E* code; nothing throws across the ABI.
 * A handle is NOT thread-safe (the RTL is one clock domain, one sequence at a time): one host thread per
 * handle.  Inside, the library owns the CUDA streams, the pinned staging / output memory and one worker
 * thread per device, for 1..8 devices (m2v_create: the device current at the call; m2v_create_multi:
 * devices 0..ndev-1).  Several handles, on the same or on different devices, can live in one process.
 *
 * There is no CPU fallback - m2v_create() fails with M2V_ENODEV when no sm_100 device is usable.
 */
#ifndef M2VENC_H
#define M2VENC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct m2v_encoder m2v_encoder;

enum {
    M2V_OK = 0,
    M2V_EINVAL = -1,   /* bad argument (parameter outside the sets the RTL documents) */
    M2V_ESTATE = -2,   /* call not legal in this state (e.g. begin while busy) */
    M2V_ENODEV = -3,   /* no usable CUDA device / kernel image */
    M2V_ENOMEM = -4,
    M2V_ECUDA  = -5,   /* CUDA runtime error; see m2v_last_error() */
    M2V_ESPACE = -6    /* caller buffer too small */
};

/* rstn + static parameters (RTL:11-14; README.md:79-84): XL,YL in 4..7, VECTOR_LEVEL in 1..3,
 * Q_LEVEL in 1..4.  Equivalent to instantiating the module and pulsing reset (README.md:96). */
int  m2v_create(int XL, int YL, int VECTOR_LEVEL, int Q_LEVEL, m2v_encoder **out);
/* The same module instance spread over `ndev` GPUs (devices 0..ndev-1, 1 <= ndev <= 8) of this process: the streaming
 * calls below deal whole closed GOPs (RTL:2645-2656, 1820-1825) to the devices in batches and m2v_pull / m2v_drain
 * return ONE ordered word stream, byte-identical to the single-device one (RTL:10-38: one instance, one stream). */
int  m2v_create_multi(int ndev, int XL, int YL, int VECTOR_LEVEL, int Q_LEVEL, m2v_encoder **out);
int  m2v_device_count(const m2v_encoder *e);
void m2v_destroy(m2v_encoder *e);
const char *m2v_last_error(const m2v_encoder *e);

/* i_xsize16 / i_ysize16 / i_pframes_count, sampled by the RTL on the first i_en of a sequence
 * (RTL:1060-1065).  Sizes are clamped exactly like RTL:985-991 (never an error); the clamped size
 * in macroblocks is returned through mbw/mbh when non-NULL.  M2V_ESTATE while o_sequence_busy. */
int  m2v_begin(m2v_encoder *e, int xsize16, int ysize16, int pframes_count, int *mbw, int *mbh);

/* One i_en cycle: 4 horizontally adjacent YUV 4:4:4 pixels in raster order (RTL:25-28). */
int  m2v_push4(m2v_encoder *e, const uint8_t Y[4], const uint8_t U[4], const uint8_t V[4]);

/* Bulk form of the same stimulus: nframes planar yuv444p frames (Y plane, U plane, V plane per
 * frame, exactly as TB:210-218 loads them) of the CLAMPED geometry, in HOST memory.  Must start on
 * a frame boundary.  The library has copied what it needs when the call returns (the buffer can be
 * reused at once); the kernels of the last batches may still be running then and overlap the caller's
 * next read.  Memory from m2v_alloc_host (or registered with m2v_register_host) is copied at PCIe
 * rate; pageable memory works but goes through the driver's bounce buffer.  nframes == 0 is a no-op
 * (no i_en, the sequence is not armed: RTL:1060-1065). */
int  m2v_push_frames(m2v_encoder *e, const uint8_t *yuv444p, long nframes);

/* i_sequence_stop (RTL:1082-1083,1090-1091): an unfinished frame is padded with Y=0,U=V=0x80
 * (RTL:1036-1037,1049-1056); the end code and the final padded word are produced.  Returns when every
 * batch has been encoded, i.e. the whole stream can be pulled. */
int  m2v_stop(m2v_encoder *e);

/* o_sequence_busy (RTL:1095): 1 from the first pixel until the o_last word has been pulled. */
int  m2v_busy(const m2v_encoder *e);

/* o_en/o_data/o_last (RTL:2961-2994): returns 1 and fills one 32-byte word (stream order, first
 * stream byte in out[0]), 0 if no word is available yet, <0 on error.  The GPU path emits words
 * later than the RTL (whole batches of GOPs) but in identical order and content. */
int  m2v_pull(m2v_encoder *e, uint8_t out[32], int *last);

/* Bulk pull: copies as many whole available words as fit in cap; *n = bytes written;
 * *last = 1 when the o_last word was among them.  Never blocks: words of batches still on a device
 * are simply not available yet (all of them are after m2v_stop). */
int  m2v_drain(m2v_encoder *e, uint8_t *dst, size_t cap, size_t *n, int *last);

/* ---- device-resident bulk path (used for GOP sharding across GPUs and by bench.py) -----------
 * Encodes frames [n0, n0+nframes) of a sequence whose frames are ALREADY in device memory
 * (planar yuv444p, clamped geometry given by mbw/mbh).  n0 must be the absolute index of an
 * I-frame (n0 % (pframes_count+1) == 0) - closed GOPs are independent (RTL:2645-2656,1820-1825).
 * Produces the byte-aligned GOP/picture/slice layers of those frames ("body"), i.e. everything
 * between the 34-byte sequence header and the sequence end code.
 *   d_body/body_len : internal DEVICE buffer holding the body, valid until the next call
 * Runs on the handle's first device.  Preconditions: d_yuv444p is 16-byte aligned (the frames are read
 * by TMA; M2V_EINVAL otherwise); the caller's writes to the frames have completed (the kernels run on
 * the library's own non-blocking stream, which does not order against the caller's streams: synchronise
 * first); no streamed batch is in flight (M2V_ESTATE).  Does not touch the streaming state. */
int  m2v_encode_gops_device(m2v_encoder *e, int mbw, int mbh, int pframes_count,
                            const uint8_t *d_yuv444p, long nframes, long n0,
                            const uint8_t **d_body, size_t *body_len);

/* Same, then copies the body to host memory (h_body may be pageable or pinned). */
int  m2v_encode_gops_host(m2v_encoder *e, int mbw, int mbh, int pframes_count,
                          const uint8_t *d_yuv444p, long nframes, long n0,
                          uint8_t *h_body, size_t cap, size_t *body_len);

/* Asynchronous form of m2v_encode_gops_device for callers that shard GOPs over processes and want the body of chunk i
 * on its way to the host while chunk i+1 is encoded.  Two slots (0/1), each with its own body buffer:
 *   m2v_gops_submit : queue the whole hot path of a chunk into `slot`; returns at once
 *   m2v_gops_size   : body length of that chunk - known after the scans, while the write pass is still running
 *   m2v_gops_fetch  : queue the device->host copy of the body (after the write pass) on the copy-out stream
 *   m2v_gops_wait   : wait for the chunk's kernels and fetch;  m2v_gops_body : the slot's device buffer
 * A chunk is at most what one launch sequence takes (m2v_encode_gops_device cuts longer jobs itself). */
int  m2v_gops_submit(m2v_encoder *e, int mbw, int mbh, int pframes_count, const uint8_t *d_yuv444p, long nframes, long n0, int slot);
int  m2v_gops_size(m2v_encoder *e, int slot, size_t *body_len);
int  m2v_gops_fetch(m2v_encoder *e, int slot, uint8_t *h_dst, size_t len);
int  m2v_gops_wait(m2v_encoder *e, int slot);
int  m2v_gops_body(m2v_encoder *e, int slot, const uint8_t **d_body);

/* Pinned host memory for frames and streams (host<->device copies at PCIe rate), and pinning of memory the caller
 * already owns (a shared-memory arena, an mmap). */
void *m2v_alloc_host(size_t bytes);
void  m2v_free_host(void *p);
int  m2v_register_host(void *p, size_t bytes);
int  m2v_unregister_host(void *p);

/* Host-side framing helpers (RTL:2596-2617; RTL:2621-2628 + 2932-2937). */
int  m2v_sequence_header(int mbw, int mbh, uint8_t out34[34]);
/* Given a buffer holding header+bodies in its first `len` bytes, appends 00 00 01 B7 and zero
 * padding up to the next 32-byte multiple with the RTL's "always one more word" rule. */
int  m2v_finish_stream(uint8_t *buf, size_t len, size_t cap, size_t *total);

/* Debug taps used by the parity tests (device -> host copies of the last encode_gops call).  A tile
 * of `coefs` is only meaningful where the macroblock's cbp bit is set: uncoded tiles are not written. */
int  m2v_debug_copy(m2v_encoder *e, uint32_t *mbinfo, int16_t *coefs, long nframes_times_nmb);
/* Kernel launch counter (bench.py "gpu_launches") and device time in ms (CUDA events on the
 * launching stream) of the last encode_gops call (or, for the m2v_gops_* form, accumulated by m2v_gops_wait): idx 0 = all
 * mb_encode (K1) launches, 1 = vlc count, 2 = scans + body zeroing + headers, 3 = vlc write, 4 = first launch -> last kernel end. */
long m2v_launch_count(const m2v_encoder *e);
/* Test knob: overrides the sizes the library picks by itself - the streaming flush threshold (frames per batch of
 * m2v_push*; rounded down to whole GOPs, at least one) and the frames per internal chunk of m2v_encode_gops_* -
 * so that the multi-batch / multi-chunk paths can be driven with small clips.  0 = automatic.  The stream does not
 * depend on either value (closed GOPs, RTL:2645-2656). */
int  m2v_set_limits(m2v_encoder *e, long batch_frames, long chunk_frames);
/* Test knob: bytes of body buffer reserved per macroblock (default 192, several times a typical body).  A batch whose
 * body does not fit is detected after its scans, the buffer grows to what the scan asked for and the batch runs again;
 * a tiny value forces that path. */
int  m2v_set_body_reserve(m2v_encoder *e, long bytes_per_macroblock);
int  m2v_kernel_ms(const m2v_encoder *e, float ms[5]);
int  m2v_set_timing(m2v_encoder *e, int enable);

#ifdef __cplusplus
}
#endif
#endif
