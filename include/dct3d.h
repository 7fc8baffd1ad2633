/*
 * dct3d.h -- C ABI of libdct3d.so: the 3D-DCT video codec hot path on NVIDIA B200
 * (sm_100a).  Plain pointers and sizes only; no C++ or torch types.
 *
 * The reference (julianopiccoli/3dDCTVideoEncoding) has no FFI of its own; each entry point
 * below names the reference code it stands in for.  Paths are relative to the reference
 * root:  J/ = 3d-DCT-video-encoding/src/br/jpiccoli/video/ ,  C/ = 3d-DCT-video-encoding-OpenCL/ .
 *
 * Conventions
 *   - every function returns DCT3D_OK (0) or a negative DCT3D_E_* code; a description of the
 *     last failure is available from dct3d_last_error();
 *   - the caller owns every buffer it passes; a context is not thread-safe (one per thread);
 *   - raw video = frames x height x width unsigned bytes, frame-major, no header
 *     (J/Encoder.java:47-56, C/encoder.c:21-35); width and height must be multiples of the
 *     cube edge, trailing frames that do not fill a cube are ignored (J/Encoder.java:39-40);
 *   - the Exp-Golomb stream is MSB-first and bit-continuous across cubes and slabs; its length
 *     in bytes is floor(bits/8)+1 (J/Encoder.java:117, C/encoder.c:270).  zlib wrapping stays
 *     with the caller, as in the reference (J/Encoder.java:114-125, C/encoder.c:266-274);
 *   - cube-major order is [slab][block row][block col][k0=frame][k1=row][k2=col]
 *     (J/Encoder.java:75-89, C/encoder.c:29-41);
 *   - there is no CPU fallback: without a CUDA device every call fails with DCT3D_E_CUDA.
 */
#ifndef DCT3D_H_
#define DCT3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCT3D_OK 0
#define DCT3D_E_INVALID (-1)   /* bad argument (dimensions, cube size, null pointer) */
#define DCT3D_E_CUDA (-2)      /* CUDA runtime / driver error, or no device */
#define DCT3D_E_OVERFLOW (-3)  /* output buffer too small for the stream */
#define DCT3D_E_STREAM (-4)    /* malformed or truncated Exp-Golomb stream */
#define DCT3D_E_NEED_MORE (-5) /* streaming decode: not enough input buffered yet */

typedef struct dct3d_ctx dct3d_ctx;

/* ---- devices and contexts ------------------------------------------------------------- */

/* Number of CUDA devices (negative on error). */
int dct3d_device_count(void);

/* Writes a newline-separated list "index - name (SMs, memory)" into buf.
 * Replaces `codec list_platforms` / printAvailablePlatforms (C/OpenCLUtils.h:13, C/main.c:17-18). */
int dct3d_list_devices(char *buf, size_t cap);

/* Creates a context bound to one GPU for frames of width x height and cubes of edge `cube`
 * (8 or 4; DCT_BLOCK_WIDTH/HEIGHT/DEPTH in C/codec.h:11-13, cubeWidth/Height/Depth in
 * J/Encoder.java:28-30).  Replaces getDeviceId + clCreateContext + buildKernel + clCreateBuffer
 * + clCreateCommandQueue + clCreateKernel (C/encoder.c:148-197, C/decoder.c:153-202). */
int dct3d_create(dct3d_ctx **out, int device, int width, int height, int cube);
void dct3d_destroy(dct3d_ctx *ctx);

/* Last error text of the context (or of the last failed dct3d_create if ctx is NULL). */
const char *dct3d_last_error(const dct3d_ctx *ctx);

/* Page-locked host memory (cudaHostAlloc): buffers passed to the host-buffer entry points move at PCIe
 * speed (about 10x pageable memory) when they come from here.  Replaces the reference's malloc'd
 * inputData/outputData (C/encoder.c:131-134).  dct3d_host_alloc returns NULL on failure. */
void *dct3d_host_alloc(size_t bytes);
void dct3d_host_free(void *p);

/* Options: "tma" (1 = TMA tile loads [default when width % 16 == 0], 0 = plain vector loads);
 * "reuse_zeroed" (default 0): the device-resident encoders zero-fill the stream buffer before packing;
 * with 1, a buffer the context packed into on its previous call (same pointer, same capacity, end bit
 * read back) is only wiped up to where that call wrote -- the caller promises not to have written
 * beyond it in between.
 * "debug" (1 = print per-stage diagnostics to stderr).
 * "chunk_frames" (default 0 = about 32 MB of pixels): frames per pipeline chunk of the host-buffer calls.
 * "piece_bytes" (default 0 = 32 MiB): stream bytes per upload piece of the host-buffer decoders, which parse a stream piece
 * by piece while the later pieces are still being copied to the GPU.
 * "kernel_times_reset": restarts the ring of kernel-time events behind the ns_*_kernel_avg statistics.
 * Kernel variants, all bit-identical in their results (defaults can also be set for a whole process through the
 * environment: DCT3D_ZERO_SKIP, DCT3D_TMA_STORE, DCT3D_PACK_SORT): "zero_skip" (default 1; 8x8x8 fused encoder: groups
 * of high diagonals are tested for zero before they are quantised), "tma_store" (default 1: the inverse kernel stores
 * its pixel tiles by TMA when width % 32 == 0 and the frame buffer is 16-byte aligned; statistic "tma_store_used"
 * tells whether the last launch did), "pack_sort" (0 = the bit packer takes the cubes of a tile in stream order,
 * 1 = dealt to its threads by chunk count, 2 = the same with balanced warps).
 * "precision" (32 [default] or 64): with 64 the fused and stage entry points (encode_u8, decode_u8, quantize_u8,
 * reconstruct_i16 and the streaming calls) compute in double like the Java reference (J/dct/DCT.java:41-59,
 * J/Encoder.java:82, J/Decoder.java:89,112): quantised cubes then equal the fp64 oracle without rounding-tie
 * flips, at a fraction of the fp32 path's speed.  "rounding" (fp64 mode only): 0 = Math.round = floor(v + 0.5)
 * (J/Encoder.java:82), 1 = C round(), ties away from zero (C/encoder.c:53).
 * Statistics (dct3d_get_stat): "launches" = kernels launched by the context so far; "tma" = 1 when the
 * TMA path is active; "num_sms"; "chunks" = pipeline chunks of the last host-buffer call; "ns_encode_kernel" /
 * "ns_reconstruct_kernel" = duration in ns of the last transform kernel of each direction, from CUDA events recorded around
 * it on the launching stream; "ns_encode_kernel_avg" / "ns_reconstruct_kernel_avg" = the mean over the launches since option
 * "kernel_times_reset" was set (a ring of the last 32), read without having synchronised in between. */
int dct3d_set_option(dct3d_ctx *ctx, const char *key, long value);
long dct3d_get_stat(const dct3d_ctx *ctx, const char *key);

/* ---- fused hot path, host buffers ------------------------------------------------------- */

/* u8 frames -> Exp-Golomb stream: DCT + quantise + zig-zag + Exp-Golomb in one pass on the GPU.
 * Replaces J/Encoder.java:51-111 and, per slab, readCubes + the cl* sequence + applyQuantization
 * + applyExpGolombCoding (C/encoder.c:206-263).  On success *nbits = stream length in bits and
 * *nbytes = floor(nbits/8)+1 bytes were written to `stream` (unused trailing bits zero). */
int dct3d_encode_u8(dct3d_ctx *ctx, const uint8_t *frames, int nframes,
                    uint8_t *stream, size_t cap, uint64_t *nbits, size_t *nbytes);

/* Exp-Golomb stream -> u8 frames: parse + dequantise + inverse DCT + clamp + truncate.
 * Replaces J/Decoder.java:61-117 and, per slab, C/decoder.c:229-295. */
int dct3d_decode_u8(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, int nframes, uint8_t *frames);

/* Both calls above move the clip as a pipeline of slab-range chunks (option "chunk_frames", default about 32 MB of
 * pixels): the H2D copy of one chunk runs beside the kernels of the previous one and beside the D2H copy of what is
 * finished, and the chunks of one call form ONE stream (the bit position is chained on the device), exactly as the
 * reference's slab loop carries its partial byte (C/encoder.c:203-278, C/ExpGolomb.c:112-130). */

/* ---- sharded coding: one slab range per GPU (or per process), one stream ------------------------
 * Every slab of `cube` frames is an independent key-frame group (reference README.md:10); only the bit position
 * couples them.  A range is coded from bit 0 on its own GPU (phase 1), the bit counts of all ranges are prefix-summed
 * by the caller (G scalars, any channel), and every range is then moved to its place in the clip's one stream
 * (phase 2).  This is the concatenation rule of expGolomb_freeBuffer (C/ExpGolomb.c:112-130) with C/encoder.c:263-271,
 * applied across GPUs instead of across loop iterations. */

/* Phase 1: codes `nframes` host frames into the context's device buffer from bit 0; *nbits = the range's bit count. */
int dct3d_encode_u8_range(dct3d_ctx *ctx, const uint8_t *frames, int nframes, uint64_t *nbits);

/* Phase 2: moves the range coded by the last dct3d_encode_u8_range to global bit `start_bit` of the host buffer `stream`
 * (capacity `cap` bytes).  The bits are shifted to phase start_bit % 8 on the GPU and copied straight to byte
 * start_bit / 8 onwards.  When start_bit % 8 != 0 the range's first byte is shared with its predecessor: it is NOT
 * written but returned in *first_byte, to be OR-ed into stream[start_bit / 8] by the caller once the predecessor's bytes
 * have landed.  The byte after the range's last bit is written only when the range has bits in it or `last` != 0 (the
 * stream's closing byte: floor(bits/8)+1 bytes in all, J/Encoder.java:117, C/encoder.c:270). */
int dct3d_encode_u8_place(dct3d_ctx *ctx, uint64_t start_bit, int last, uint8_t *stream, size_t cap, uint8_t *first_byte);

/* Decodes `nframes` frames whose first code starts at bit `start_bit` of the host buffer `stream` (any bit, no
 * alignment).  end_bit_hint (0 = unknown): the bit at which the range ends when the caller knows it; it only limits how
 * much of the stream is copied to the GPU.  *end_bit (may be NULL) = first bit after the range. */
int dct3d_decode_u8_range(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit, uint64_t end_bit_hint,
                          int nframes, uint8_t *frames, uint64_t *end_bit);

/* Page-locks an existing host buffer (e.g. a shared-memory mapping that several processes place their ranges into). */
int dct3d_host_register(void *p, size_t bytes);
int dct3d_host_unregister(void *p);

/* Several GPUs in one process.  `devices` = CUDA ordinals (NULL = 0..ndevices-1).  GPU g codes slabs
 * [g*n/G, (g+1)*n/G) on its own host thread; dct3d_multi_encode_u8 returns the ONE stream the reference would have
 * written for the whole clip (bit-identical to dct3d_encode_u8 on a single GPU) and, in range_start_bits (G+1 entries,
 * may be NULL), the bit at which every range starts.  dct3d_multi_decode_u8 takes those as side information; with
 * range_start_bits == NULL (a file from the reference itself: the format stores no index, J/ExpGolombReader.java:19-63)
 * the start bits are found by index discovery first (dct3d_multi_locate).  Replaces the single-device flow of
 * C/encoder.c:148-278 / C/decoder.c:153-295; there is no collective on the data path. */
typedef struct dct3d_multi dct3d_multi;
int dct3d_multi_create(dct3d_multi **out, const int *devices, int ndevices, int width, int height, int cube);
void dct3d_multi_destroy(dct3d_multi *m);
const char *dct3d_multi_last_error(const dct3d_multi *m);
int dct3d_multi_set_option(dct3d_multi *m, const char *key, long value);
dct3d_ctx *dct3d_multi_context(dct3d_multi *m, int index);   /* the per-GPU context (statistics, options) */
/* Shares of the slabs: GPU g codes a contiguous range of about weights[g] / sum(weights) of them (NULL = equal shares, the
 * default).  The stream does not depend on the shares.  dct3d_multi_probe_links measures the host<->device copy rate of every
 * GPU while all of them copy at once (GB/s each way, arrays of ndevices entries, any may be NULL) and proposes weights
 * 1 / (1/h2d + 1/d2h): on hosts whose GPUs do not get the same share of the host links (this pool: 24/12 GB/s for four of
 * eight GPUs, 39/20 for the others) the end-to-end calls are bound by the slowest link unless the ranges follow them. */
int dct3d_multi_set_weights(dct3d_multi *m, const double *weights);
int dct3d_multi_probe_links(dct3d_multi *m, double *h2d_gbs, double *d2h_gbs, double *weights);
int dct3d_multi_encode_u8(dct3d_multi *m, const uint8_t *frames, int nframes, uint8_t *stream, size_t cap,
                          uint64_t *nbits, size_t *nbytes, uint64_t *range_start_bits);
int dct3d_multi_locate(dct3d_multi *m, const uint8_t *stream, size_t nbytes, int nframes, uint64_t *range_start_bits);
int dct3d_multi_decode_u8(dct3d_multi *m, const uint8_t *stream, size_t nbytes, int nframes, uint8_t *frames,
                          const uint64_t *range_start_bits);

/* The streaming calls below (dct3d_stream_*) over several GPUs: a batch of slabs per call, shared out as slab ranges, the
 * partial byte carried between calls.  dct3d_multi_stream_decode finds the ranges' start bits by distributed index
 * discovery on the buffered bytes (every GPU counts the codes of 1/G of the bytes) and returns
 * DCT3D_E_NEED_MORE, changing nothing, while they do not hold the whole batch. */
int dct3d_multi_stream_begin(dct3d_multi *m);
int dct3d_multi_stream_encode(dct3d_multi *m, const uint8_t *frames, int nframes, int last, uint8_t *out, size_t cap, size_t *nbytes);
int dct3d_multi_stream_decode(dct3d_multi *m, const uint8_t *in, size_t nbytes, uint64_t *bitpos, int nframes, uint8_t *frames);

/* ---- streaming (the C codec's slab loop with a carried bit position) -------------------- */

/* Resets the carried bit state (expGolomb_createStream, C/ExpGolomb.c:24-30). */
int dct3d_stream_begin(dct3d_ctx *ctx);

/* Encodes `nframes` more frames (a multiple of the cube edge).  Writes the COMPLETE bytes
 * produced so far to `out` and keeps the partial byte inside the context, exactly as
 * applyExpGolombCoding + expGolomb_freeBuffer do (C/encoder.c:263-268, C/ExpGolomb.c:112-122).
 * With last != 0 the partial byte is flushed too (C/encoder.c:269-271: size+1 bytes). */
int dct3d_stream_encode(dct3d_ctx *ctx, const uint8_t *frames, int nframes, int last,
                        uint8_t *out, size_t cap, size_t *nbytes);

/* Decodes `nframes` frames from `in`, starting at bit *bitpos of `in` (0..7 after the caller
 * dropped consumed bytes, as expGolomb_freeBuffer(..., 0) does, C/decoder.c:229-235).  If the
 * buffered input does not hold all codes yet returns DCT3D_E_NEED_MORE and changes nothing;
 * otherwise advances *bitpos to the first unread bit. */
int dct3d_stream_decode(dct3d_ctx *ctx, const uint8_t *in, size_t nbytes, uint64_t *bitpos,
                        int nframes, uint8_t *frames);

/* ---- transform seams, host buffers ------------------------------------------------------ */

/* float cube-major slabs in, float cube-major coefficients out: the reference's own device
 * boundary, replacing clEnqueueWriteBuffer + dct_calculate_partial_sums +
 * dct_aggregate_partial_sums + clEnqueueReadBuffer (C/encoder.c:209-254, C/3dDCT.cl:43-143). */
int dct3d_forward_f32(dct3d_ctx *ctx, const float *cubes_in, float *coef_out, int nslabs);
/* inverse + clamp to [0,255] (C/decoder.c:249-292, C/3dDCT.cl:164-265). */
int dct3d_inverse_f32(dct3d_ctx *ctx, const float *coef_in, float *pixels_out, int nslabs);

/* double planar [frames][height][width] in and out: `new DCT(in, out, w, h, c, c, c).run()`
 * (J/Encoder.java:63-64, J/dct/DCT.java:41-59) and `new InverseDCT(...).run()`
 * (J/Decoder.java:102-103, J/dct/InverseDCT.java:33-82; clamps to [0,255]).  fp64 arithmetic. */
int dct3d_forward_f64(dct3d_ctx *ctx, const double *planar_in, double *planar_out, int nframes);
int dct3d_inverse_f64(dct3d_ctx *ctx, const double *planar_in, double *planar_out, int nframes);

/* ---- codec stages, host buffers (flow-preserving mode and parity tests) ------------------ */

/* u8 frames -> quantised cube-major int16 cubes, natural order inside the cube.
 * DCT + round(coef / max(1, 5*(k0+k1+k2))) (J/Encoder.java:69-89, C/encoder.c:47-58). */
int dct3d_quantize_u8(dct3d_ctx *ctx, const uint8_t *frames, int nframes, int16_t *qcubes);
/* quantised cubes -> u8 frames (J/Decoder.java:78-117, C/decoder.c:48-59 + inverse + :29). */
int dct3d_reconstruct_i16(dct3d_ctx *ctx, const int16_t *qcubes, int nframes, uint8_t *frames);
/* quantised cubes -> Exp-Golomb stream in zig-zag order, starting at bit `start_bit` of
 * `stream` (which must be zero from that bit on).  Bit-exact with applyExpGolombCoding
 * (C/encoder.c:60-71) / J/Encoder.java:101-111 on identical cubes. */
int dct3d_eg_encode_i16(dct3d_ctx *ctx, const int16_t *qcubes, size_t ncubes, uint64_t start_bit,
                        uint8_t *stream, size_t cap, uint64_t *end_bit);
/* stream -> quantised cubes (J/Decoder.java:61-76, C/decoder.c:229-239). */
int dct3d_eg_decode_i16(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit,
                        size_t ncubes, int16_t *qcubes, uint64_t *end_bit);

/* Finds where the first `ncubes` cubes of the stream end (*end_bit) without decoding them: index discovery only.
 * The stream has no index (J/ExpGolombReader.java:19-63 is the only way the reference knows a code boundary), so
 * this is how a GPU that decodes a later slab range learns its start bit when the encoder's bit counts are not at
 * hand. */
int dct3d_eg_locate(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit, size_t ncubes, uint64_t *end_bit);

/* ---- device-resident variants ------------------------------------------------------------
 * All pointers are device pointers on the context's GPU; `cuda_stream` is a cudaStream_t (NULL =
 * the context's own stream, which is non-blocking: it does not synchronise with the legacy default
 * stream, so work the caller has pending on the buffers must be complete, or the caller passes the
 * stream that work was enqueued on).  Work is enqueued on that stream; scalar results are written to
 * host memory after an internal stream synchronisation unless the pointer is NULL.
 * d_stream must be 4-byte aligned, zero-filled from start_bit on, and `cap` must include 8
 * bytes of slack.  d_frames must be aligned to the cube edge (8 or 4 bytes; 16 for the TMA path, else plain loads are
 * used), d_qcubes and the f32/f64 buffers to 16 bytes; a misaligned pointer is rejected with DCT3D_E_INVALID. */
int dct3d_encode_u8_dev(dct3d_ctx *ctx, const void *d_frames, int nframes, void *d_stream, size_t cap,
                        uint64_t start_bit, uint64_t *end_bit, void *cuda_stream);
/* The cube boundaries are discovered from the stream itself (it stores no index); start_bit is any bit of d_stream. */
int dct3d_decode_u8_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit,
                        int nframes, void *d_frames, uint64_t *end_bit, void *cuda_stream);
int dct3d_forward_f32_dev(dct3d_ctx *ctx, const void *d_cubes_in, void *d_coef_out, int nslabs, void *cuda_stream);
int dct3d_inverse_f32_dev(dct3d_ctx *ctx, const void *d_coef_in, void *d_pixels_out, int nslabs, void *cuda_stream);
int dct3d_forward_f64_dev(dct3d_ctx *ctx, const void *d_planar_in, void *d_planar_out, int nframes, void *cuda_stream);
int dct3d_inverse_f64_dev(dct3d_ctx *ctx, const void *d_planar_in, void *d_planar_out, int nframes, void *cuda_stream);
int dct3d_quantize_u8_dev(dct3d_ctx *ctx, const void *d_frames, int nframes, void *d_qcubes, void *cuda_stream);
int dct3d_reconstruct_i16_dev(dct3d_ctx *ctx, const void *d_qcubes, int nframes, void *d_frames, void *cuda_stream);
int dct3d_eg_encode_i16_dev(dct3d_ctx *ctx, const void *d_qcubes, size_t ncubes, uint64_t start_bit,
                            void *d_stream, size_t cap, uint64_t *end_bit, void *cuda_stream);
int dct3d_eg_decode_i16_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit,
                            size_t ncubes, void *d_qcubes, uint64_t *end_bit, void *cuda_stream);
int dct3d_eg_locate_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit, size_t ncubes,
                        uint64_t *end_bit, void *cuda_stream);
/* d_dst bit (phase + i) = d_src bit i for i < nbits, the first `phase` (0..7) bits of d_dst zero: a range coded from bit 0
 * moved to its phase in the clip's stream (device-resident form of dct3d_encode_u8_place).  d_src must be zero beyond
 * its last bit; cap >= 4 * ((nbits + phase + 31) / 32 + 1). */
int dct3d_stream_shift_dev(dct3d_ctx *ctx, const void *d_src, uint64_t nbits, unsigned phase, void *d_dst, size_t cap, void *cuda_stream);

/* ---- colour planes (the reference codes colour video as three gray streams) --------------------
 * RGBUtils split / mix (J/RGBUtils.java:39-92 and :94-131): byte i of a raw RGB24 buffer belongs to
 * plane i % 3.  split: nbytes of RGB -> planes of (nbytes+2)/3, (nbytes+1)/3 and nbytes/3 bytes;
 * mix: three planes of npixels bytes -> 3*npixels bytes of RGB.  The _dev variants use the 16-byte
 * vector path when all four pointers are 16-byte aligned. */
int dct3d_rgb_split(dct3d_ctx *ctx, const uint8_t *rgb, size_t nbytes, uint8_t *r, uint8_t *g, uint8_t *b);
int dct3d_rgb_mix(dct3d_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, size_t npixels, uint8_t *rgb);
int dct3d_rgb_split_dev(dct3d_ctx *ctx, const void *d_rgb, size_t nbytes, void *d_r, void *d_g, void *d_b, void *cuda_stream);
int dct3d_rgb_mix_dev(dct3d_ctx *ctx, const void *d_r, const void *d_g, const void *d_b, size_t npixels, void *d_rgb, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DCT3D_H_ */
