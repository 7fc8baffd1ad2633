/* pdeflate.h -- multi-threaded deflate that still emits ONE valid RFC 1950 zlib stream.
 *
 * Replaces the single-threaded deflate() calls of the reference's container stage
 * (3d-DCT-video-encoding-OpenCL/encoder.c:73-86,266-274; Java: Encoder.java:114-125), which is more
 * than 95% of the encoder's wall time once the transform runs on the GPU.
 * The stream is what `inflate` (decoder.c:213-227, Decoder.java:41-59) and any zlib reader expect;
 * the bytes differ from a single-threaded deflate of the same input (block boundaries), the
 * inflated data does not.
 */
#ifndef PDEFLATE_H_
#define PDEFLATE_H_

#include <stddef.h>
#include <stdio.h>

typedef struct pdeflate pdeflate;

/* level: zlib level 0..9 (the reference uses Z_BEST_COMPRESSION); threads <= 0: all online cores;
 * block: bytes of input per independent deflate block (0 = 256 KiB). */
pdeflate *pdeflate_open(FILE *out, int level, int threads, size_t block);

/* Queues n bytes (copied; the caller may reuse its buffer at once).  Blocks only when too much input is
 * in flight.  Returns 0, or -1 after an I/O or zlib failure. */
int pdeflate_write(pdeflate *p, const unsigned char *data, size_t n);

/* Compresses what is left, writes the final block and the Adler-32 trailer, joins the workers.
 * Returns 0 on success.  *in_bytes / *out_bytes (optional) receive the totals. */
int pdeflate_close(pdeflate *p, unsigned long long *in_bytes, unsigned long long *out_bytes);

#endif /* PDEFLATE_H_ */
