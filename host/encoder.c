/* encoder.c -- the C codec's encode flow over libdct3d.so.
 *
 * Same flow as the reference's 3d-DCT-video-encoding-OpenCL/encoder.c:88-293: read frames, code them, deflate the
 * complete bytes, carry the partial byte into the next round, Z_FINISH after the last one.  What changed:
 *   - what sits between the read and deflate: readCubes' float reshuffle (:10-45), the clEnqueueWriteBuffer / two
 *     kernels / clEnqueueReadBuffer sequence (:209-254), applyQuantization (:47-58) and applyExpGolombCoding (:60-71)
 *     with expGolomb_freeBuffer (ExpGolomb.c:112-122) are ONE call, dct3d_multi_stream_encode, which takes raw u8 frames and
 *     returns the complete stream bytes; OpenCLUtils.c is replaced by dct3d_multi_create (one GPU, or the list in
 *     DCT3D_DEVICES / a comma-separated last argument: the batch is then shared out as slab ranges, one stream);
 *   - the unit of work: a BATCH of slabs per call (DCT3D_BATCH_SLABS, default 16 = 128 frames per GPU) instead of one slab
 *     (:203-206), so that the library's chunk pipeline overlaps the H2D copies with the kernels, and there is one
 *     host round trip per batch instead of one per 8 frames;
 *   - the read (:21-27, a blocking fread per slab): a reader thread fills the other of two page-locked batch buffers
 *     while the current one is coded, with O_DIRECT reads when the batch size allows it (falls back to buffered reads);
 *   - the container stage (deflate, :73-86,266-274) keeps its format -- one zlib stream, level 9 -- but runs on all host
 *     cores (pdeflate.c) while the next batch is read and coded; DCT3D_ZLIB_LEVEL / DCT3D_ZLIB_THREADS override level
 *     and thread count.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include "../include/dct3d.h"
#include "codec.h"
#include "pdeflate.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* Input side of the batch pipeline: two page-locked buffers, filled in turn by the reader thread. */
typedef struct {
    int fd, direct;                    /* direct: the descriptor was opened with O_DIRECT */
    const char *path;
    unsigned char *buf[2];
    size_t want[2], got[2];            /* bytes asked for / delivered (short at the end of the file) */
    int full[2], stop, failed;
    off_t offset;
    pthread_mutex_t mu;
    pthread_cond_t cv;
} batch_reader;

static size_t read_fully(batch_reader *r, unsigned char *dst, size_t n)
{
    size_t got = 0;
    while (got < n) {
        ssize_t k = pread(r->fd, dst + got, n - got, r->offset + (off_t)got);
        if (k < 0 && errno == EINVAL && r->direct) {
            /* O_DIRECT refuses this length or offset (the tail of the file): reopen buffered and go on */
            int fd = open(r->path, O_RDONLY);
            if (fd < 0) break;
            close(r->fd);
            r->fd = fd;
            r->direct = 0;
            continue;
        }
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) break;
        got += (size_t)k;
    }
    r->offset += (off_t)got;
    return got;
}

static void *batch_reader_main(void *arg)
{
    batch_reader *r = (batch_reader *)arg;
    int b = 0;
    for (;;) {
        pthread_mutex_lock(&r->mu);
        while (r->full[b] && !r->stop) pthread_cond_wait(&r->cv, &r->mu);
        if (r->stop) { pthread_mutex_unlock(&r->mu); return NULL; }
        const size_t want = r->want[b];
        pthread_mutex_unlock(&r->mu);
        const size_t got = want ? read_fully(r, r->buf[b], want) : 0;
        pthread_mutex_lock(&r->mu);
        r->got[b] = got;
        r->full[b] = 1;
        pthread_cond_broadcast(&r->cv);
        pthread_mutex_unlock(&r->mu);
        if (want == 0) return NULL;    /* a zero-byte request ends the thread */
        b ^= 1;
    }
}

int encode(char *inputFileName, char *outputFileName, int width, int height, int framesToEncode, int platformIndex)
{
    const double t_start = now_s();
    const size_t slabBytes = (size_t)width * height * DCT_BLOCK_DEPTH;
    const char *bs = getenv("DCT3D_BATCH_SLABS");
    int devices[64];
    const int ndevices = codec_devices(platformIndex, devices, 64);
    int batchSlabs = bs ? atoi(bs) : 16 * ndevices;              /* every GPU gets 16 slabs of a batch */
    if (batchSlabs < 1) batchSlabs = 1;
    const int totalSlabs = (framesToEncode + DCT_BLOCK_DEPTH - 1) / DCT_BLOCK_DEPTH;   /* the reference codes whole slabs (:203) */
    if (batchSlabs > totalSlabs) batchSlabs = totalSlabs > 0 ? totalSlabs : 1;
    const size_t batchBytes = slabBytes * batchSlabs;

    FILE *outputFile = fopen(outputFileName, "wb");
    batch_reader reader;
    memset(&reader, 0, sizeof reader);
    reader.path = inputFileName;
    reader.direct = (batchBytes % 4096 == 0) && !getenv("DCT3D_NO_ODIRECT");
    reader.fd = reader.direct ? open(inputFileName, O_RDONLY | O_DIRECT) : -1;
    if (reader.fd < 0) { reader.direct = 0; reader.fd = open(inputFileName, O_RDONLY); }
    if (reader.fd < 0 || !outputFile) { printf("Error opening files\n"); return 1; }
    /* page-locked staging buffers: the batch goes to the GPU at PCIe speed */
    reader.buf[0] = (unsigned char *)dct3d_host_alloc(batchBytes);
    reader.buf[1] = (unsigned char *)dct3d_host_alloc(batchBytes);
    size_t egCap = batchBytes / 2 + 4096;                        /* 4 bit/sample; grown on DCT3D_E_OVERFLOW (worst case 4 bytes/sample) */
    unsigned char *expGolombBuffer = (unsigned char *)dct3d_host_alloc(egCap);
    if (!reader.buf[0] || !reader.buf[1] || !expGolombBuffer) { printf("Error allocating host buffers\n"); return 1; }

    const char *lv = getenv("DCT3D_ZLIB_LEVEL"), *th = getenv("DCT3D_ZLIB_THREADS");
    pdeflate *zlibStream = pdeflate_open(outputFile, lv ? atoi(lv) : Z_BEST_COMPRESSION, th ? atoi(th) : 0, 0);
    if (!zlibStream) { printf("Error starting deflate\n"); return 1; }

    printf("Getting device id\n");
    dct3d_multi *ctx = NULL;
    if (dct3d_multi_create(&ctx, devices, ndevices, width, height, DCT_BLOCK_WIDTH) != DCT3D_OK) {
        printf("Error creating dct3d context: %s\n", dct3d_last_error(NULL));
        return 1;
    }
    dct3d_multi_stream_begin(ctx);

    pthread_mutex_init(&reader.mu, NULL);
    pthread_cond_init(&reader.cv, NULL);
    /* ask for the first two batches before the thread starts */
    int slabsAsked = 0;
    for (int b = 0; b < 2; b++) {
        const int n = totalSlabs - slabsAsked < batchSlabs ? totalSlabs - slabsAsked : batchSlabs;
        reader.want[b] = slabBytes * (size_t)n;
        slabsAsked += n;
    }
    pthread_t readerThread;
    pthread_create(&readerThread, NULL, batch_reader_main, &reader);

    printf("Starting encoding process\n");
    const double t_loop = now_s();
    int framesRead = 0, slabsDone = 0, cur = 0, rc = 0;
    while (slabsDone < totalSlabs) {
        const int n = totalSlabs - slabsDone < batchSlabs ? totalSlabs - slabsDone : batchSlabs;
        const size_t bytes = slabBytes * (size_t)n;
        pthread_mutex_lock(&reader.mu);
        while (!reader.full[cur]) pthread_cond_wait(&reader.cv, &reader.mu);
        const size_t got = reader.got[cur];
        pthread_mutex_unlock(&reader.mu);
        unsigned char *frames = reader.buf[cur];
        /* a short last slab is zero padded (the reference codes stale bytes of its buffer there) */
        if (got < bytes) memset(frames + got, 0, bytes - got);
        slabsDone += n;
        framesRead += n * DCT_BLOCK_DEPTH;
        const int last = slabsDone >= totalSlabs;

        /* DCT + quantization + zig-zag + Exp-Golomb on the GPU; complete bytes come back */
        size_t expGolombCodedDataSize = 0;
        int er;
        /* a failed call leaves the carried partial byte untouched, so it can be repeated with a larger buffer */
        while ((er = dct3d_multi_stream_encode(ctx, frames, n * DCT_BLOCK_DEPTH, last, expGolombBuffer, egCap,
                                               &expGolombCodedDataSize)) == DCT3D_E_OVERFLOW && egCap < 4 * batchBytes + 64) {
            dct3d_host_free(expGolombBuffer);
            egCap = egCap * 4 < 4 * batchBytes + 64 ? egCap * 4 : 4 * batchBytes + 64;
            expGolombBuffer = (unsigned char *)dct3d_host_alloc(egCap);
            if (!expGolombBuffer) { printf("Error allocating host buffers\n"); return 1; }
        }
        if (er != DCT3D_OK) {
            printf("Error encoding slab: %s\n", dct3d_multi_last_error(ctx));
            rc = 1;
            break;
        }
        /* the buffer goes back to the reader with the next request */
        pthread_mutex_lock(&reader.mu);
        {
            const int m = totalSlabs - slabsAsked < batchSlabs ? totalSlabs - slabsAsked : batchSlabs;
            reader.want[cur] = slabBytes * (size_t)(m > 0 ? m : 0);
            slabsAsked += m > 0 ? m : 0;
            reader.full[cur] = 0;
        }
        pthread_cond_broadcast(&reader.cv);
        pthread_mutex_unlock(&reader.mu);
        cur ^= 1;

        /* Deflating the Exp-Golomb coded data (queued; the workers run while the next batch is read and coded). */
        if (pdeflate_write(zlibStream, expGolombBuffer, expGolombCodedDataSize)) { printf("Error deflating output\n"); rc = 1; break; }
        printf("Frames processed: %d\n", framesRead);
    }
    pthread_mutex_lock(&reader.mu);
    reader.stop = 1;
    pthread_cond_broadcast(&reader.cv);
    pthread_mutex_unlock(&reader.mu);
    pthread_join(readerThread, NULL);
    if (pdeflate_close(zlibStream, NULL, NULL) && !rc) { printf("Error deflating output\n"); rc = 1; }   /* waits for the deflate workers */
    fflush(outputFile);
    fclose(outputFile);
    close(reader.fd);
    dct3d_multi_destroy(ctx);
    dct3d_host_free(reader.buf[0]); dct3d_host_free(reader.buf[1]); dct3d_host_free(expGolombBuffer);
    if (getenv("DCT3D_CLI_TIMING"))
        fprintf(stderr, "dct3d-cli encode: setup %.3f s, loop %.3f s, total %.3f s\n", t_loop - t_start, now_s() - t_loop, now_s() - t_start);
    if (!rc) printf("Encoding process completed");
    return rc;
}
