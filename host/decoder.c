/* decoder.c -- the C codec's decode flow over libdct3d.so.
 *
 * Same flow as the reference's 3d-DCT-video-encoding-OpenCL/decoder.c:85-314: read and inflate until enough codes are
 * buffered, decode them, write the frames, drop the consumed bytes and keep the bit position of the partial byte
 * (expGolomb_freeBuffer(..., 0), ExpGolomb.c:123-129).  What changed:
 *   - expGolomb_readValue + reorderDctCoeffs + applyDequantization + the cl* sequence + writeCubes (:229-295) are ONE call,
 *     dct3d_multi_stream_decode (one GPU or several, see encoder.c), which reports DCT3D_E_NEED_MORE while the buffered input does not yet hold all the codes;
 *   - the unit of work is a BATCH of slabs per call (DCT3D_BATCH_SLABS, default 16 = 128 frames per GPU) instead of one slab
 *     (:207), so the library's chunk pipeline overlaps the inverse kernels with the D2H copies;
 *   - three stages run side by side instead of in turn: an inflate thread (the reference's fread + inflate loop,
 *     :210-227; inflate of one zlib stream is inherently serial, so it gets a core of its own and runs ahead), the GPU
 *     decode on the main thread, and a writer thread (the fwrite of :294-295) on the other of two page-locked buffers.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#include "../include/dct3d.h"
#include "codec.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* ---- output: frames of batch i are written while batch i+1 is decoded into the other buffer ------------------ */
typedef struct {
    FILE *out;
    unsigned char *buf[2];
    size_t bytes[2];
    int filled[2], stop, failed;       /* filled[b]: buffer b holds a batch that is not on disk yet */
    pthread_mutex_t mu;
    pthread_cond_t cv;
} batch_writer;

static void *batch_writer_main(void *arg)
{
    batch_writer *w = (batch_writer *)arg;
    int b = 0;
    for (;;) {
        pthread_mutex_lock(&w->mu);
        while (!w->filled[b] && !w->stop) pthread_cond_wait(&w->cv, &w->mu);
        if (!w->filled[b]) { pthread_mutex_unlock(&w->mu); return NULL; }
        const size_t n = w->bytes[b];
        pthread_mutex_unlock(&w->mu);
        const int bad = fwrite(w->buf[b], 1, n, w->out) != n;
        pthread_mutex_lock(&w->mu);
        if (bad) w->failed = 1;
        w->filled[b] = 0;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
        b ^= 1;
    }
}

/* ---- input: the inflate thread appends to a queue of blocks; the main thread moves them into its window -------- */
#define INFLATE_BLOCK ((size_t)8 << 20)
#define INFLATE_AHEAD 64               /* blocks the inflater may run ahead of the decoder (512 MiB) */

typedef struct inflate_block {
    struct inflate_block *next;
    size_t len;
    unsigned char data[];
} inflate_block;

typedef struct {
    FILE *in;
    inflate_block *head, *tail;
    int queued, done, failed, stop;    /* done: the zlib stream (or the file) has ended */
    pthread_mutex_t mu;
    pthread_cond_t cv;
} inflater;

static void *inflater_main(void *arg)
{
    inflater *f = (inflater *)arg;
    z_stream z;
    memset(&z, 0, sizeof z);
    inflateInit(&z);
    unsigned char *src = (unsigned char *)malloc(1 << 20);
    int ended = 0, bad = src == NULL;
    while (!ended && !bad) {
        inflate_block *blk = (inflate_block *)malloc(sizeof(inflate_block) + INFLATE_BLOCK);
        if (!blk) { bad = 1; break; }
        blk->next = NULL;
        z.next_out = blk->data;
        z.avail_out = (uInt)INFLATE_BLOCK;
        while (z.avail_out && !ended) {
            if (z.avail_in == 0) {
                const size_t got = fread(src, 1, 1 << 20, f->in);
                if (got == 0) { ended = 1; break; }                   /* the file ends (possibly before the stream does) */
                z.next_in = src;
                z.avail_in = (uInt)got;
            }
            const int zr = inflate(&z, Z_NO_FLUSH);
            if (zr == Z_STREAM_END) ended = 1;
            else if (zr != Z_OK && zr != Z_BUF_ERROR) { printf("Error inflating input: %d\n", zr); bad = 1; ended = 1; }
        }
        blk->len = INFLATE_BLOCK - z.avail_out;
        pthread_mutex_lock(&f->mu);
        while (f->queued >= INFLATE_AHEAD && !f->stop) pthread_cond_wait(&f->cv, &f->mu);
        if (f->stop) { pthread_mutex_unlock(&f->mu); free(blk); break; }
        if (f->tail) f->tail->next = blk; else f->head = blk;
        f->tail = blk;
        f->queued++;
        pthread_cond_broadcast(&f->cv);
        pthread_mutex_unlock(&f->mu);
    }
    inflateEnd(&z);
    free(src);
    pthread_mutex_lock(&f->mu);
    f->done = 1;
    f->failed = bad;
    pthread_cond_broadcast(&f->cv);
    pthread_mutex_unlock(&f->mu);
    return NULL;
}

int decode(char *inputFileName, char *outputFileName, int width, int height, int framesToDecode, int platformIndex)
{
    const double t_start = now_s();
    const size_t slabBytes = (size_t)width * height * DCT_BLOCK_DEPTH;
    const char *bs = getenv("DCT3D_BATCH_SLABS");
    int devices[64];
    const int ndevices = codec_devices(platformIndex, devices, 64);
    int batchSlabs = bs ? atoi(bs) : 16 * ndevices;              /* every GPU gets 16 slabs of a batch */
    if (batchSlabs < 1) batchSlabs = 1;
    const int totalSlabs = (framesToDecode + DCT_BLOCK_DEPTH - 1) / DCT_BLOCK_DEPTH;   /* decoder.c:207 decodes whole slabs */
    if (batchSlabs > totalSlabs) batchSlabs = totalSlabs > 0 ? totalSlabs : 1;
    const size_t batchBytes = slabBytes * batchSlabs;

    FILE *inputFile = fopen(inputFileName, "rb");
    FILE *outputFile = fopen(outputFileName, "wb");
    if (!inputFile || !outputFile) { printf("Error opening files\n"); return 1; }
    /* the window of inflated, not yet consumed bytes (page-locked: it is the source of the H2D copies) */
    size_t cap = batchBytes / 2 + 2 * INFLATE_BLOCK + 64, have = 0;
    unsigned char *expGolombCodedData = (unsigned char *)dct3d_host_alloc(cap);
    batch_writer writer;
    memset(&writer, 0, sizeof writer);
    writer.out = outputFile;
    writer.buf[0] = (unsigned char *)dct3d_host_alloc(batchBytes);
    writer.buf[1] = (unsigned char *)dct3d_host_alloc(batchBytes);
    if (!expGolombCodedData || !writer.buf[0] || !writer.buf[1]) { printf("Error allocating host buffers\n"); return 1; }
    pthread_mutex_init(&writer.mu, NULL);
    pthread_cond_init(&writer.cv, NULL);
    pthread_t writerThread, inflateThread;
    pthread_create(&writerThread, NULL, batch_writer_main, &writer);
    inflater inf;
    memset(&inf, 0, sizeof inf);
    inf.in = inputFile;
    pthread_mutex_init(&inf.mu, NULL);
    pthread_cond_init(&inf.cv, NULL);
    pthread_create(&inflateThread, NULL, inflater_main, &inf);

    printf("Getting device id\n");
    dct3d_multi *ctx = NULL;
    if (dct3d_multi_create(&ctx, devices, ndevices, width, height, DCT_BLOCK_WIDTH) != DCT3D_OK) {
        printf("Error creating dct3d context: %s\n", dct3d_last_error(NULL));
        return 1;
    }

    printf("Starting decoding process\n");
    const double t_loop = now_s();
    int slabsDone = 0, cur = 0, rc = 0, inflated_all = 0;
    uint64_t bitpos = 0;
    size_t want = 0;                   /* bytes the window should hold before the next attempt (an estimate that only grows on NEED_MORE) */
    size_t bytesPerSlab = 0;           /* measured on the batches decoded so far */
    while (slabsDone < totalSlabs && !rc) {
        const int n = totalSlabs - slabsDone < batchSlabs ? totalSlabs - slabsDone : batchSlabs;
        /* ---- fill the window from the inflater's queue up to `want` bytes (all that is queued if want is unknown) ---- */
        if (want == 0) want = bytesPerSlab ? bytesPerSlab * (size_t)n + bytesPerSlab / 4 + 4096 : slabBytes * (size_t)n / 8;   /* first guess: 1 bit per sample */
        while (have < want && !inflated_all) {
            pthread_mutex_lock(&inf.mu);
            while (!inf.head && !inf.done) pthread_cond_wait(&inf.cv, &inf.mu);
            inflate_block *blk = inf.head;
            if (blk) {
                inf.head = blk->next;
                if (!inf.head) inf.tail = NULL;
                inf.queued--;
                pthread_cond_broadcast(&inf.cv);
            } else {
                inflated_all = 1;
                if (inf.failed) rc = 1;
            }
            pthread_mutex_unlock(&inf.mu);
            if (!blk) break;
            if (cap - have < blk->len) {
                size_t bigger = 2 * cap + blk->len;
                unsigned char *p = (unsigned char *)dct3d_host_alloc(bigger);
                if (!p) { printf("Error allocating host buffers\n"); rc = 1; free(blk); break; }
                memcpy(p, expGolombCodedData, have);
                dct3d_host_free(expGolombCodedData);
                expGolombCodedData = p;
                cap = bigger;
            }
            memcpy(expGolombCodedData + have, blk->data, blk->len);
            have += blk->len;
            free(blk);
        }
        if (rc) break;
        /* the buffer must be back from the writer before it is decoded into again */
        pthread_mutex_lock(&writer.mu);
        while (writer.filled[cur]) pthread_cond_wait(&writer.cv, &writer.mu);
        pthread_mutex_unlock(&writer.mu);
        unsigned char *frames = writer.buf[cur];
        const uint64_t before = bitpos;
        int dr = have ? dct3d_multi_stream_decode(ctx, expGolombCodedData, have, &bitpos, n * DCT_BLOCK_DEPTH, frames) : DCT3D_E_NEED_MORE;
        if (dr == DCT3D_E_NEED_MORE) {
            if (inflated_all) { printf("Input ended before all frames were decoded\n"); rc = 1; break; }
            want = have + (have / 4 > INFLATE_BLOCK ? have / 4 : INFLATE_BLOCK);          /* buffer more and try again */
            continue;
        }
        if (dr != DCT3D_OK) { printf("Error decoding slab: %s\n", dct3d_multi_last_error(ctx)); rc = 1; break; }
        /* Writing the resulting pixels to the output file (handed to the writer thread) */
        pthread_mutex_lock(&writer.mu);
        writer.bytes[cur] = slabBytes * (size_t)n;
        writer.filled[cur] = 1;
        pthread_cond_broadcast(&writer.cv);
        pthread_mutex_unlock(&writer.mu);
        cur ^= 1;
        slabsDone += n;
        /* drop the consumed bytes, keep the partial byte's bit position */
        size_t consumed = (size_t)(bitpos / 8);
        if (consumed > have) consumed = have;                   /* cannot happen: the library never reads past the buffer */
        bytesPerSlab = (size_t)((bitpos - before) / 8 / (uint64_t)n) + 1;
        memmove(expGolombCodedData, expGolombCodedData + consumed, have - consumed);
        have -= consumed;
        bitpos %= 8;
        want = 0;
        printf("Frames processed: %d\n", slabsDone * DCT_BLOCK_DEPTH);
    }
    /* stop the helpers */
    pthread_mutex_lock(&inf.mu);
    inf.stop = 1;
    pthread_cond_broadcast(&inf.cv);
    pthread_mutex_unlock(&inf.mu);
    pthread_join(inflateThread, NULL);
    while (inf.head) { inflate_block *b = inf.head; inf.head = b->next; free(b); }
    pthread_mutex_lock(&writer.mu);
    writer.stop = 1;
    pthread_cond_broadcast(&writer.cv);
    pthread_mutex_unlock(&writer.mu);
    pthread_join(writerThread, NULL);
    if (writer.failed) { printf("Error writing output\n"); rc = 1; }
    fflush(outputFile);
    fclose(outputFile);
    fclose(inputFile);
    dct3d_multi_destroy(ctx);
    dct3d_host_free(expGolombCodedData); dct3d_host_free(writer.buf[0]); dct3d_host_free(writer.buf[1]);
    if (getenv("DCT3D_CLI_TIMING"))
        fprintf(stderr, "dct3d-cli decode: setup %.3f s, loop %.3f s, total %.3f s\n", t_loop - t_start, now_s() - t_loop, now_s() - t_start);
    if (!rc) printf("Decoding process completed");
    return rc;
}
