/* decoder.c -- the C codec's decode flow over libdct3d.so.
 *
 * Same flow as the reference's 3d-DCT-video-encoding-OpenCL/decoder.c:85-314: read and inflate until
 * one slab's worth of codes is buffered, decode it, write DCT_BLOCK_DEPTH frames, drop the consumed
 * bytes and keep the bit position of the partial byte (expGolomb_freeBuffer(..., 0),
 * ExpGolomb.c:123-129).  expGolomb_readValue + reorderDctCoeffs + applyDequantization + the cl*
 * sequence + writeCubes (:229-295) are ONE call, dct3d_stream_decode, which reports
 * DCT3D_E_NEED_MORE while the buffered input does not yet hold the whole slab.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../include/dct3d.h"
#include "codec.h"

/* Output side of the slab pipeline: the frames of slab i are written by this thread while the main thread
 * inflates and decodes slab i+1 into the other page-locked buffer (replaces the blocking fwrite of
 * writeCubes' caller, decoder.c:294-295). */
typedef struct {
    FILE *out;
    unsigned char *buf[2];
    size_t bytes;
    int filled[2], stop, failed;       /* filled[b]: buffer b holds a slab that is not on disk yet */
    pthread_mutex_t mu;
    pthread_cond_t cv;
} slab_writer;

static void *slab_writer_main(void *arg)
{
    slab_writer *w = (slab_writer *)arg;
    int b = 0;
    for (;;) {
        pthread_mutex_lock(&w->mu);
        while (!w->filled[b] && !w->stop) pthread_cond_wait(&w->cv, &w->mu);
        if (!w->filled[b]) { pthread_mutex_unlock(&w->mu); return NULL; }
        pthread_mutex_unlock(&w->mu);
        const int bad = fwrite(w->buf[b], 1, w->bytes, w->out) != w->bytes;
        pthread_mutex_lock(&w->mu);
        if (bad) w->failed = 1;
        w->filled[b] = 0;
        pthread_cond_broadcast(&w->cv);
        pthread_mutex_unlock(&w->mu);
        b ^= 1;
    }
}

int decode(char *inputFileName, char *outputFileName, int width, int height, int framesToDecode, int platformIndex)
{
    const size_t bufferSize = (size_t)width * height * DCT_BLOCK_DEPTH;
    FILE *inputFile = fopen(inputFileName, "rb");
    FILE *outputFile = fopen(outputFileName, "wb");
    if (!inputFile || !outputFile) { printf("Error opening files\n"); return 1; }
    unsigned char *zlibCompressedData = (unsigned char *)malloc(bufferSize);
    size_t cap = 2 * bufferSize + 64, have = 0;                  /* inflated, not yet consumed */
    /* page-locked buffers on both sides of the GPU call */
    unsigned char *expGolombCodedData = (unsigned char *)dct3d_host_alloc(cap);
    slab_writer writer;
    memset(&writer, 0, sizeof writer);
    writer.out = outputFile;
    writer.bytes = bufferSize;
    writer.buf[0] = (unsigned char *)dct3d_host_alloc(bufferSize);
    writer.buf[1] = (unsigned char *)dct3d_host_alloc(bufferSize);
    if (!zlibCompressedData || !expGolombCodedData || !writer.buf[0] || !writer.buf[1]) { printf("Error allocating host buffers\n"); return 1; }
    pthread_mutex_init(&writer.mu, NULL);
    pthread_cond_init(&writer.cv, NULL);
    pthread_t writerThread;
    pthread_create(&writerThread, NULL, slab_writer_main, &writer);
    int cur = 0;                                                 /* buffer the next slab is decoded into */

    z_stream zlibStream;
    memset(&zlibStream, 0, sizeof zlibStream);
    inflateInit(&zlibStream);

    printf("Getting device id\n");
    dct3d_ctx *ctx = NULL;
    if (dct3d_create(&ctx, platformIndex - 1, width, height, DCT_BLOCK_WIDTH) != DCT3D_OK) {
        printf("Error creating dct3d context: %s\n", dct3d_last_error(NULL));
        return 1;
    }

    printf("Starting decoding process\n");
    int framesRead = 0, eof = 0, inflated_all = 0;
    uint64_t bitpos = 0;
    while (framesRead < framesToDecode) {
        /* the buffer must be back from the writer before it is decoded into again */
        pthread_mutex_lock(&writer.mu);
        while (writer.filled[cur]) pthread_cond_wait(&writer.cv, &writer.mu);
        pthread_mutex_unlock(&writer.mu);
        unsigned char *frames = writer.buf[cur];
        int rc = have ? dct3d_stream_decode(ctx, expGolombCodedData, have, &bitpos, DCT_BLOCK_DEPTH, frames) : DCT3D_E_NEED_MORE;
        if (rc == DCT3D_E_NEED_MORE) {
            if (inflated_all) { printf("Input ended before all frames were decoded\n"); return 1; }
            /* Reading data from file and applying the inflate algorithm */
            if (zlibStream.avail_in == 0 && !eof) {
                size_t got = fread(zlibCompressedData, 1, bufferSize, inputFile);
                if (got == 0) eof = 1;
                zlibStream.next_in = zlibCompressedData;
                zlibStream.avail_in = (uInt)got;
            }
            if (cap - have < bufferSize) {
                unsigned char *bigger = (unsigned char *)dct3d_host_alloc(2 * cap);
                if (!bigger) { printf("Error allocating host buffers\n"); return 1; }
                memcpy(bigger, expGolombCodedData, have);
                dct3d_host_free(expGolombCodedData);
                expGolombCodedData = bigger;
                cap *= 2;
            }
            zlibStream.next_out = expGolombCodedData + have;
            zlibStream.avail_out = (uInt)(cap - have);
            int zr = inflate(&zlibStream, Z_NO_FLUSH);
            have = cap - zlibStream.avail_out;
            if (zr == Z_STREAM_END || (eof && zlibStream.avail_in == 0)) inflated_all = (zr == Z_STREAM_END) || eof;
            if (zr != Z_OK && zr != Z_STREAM_END && zr != Z_BUF_ERROR) { printf("Error inflating input: %d\n", zr); return 1; }
            continue;
        }
        if (rc != DCT3D_OK) { printf("Error decoding slab: %s\n", dct3d_last_error(ctx)); return 1; }
        /* Writing the resulting pixels to the output file (handed to the writer thread) */
        pthread_mutex_lock(&writer.mu);
        writer.filled[cur] = 1;
        pthread_cond_broadcast(&writer.cv);
        pthread_mutex_unlock(&writer.mu);
        cur ^= 1;
        framesRead += DCT_BLOCK_DEPTH;
        /* drop the consumed bytes, keep the partial byte's bit position */
        const size_t consumed = (size_t)(bitpos / 8);
        memmove(expGolombCodedData, expGolombCodedData + consumed, have - consumed);
        have -= consumed;
        bitpos %= 8;
        printf("Frames processed: %d\n", framesRead);
    }
    inflateEnd(&zlibStream);
    pthread_mutex_lock(&writer.mu);
    writer.stop = 1;
    pthread_cond_broadcast(&writer.cv);
    pthread_mutex_unlock(&writer.mu);
    pthread_join(writerThread, NULL);
    if (writer.failed) { printf("Error writing output\n"); return 1; }
    fflush(outputFile);
    fclose(outputFile);
    fclose(inputFile);
    dct3d_destroy(ctx);
    free(zlibCompressedData); dct3d_host_free(expGolombCodedData); dct3d_host_free(writer.buf[0]); dct3d_host_free(writer.buf[1]);
    printf("Decoding process completed");
    return 0;
}
