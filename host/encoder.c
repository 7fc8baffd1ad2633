/* encoder.c -- the C codec's encode flow over libdct3d.so.
 *
 * Same flow as the reference's 3d-DCT-video-encoding-OpenCL/encoder.c:88-293: read one slab of
 * DCT_BLOCK_DEPTH frames, code it, deflate the complete bytes, carry the partial byte into the
 * next slab, Z_FINISH after the last one.  What changed is what sits between fread and deflate:
 * readCubes' float reshuffle (:10-45), the clEnqueueWriteBuffer / two kernels / clEnqueueReadBuffer
 * sequence (:209-254), applyQuantization (:47-58) and applyExpGolombCoding (:60-71) with
 * expGolomb_freeBuffer (ExpGolomb.c:112-122) are ONE call, dct3d_stream_encode, which takes the raw
 * u8 frames and returns the complete stream bytes; OpenCLUtils.c is replaced by dct3d_create.
 * The container stage (deflate, :73-86,266-274) keeps its format -- one zlib stream, level 9 -- but runs
 * on all host cores (pdeflate.c) while the next slab is read and coded; environment variables
 * DCT3D_ZLIB_LEVEL / DCT3D_ZLIB_THREADS override level and thread count.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "../include/dct3d.h"
#include "codec.h"
#include "pdeflate.h"

int encode(char *inputFileName, char *outputFileName, int width, int height, int framesToEncode, int platformIndex)
{
    const size_t bufferSize = (size_t)width * height * DCT_BLOCK_DEPTH;
    FILE *inputFile = fopen(inputFileName, "rb");
    FILE *outputFile = fopen(outputFileName, "wb");
    if (!inputFile || !outputFile) { printf("Error opening files\n"); return 1; }
    /* page-locked staging buffers: the slab goes to the GPU at PCIe speed */
    unsigned char *frames = (unsigned char *)dct3d_host_alloc(bufferSize);
    const size_t egCap = 4 * bufferSize + 64;                                          /* worst case, bounds-checked by the library */
    unsigned char *expGolombBuffer = (unsigned char *)dct3d_host_alloc(egCap);
    if (!frames || !expGolombBuffer) { printf("Error allocating host buffers\n"); return 1; }

    const char *lv = getenv("DCT3D_ZLIB_LEVEL"), *th = getenv("DCT3D_ZLIB_THREADS");
    pdeflate *zlibStream = pdeflate_open(outputFile, lv ? atoi(lv) : Z_BEST_COMPRESSION, th ? atoi(th) : 0, 0);
    if (!zlibStream) { printf("Error starting deflate\n"); return 1; }

    printf("Getting device id\n");
    dct3d_ctx *ctx = NULL;
    if (dct3d_create(&ctx, platformIndex - 1, width, height, DCT_BLOCK_WIDTH) != DCT3D_OK) {
        printf("Error creating dct3d context: %s\n", dct3d_last_error(NULL));
        return 1;
    }
    dct3d_stream_begin(ctx);

    printf("Starting encoding process\n");
    int framesRead = 0;
    while (framesRead < framesToEncode) {
        /* Reading frames (a short last slab is zero padded; the reference codes stale bytes there) */
        size_t got = 0, r;
        while (got < bufferSize && (r = fread(frames + got, 1, bufferSize - got, inputFile)) > 0) got += r;
        if (got < bufferSize) memset(frames + got, 0, bufferSize - got);
        framesRead += DCT_BLOCK_DEPTH;
        const int last = !(framesToEncode > framesRead);

        /* DCT + quantization + zig-zag + Exp-Golomb on the GPU; complete bytes come back */
        size_t expGolombCodedDataSize = 0;
        if (dct3d_stream_encode(ctx, frames, DCT_BLOCK_DEPTH, last, expGolombBuffer, egCap,
                                &expGolombCodedDataSize) != DCT3D_OK) {
            printf("Error encoding slab: %s\n", dct3d_last_error(ctx));
            return 1;
        }

        /* Deflating the Exp-Golomb coded data (queued; the workers run while the next slab is read). */
        if (pdeflate_write(zlibStream, expGolombBuffer, expGolombCodedDataSize)) { printf("Error deflating output\n"); return 1; }
        printf("Frames processed: %d\n", framesRead);
    }
    if (pdeflate_close(zlibStream, NULL, NULL)) { printf("Error deflating output\n"); return 1; }
    fflush(outputFile);
    fclose(outputFile);
    fclose(inputFile);
    dct3d_destroy(ctx);
    dct3d_host_free(frames); dct3d_host_free(expGolombBuffer);
    printf("Encoding process completed");
    return 0;
}
