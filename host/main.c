/* main.c -- the C codec's command line, same shape as the reference's
 * 3d-DCT-video-encoding-OpenCL/main.c:5-49:
 *   codec list_platforms
 *   codec encode|decode <input file> <output file> <width> <height> <nr of frames> [device_index | i,j,k,...]
 * list_platforms lists CUDA devices (dct3d_list_devices) instead of OpenCL platforms. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dct3d.h"
#include "codec.h"

static void printUsage(void)
{
    printf("Usage\n\n");
    printf("codec list_platforms -> List available CUDA devices\n");
    printf("codec encode|decode <input file> <output file> <width> <height> <nr of frames to encode/decode> "
           "<device_index (optional)> -> Encode/Decode given file");
}

int codec_devices(int platformIndex, int *devices, int max)
{
    const char *list = getenv("DCT3D_DEVICES");
    int n = 0;
    while (list && *list && n < max) {
        devices[n++] = atoi(list) - 1;
        list = strchr(list, ',');
        if (list) list++;
    }
    if (n == 0) devices[n++] = platformIndex - 1;
    return n;
}

int main(int argc, char *argv[])
{
    if (argc < 2) { printUsage(); exit(0); }
    if (!strcmp(argv[1], "list_platforms")) {
        char buf[4096];
        if (dct3d_list_devices(buf, sizeof buf) < 0) { printf("%s\n", dct3d_last_error(NULL)); return 1; }
        printf("%s", buf);
    } else if (argc >= 7) {
        int width = atoi(argv[4]), height = atoi(argv[5]), framesToProcess = atoi(argv[6]);
        int platformIndex = argc > 7 ? atoi(argv[7]) : 1;
        if (argc > 7 && strchr(argv[7], ',')) setenv("DCT3D_DEVICES", argv[7], 1);    /* "1,2,3,4": slab ranges over several GPUs */
        if (!strcmp(argv[1], "encode")) return encode(argv[2], argv[3], width, height, framesToProcess, platformIndex);
        else if (!strcmp(argv[1], "decode")) return decode(argv[2], argv[3], width, height, framesToProcess, platformIndex);
        else printUsage();
    } else {
        printUsage();
    }
    return 0;
}
