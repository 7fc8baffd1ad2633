/* codec.h -- entry points of the C codec, same signatures as the reference's
 * 3d-DCT-video-encoding-OpenCL/codec.h:17-19.  The last argument was the 1-based OpenCL platform
 * index there; here it is the 1-based CUDA device index. */
#ifndef CODEC_H_
#define CODEC_H_

#define DCT_BLOCK_WIDTH 8
#define DCT_BLOCK_HEIGHT 8
#define DCT_BLOCK_DEPTH 8

int encode(char *inputFileName, char *outputFileName, int width, int height, int framesToEncode, int platformIndex);
int decode(char *inputFileName, char *outputFileName, int width, int height, int framesToDecode, int platformIndex);

/* The GPUs a run uses: the 1-based index of the command line, or the 1-based, comma-separated list in the environment
 * variable DCT3D_DEVICES (main.c puts a comma-separated last argument there).  Returns the count, ordinals 0-based. */
int codec_devices(int platformIndex, int *devices, int max);

#endif /* CODEC_H_ */
