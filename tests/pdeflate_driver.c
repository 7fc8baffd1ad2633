/* Test driver for host/pdeflate.c: stdin -> one zlib stream on stdout.
 * usage: pdeflate_driver LEVEL THREADS BLOCK CHUNK   (CHUNK = bytes per pdeflate_write call) */
#include <stdio.h>
#include <stdlib.h>
#include "../host/pdeflate.h"

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const int level = atoi(argv[1]), threads = atoi(argv[2]);
    const size_t block = (size_t)atol(argv[3]), chunk = (size_t)atol(argv[4]);
    pdeflate *p = pdeflate_open(stdout, level, threads, block);
    if (!p) return 3;
    unsigned char *buf = (unsigned char *)malloc(chunk ? chunk : 1);
    size_t n;
    while (chunk && (n = fread(buf, 1, chunk, stdin)) > 0)
        if (pdeflate_write(p, buf, n)) return 4;
    unsigned long long in = 0, out = 0;
    if (pdeflate_close(p, &in, &out)) return 5;
    fprintf(stderr, "%llu %llu\n", in, out);
    free(buf);
    return 0;
}
