// micro-benchmark: issue/throughput of FFMA (imm), FFMA (reg), FFMA2 (packed) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 2048
template <int MODE>
__global__ void k(float *out, float c0, float c1)
{
    float a[16];
    float2 p[8];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; i++) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
    const float2 cc0 = make_float2(c0, c0), cc1 = make_float2(c1, c1);
    for (int it = 0; it < N_ITER; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 0.999f, 0.001f);      // immediates
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], c0, c1);              // register operands
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], cc0, cc1);       // packed, register operands
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __fadd2_rn(p[i], cc1);
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = a[i] + 0.001f;
        } else if (MODE == 5) {   // mixed: packed FMA + integer ALU work, to see co-issue
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], cc0, cc1);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __int_as_float(__float_as_int(a[i]) ^ (it + i));
        } else if (MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 0.999f, 0.001f);
#pragma unroll
            for (int i = 0; i < 8; i++) p[i].x = __int_as_float(__float_as_int(p[i].x) ^ (it + i));
        }
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i];
    for (int i = 0; i < 8; i++) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, float *d, double flop_per_thread_iter)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * N_ITER * flop_per_thread_iter;
    printf("%-28s %8.3f ms  %8.2f T lane-ops/s\n", name, ms, ops / ms / 1e9);
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA imm (16/iter)", d, 16);
    run<1>("FFMA reg (16/iter)", d, 16);
    run<2>("FFMA2 reg (8 packed/iter)", d, 16);
    run<3>("FADD2 reg (8 packed/iter)", d, 16);
    run<4>("FADD imm (16/iter)", d, 16);
    run<5>("FFMA2 x8 + LOP x8", d, 16);
    run<6>("FFMA imm x16 + LOP x8", d, 16);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
