/*
 * dct3d_oracle.c -- CPU restatement of the reference codec's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: the
 * CUDA library (libdct3d.so) never links, loads or calls this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it, and only as the checker / the reported CPU
 * baseline.
 *
 * Parity pin: the reference ships no tests or golden vectors, so
 * this restatement is pinned against (a) the known-answer vectors generated
 * from the reference's own CubeUtils.c / ExpGolomb.c (SURVEY.md App. C,
 * committed under tests/golden/) and (b) the reference's unmodified C sources
 * compiled into oracle/_ref/ by oracle/Makefile (the test_ref_*_live tests of tests/test_oracle.py).
 *
 * Every function cites the reference file:line it restates.  Paths:
 *   J/ = 3d-DCT-video-encoding/src/br/jpiccoli/video/
 *   C/ = 3d-DCT-video-encoding-OpenCL/
 *
 * Index convention: n2/k2 = column (fastest), n1/k1 = row,
 * n0/k0 = frame.  Cube-major order = [slab][block row][block col][k0][k1][k2].
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ */
/* Zig-zag ("diagonal slices") table.                                        */
/* J/CubeUtils.java:7-41, C/CubeUtils.c:5-46: slices of constant x+y+z in    */
/* ascending order; inside a slice y is the outer loop, z the middle, x the  */
/* inner.  Linear index x + y*cw + z*cw*ch (J/Encoder.java:108,              */
/* C/encoder.c:65).  Returns the number of entries written.                  */
/* ------------------------------------------------------------------------ */
int orc_zigzag(int cw, int ch, int cd, int32_t *lin)
{
    int n = 0;
    int top = (cw - 1) + (ch - 1) + (cd - 1);
    for (int s = 0; s <= top; s++)
        for (int y = 0; y < ch; y++)
            for (int z = 0; z < cd; z++) {
                int x = s - y - z;
                if (x >= 0 && x < cw)
                    lin[n++] = x + y * cw + z * cw * ch;
            }
    return n;
}

/* ------------------------------------------------------------------------ */
/* Signed order-0 Exp-Golomb, MSB first.                                     */
/* J/ExpGolombWriter.java:19-49, C/ExpGolomb.c:32-64: v<=0 -> -2v, v>0 ->     */
/* 2v-1, +1, then (L-1) zero bits followed by the L-bit value.               */
/* Bits are OR-ed into buf starting at absolute bit `start_bit` (bit 0 = MSB  */
/* of byte 0); bytes past the start must be zero on entry.  Returns the end   */
/* bit, or UINT64_MAX if the buffer (cap bytes) would overflow.              */
/* ------------------------------------------------------------------------ */
static inline uint64_t eg_map(int32_t v)
{
    int64_t w = v;
    return (uint64_t)(w <= 0 ? -2 * w : 2 * w - 1) + 1;
}

static inline int bitlen64(uint64_t m)
{
    int l = 0;
    while (m) { l++; m >>= 1; }
    return l;
}

uint64_t orc_eg_encode(const int32_t *v, size_t n, uint8_t *buf, size_t cap, uint64_t start_bit)
{
    uint64_t pos = start_bit;
    for (size_t i = 0; i < n; i++) {
        uint64_t m = eg_map(v[i]);
        int L = bitlen64(m);
        uint64_t end = pos + (uint64_t)(2 * L - 1);
        if ((end >> 3) >= cap) return UINT64_MAX; /* keeps the "+1" byte in range too */
        pos += (uint64_t)(L - 1);
        for (int b = L - 1; b >= 0; b--, pos++)
            if ((m >> b) & 1) buf[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
    }
    return pos;
}

/* Code length in bits of one value (2L-1). */
int orc_eg_codelen(int32_t v) { return 2 * bitlen64(eg_map(v)) - 1; }

/* J/ExpGolombReader.java:19-63, C/ExpGolomb.c:66-110.  Reads n values from   */
/* absolute bit start_bit; returns the end bit, UINT64_MAX if it runs past    */
/* nbytes.                                                                   */
uint64_t orc_eg_decode(const uint8_t *buf, size_t nbytes, uint64_t start_bit, size_t n, int32_t *out)
{
    uint64_t pos = start_bit, lim = (uint64_t)nbytes * 8;
    for (size_t i = 0; i < n; i++) {
        int z = 0;
        for (;;) {
            if (pos >= lim) return UINT64_MAX;
            if (buf[pos >> 3] & (0x80u >> (pos & 7))) break;
            z++; pos++;
        }
        uint64_t m = 0;
        for (int b = 0; b <= z; b++, pos++) {
            if (pos >= lim) return UINT64_MAX;
            m = (m << 1) | ((buf[pos >> 3] >> (7 - (pos & 7))) & 1u);
        }
        int64_t w = (int64_t)m - 1;
        out[i] = (int32_t)((w & 1) ? (w + 1) / 2 : -(w / 2));
    }
    return pos;
}

/* ------------------------------------------------------------------------ */
/* Transform constants.  J/dct/Transform.java:20-21, J/dct/DCT.java:81-85,112 */
/*   scale = sqrt(2^3) / sqrt(cw*ch*cd); c(0) = 1/sqrt(2); angle =            */
/*   (PI / (float)N) * (n + 0.5f) * k evaluated in double.                   */
/* ------------------------------------------------------------------------ */
static double basis_cos(int N, int n, int k)
{
    return cos((M_PI / (double)(float)N) * (double)((n + 0.5f) * (float)k));
}

static double ck(int k) { return k == 0 ? 1.0 / sqrt(2.0) : 1.0; }

/* Direct O(cs^2) evaluation per cube on planar data; the plainest statement  */
/* of the maths: J/dct/DCT.java:184-225 (dct3d).  in/out are planar           */
/* [F][H][W]; out is overwritten.                                            */
void orc_dct3d_direct_f64(const double *in, double *out, int W, int H, int F, int cube)
{
    const int c = cube;
    const double scale = sqrt(8.0) / sqrt((double)(c * c * c));
    const size_t fs = (size_t)W * H;
    for (int z = 0; z + c <= F; z += c)
    for (int y = 0; y < H; y += c)
    for (int x = 0; x < W; x += c) {
        size_t off = (size_t)z * fs + (size_t)y * W + x;
        for (int k0 = 0; k0 < c; k0++)
        for (int k1 = 0; k1 < c; k1++)
        for (int k2 = 0; k2 < c; k2++) {
            double acc = 0.0;
            for (int n0 = 0; n0 < c; n0++)
            for (int n1 = 0; n1 < c; n1++)
            for (int n2 = 0; n2 < c; n2++)
                acc += in[off + n0 * fs + (size_t)n1 * W + n2] *
                       basis_cos(c, n0, k0) * basis_cos(c, n1, k1) * basis_cos(c, n2, k2);
            out[off + k0 * fs + (size_t)k1 * W + k2] = scale * ck(k0) * ck(k1) * ck(k2) * acc;
        }
    }
}

/* J/dct/InverseDCT.java:135-178 (idct3d): inverse + clamp to [0,255].        */
void orc_idct3d_direct_f64(const double *in, double *out, int W, int H, int F, int cube)
{
    const int c = cube;
    const double scale = sqrt(8.0) / sqrt((double)(c * c * c));
    const size_t fs = (size_t)W * H;
    for (int z = 0; z + c <= F; z += c)
    for (int y = 0; y < H; y += c)
    for (int x = 0; x < W; x += c) {
        size_t off = (size_t)z * fs + (size_t)y * W + x;
        for (int n0 = 0; n0 < c; n0++)
        for (int n1 = 0; n1 < c; n1++)
        for (int n2 = 0; n2 < c; n2++) {
            double acc = 0.0;
            for (int k0 = 0; k0 < c; k0++)
            for (int k1 = 0; k1 < c; k1++)
            for (int k2 = 0; k2 < c; k2++)
                acc += ck(k0) * ck(k1) * ck(k2) * in[off + k0 * fs + (size_t)k1 * W + k2] *
                       basis_cos(c, n0, k0) * basis_cos(c, n1, k1) * basis_cos(c, n2, k2);
            acc *= scale;
            if (acc > 255.0) acc = 255.0;
            if (acc < 0.0) acc = 0.0;
            out[off + n0 * fs + (size_t)n1 * W + n2] = acc;
        }
    }
}

/* Separable fp64 evaluation of the same transform (per axis                  */
/* sqrt(2/N)*c(k)*cos(...)).  Mathematically identical to   */
/* the direct form; differs by summation order only (~1e-13 relative), which  */
/* tests/test_oracle.py checks.  Used where the direct form is too slow.      */
static void axis_matrix(int N, double *m /* [k][n] */)
{
    for (int k = 0; k < N; k++)
        for (int n = 0; n < N; n++)
            m[k * N + n] = sqrt(2.0 / N) * ck(k) * basis_cos(N, n, k);
}

static void cube_sep(const double *src, double *dst, int c, const double *m, int inverse)
{
    /* src/dst are c*c*c cubes in [a0][a1][a2] order; apply along each axis. */
    double t1[512], t2[512];
    const int cc = c * c;
    for (int a0 = 0; a0 < c; a0++) for (int a1 = 0; a1 < c; a1++) for (int o = 0; o < c; o++) {
        double s = 0;
        for (int i = 0; i < c; i++) s += src[a0 * cc + a1 * c + i] * (inverse ? m[i * c + o] : m[o * c + i]);
        t1[a0 * cc + a1 * c + o] = s;
    }
    for (int a0 = 0; a0 < c; a0++) for (int a2 = 0; a2 < c; a2++) for (int o = 0; o < c; o++) {
        double s = 0;
        for (int i = 0; i < c; i++) s += t1[a0 * cc + i * c + a2] * (inverse ? m[i * c + o] : m[o * c + i]);
        t2[a0 * cc + o * c + a2] = s;
    }
    for (int a1 = 0; a1 < c; a1++) for (int a2 = 0; a2 < c; a2++) for (int o = 0; o < c; o++) {
        double s = 0;
        for (int i = 0; i < c; i++) s += t2[i * cc + a1 * c + a2] * (inverse ? m[i * c + o] : m[o * c + i]);
        dst[o * cc + a1 * c + a2] = s;
    }
}

static void planar_sep(const double *in, double *out, int W, int H, int F, int c, int inverse)
{
    double m[64], a[512], b[512];
    axis_matrix(c, m);
    const size_t fs = (size_t)W * H;
    for (int z = 0; z + c <= F; z += c)
    for (int y = 0; y < H; y += c)
    for (int x = 0; x < W; x += c) {
        size_t off = (size_t)z * fs + (size_t)y * W + x;
        for (int i0 = 0; i0 < c; i0++) for (int i1 = 0; i1 < c; i1++) for (int i2 = 0; i2 < c; i2++)
            a[(i0 * c + i1) * c + i2] = in[off + i0 * fs + (size_t)i1 * W + i2];
        cube_sep(a, b, c, m, inverse);
        for (int i0 = 0; i0 < c; i0++) for (int i1 = 0; i1 < c; i1++) for (int i2 = 0; i2 < c; i2++) {
            double v = b[(i0 * c + i1) * c + i2];
            if (inverse) { if (v > 255.0) v = 255.0; if (v < 0.0) v = 0.0; } /* J/dct/InverseDCT.java:74-80 */
            out[off + i0 * fs + (size_t)i1 * W + i2] = v;
        }
    }
}

void orc_dct3d_sep_f64(const double *in, double *out, int W, int H, int F, int cube) { planar_sep(in, out, W, H, F, cube, 0); }
void orc_idct3d_sep_f64(const double *in, double *out, int W, int H, int F, int cube) { planar_sep(in, out, W, H, F, cube, 1); }

/* ------------------------------------------------------------------------ */
/* Quantise + planar -> cube-major reshuffle.                                */
/* J/Encoder.java:69-89 (Math.round = floor(v+0.5), :82) and                  */
/* C/encoder.c:47-58 (libm round = half away from zero, :53).                 */
/* mode 0 = Java rounding, 1 = C rounding.                                    */
/* ------------------------------------------------------------------------ */
static double qdiv(int k0, int k1, int k2)
{
    int s = 5 * (k0 + k1 + k2);
    return s < 1 ? 1.0 : (double)s;
}

static int32_t round_mode(double v, int mode)
{
    return (int32_t)(mode == 0 ? floor(v + 0.5) : round(v));
}

void orc_quantize_planar(const double *coef, int32_t *q, int W, int H, int F, int cube, int mode)
{
    const int c = cube;
    const size_t fs = (size_t)W * H;
    size_t o = 0;
    for (int z = 0; z + c <= F; z += c)
    for (int y = 0; y < H; y += c)
    for (int x = 0; x < W; x += c)
        for (int k0 = 0; k0 < c; k0++) for (int k1 = 0; k1 < c; k1++) for (int k2 = 0; k2 < c; k2++)
            q[o++] = round_mode(coef[(size_t)(z + k0) * fs + (size_t)(y + k1) * W + x + k2] / qdiv(k0, k1, k2), mode);
}

/* Dequantise + cube-major -> planar.  J/Decoder.java:78-96 (:89),            */
/* C/decoder.c:48-59 (:54).  Exact in integers.                               */
void orc_dequantize_planar(const int32_t *q, double *coef, int W, int H, int F, int cube)
{
    const int c = cube;
    const size_t fs = (size_t)W * H;
    size_t o = 0;
    for (int z = 0; z + c <= F; z += c)
    for (int y = 0; y < H; y += c)
    for (int x = 0; x < W; x += c)
        for (int k0 = 0; k0 < c; k0++) for (int k1 = 0; k1 < c; k1++) for (int k2 = 0; k2 < c; k2++)
            coef[(size_t)(z + k0) * fs + (size_t)(y + k1) * W + x + k2] = (double)q[o++] * qdiv(k0, k1, k2);
}

/* ------------------------------------------------------------------------ */
/* Whole-path helpers on u8 frames (frame-major, J/Encoder.java:47-56).       */
/* ------------------------------------------------------------------------ */

/* u8 frames -> quantised cube-major int32 (natural order inside the cube).   */
/* Also returns the fp64 planar coefficients if coef_out != NULL.            */
void orc_quantized_cubes_u8(const uint8_t *frames, int W, int H, int F, int cube, int mode,
                            int32_t *q, double *coef_out)
{
    size_t n = (size_t)W * H * (F - F % cube);
    double *px = (double *)malloc(n * sizeof(double));
    double *cf = coef_out ? coef_out : (double *)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) px[i] = (double)frames[i];
    orc_dct3d_sep_f64(px, cf, W, H, F - F % cube, cube);
    orc_quantize_planar(cf, q, W, H, F - F % cube, cube, mode);
    free(px);
    if (!coef_out) free(cf);
}

/* Cube-major quantised values -> Exp-Golomb stream in zig-zag order.         */
/* J/Encoder.java:98-111, C/encoder.c:60-71.  Returns end bit.               */
uint64_t orc_eg_encode_cubes(const int32_t *q, size_t ncubes, int cube, uint8_t *buf, size_t cap, uint64_t start_bit)
{
    int32_t zz[512], tmp[512];
    int cs = orc_zigzag(cube, cube, cube, zz);
    uint64_t pos = start_bit;
    for (size_t c = 0; c < ncubes; c++) {
        for (int i = 0; i < cs; i++) tmp[i] = q[c * cs + zz[i]];
        pos = orc_eg_encode(tmp, (size_t)cs, buf, cap, pos);
        if (pos == UINT64_MAX) return pos;
    }
    return pos;
}

/* Inverse: J/Decoder.java:61-76, C/decoder.c:61-72 + C/ExpGolomb.c:66-110.   */
uint64_t orc_eg_decode_cubes(const uint8_t *buf, size_t nbytes, uint64_t start_bit, size_t ncubes, int cube, int32_t *q)
{
    int32_t zz[512], tmp[512];
    int cs = orc_zigzag(cube, cube, cube, zz);
    uint64_t pos = start_bit;
    for (size_t c = 0; c < ncubes; c++) {
        pos = orc_eg_decode(buf, nbytes, pos, (size_t)cs, tmp);
        if (pos == UINT64_MAX) return pos;
        for (int i = 0; i < cs; i++) q[c * cs + zz[i]] = tmp[i];
    }
    return pos;
}

/* u8 frames -> Exp-Golomb stream (the hot path of J/Encoder.java:51-111).    */
/* Stream length in bytes is floor(bits/8)+1 (J/Encoder.java:117,             */
/* C/encoder.c:270).  buf must be zeroed.  Returns total bits.               */
uint64_t orc_encode_u8(const uint8_t *frames, int W, int H, int F, int cube, int mode, uint8_t *buf, size_t cap)
{
    int Fe = F - F % cube; /* J/Encoder.java:39-40 */
    size_t n = (size_t)W * H * Fe;
    int32_t *q = (int32_t *)malloc(n * sizeof(int32_t));
    orc_quantized_cubes_u8(frames, W, H, Fe, cube, mode, q, NULL);
    uint64_t bits = orc_eg_encode_cubes(q, n / (size_t)(cube * cube * cube), cube, buf, cap, 0);
    free(q);
    return bits;
}

/* Quantised cube-major values -> u8 frames: dequantise, inverse, clamp,      */
/* truncate.  J/Decoder.java:78-117 (:112 truncation), C/decoder.c:29.        */
void orc_reconstruct_u8(const int32_t *q, int W, int H, int F, int cube, uint8_t *frames)
{
    size_t n = (size_t)W * H * F;
    double *cf = (double *)malloc(n * sizeof(double));
    double *px = (double *)malloc(n * sizeof(double));
    orc_dequantize_planar(q, cf, W, H, F, cube);
    orc_idct3d_sep_f64(cf, px, W, H, F, cube);
    for (size_t i = 0; i < n; i++) frames[i] = (uint8_t)px[i];
    free(cf); free(px);
}

/* Exp-Golomb stream -> u8 frames (J/Decoder.java:61-117). 0 ok, -1 truncated. */
int orc_decode_u8(const uint8_t *buf, size_t nbytes, int W, int H, int F, int cube, uint8_t *frames)
{
    int Fe = F - F % cube;
    size_t n = (size_t)W * H * Fe;
    int32_t *q = (int32_t *)malloc(n * sizeof(int32_t));
    uint64_t end = orc_eg_decode_cubes(buf, nbytes, 0, n / (size_t)(cube * cube * cube), cube, q);
    if (end == UINT64_MAX) { free(q); return -1; }
    orc_reconstruct_u8(q, W, H, Fe, cube, frames);
    free(q);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Float restatement of the OpenCL kernels' semantics (C/3dDCT.cl:43-87,      */
/* 103-143, 164-221, 237-265) on cube-major float data, with cosf standing in  */
/* for native_cos (implementation-defined): per coefficient */
/* a float sum of value*cos*cos*cos over the cube, then *scale*c0*c1*c2.       */
/* The summation is sequential, not the work-group tree of the kernel; both    */
/* are float sums of the same 512 products.                                   */
/* ------------------------------------------------------------------------ */
void orc_cl_dct_f32(const float *in, float *out, size_t ncubes, int cube)
{
    const int c = cube, cc = c * c, cs = cc * c;
    const float pic = (float)M_PI / (float)c;
    const float scale = sqrtf(8.0f / (float)cs);
    float cs_tab[64];
    for (int n = 0; n < c; n++) for (int k = 0; k < c; k++) cs_tab[n * c + k] = cosf(pic * (n + 0.5f) * k);
    for (size_t q = 0; q < ncubes; q++) {
        const float *src = in + q * cs; float *dst = out + q * cs;
        for (int k0 = 0; k0 < c; k0++) for (int k1 = 0; k1 < c; k1++) for (int k2 = 0; k2 < c; k2++) {
            float acc = 0.0f;
            for (int n0 = 0; n0 < c; n0++) for (int n1 = 0; n1 < c; n1++) for (int n2 = 0; n2 < c; n2++)
                acc += src[n0 * cc + n1 * c + n2] * cs_tab[n0 * c + k0] * cs_tab[n1 * c + k1] * cs_tab[n2 * c + k2];
            float c0 = k0 ? 1.0f : (float)M_SQRT1_2, c1 = k1 ? 1.0f : (float)M_SQRT1_2, c2 = k2 ? 1.0f : (float)M_SQRT1_2;
            dst[k0 * cc + k1 * c + k2] = acc * scale * c0 * c1 * c2;
        }
    }
}

void orc_cl_idct_f32(const float *in, float *out, size_t ncubes, int cube)
{
    const int c = cube, cc = c * c, cs = cc * c;
    const float pic = (float)M_PI / (float)c;
    const float scale = sqrtf(8.0f / (float)cs);
    float cs_tab[64];
    for (int n = 0; n < c; n++) for (int k = 0; k < c; k++) cs_tab[n * c + k] = cosf(pic * (n + 0.5f) * k);
    for (size_t q = 0; q < ncubes; q++) {
        const float *src = in + q * cs; float *dst = out + q * cs;
        for (int n0 = 0; n0 < c; n0++) for (int n1 = 0; n1 < c; n1++) for (int n2 = 0; n2 < c; n2++) {
            float acc = 0.0f;
            for (int k0 = 0; k0 < c; k0++) for (int k1 = 0; k1 < c; k1++) for (int k2 = 0; k2 < c; k2++) {
                float c0 = k0 ? 1.0f : (float)M_SQRT1_2, c1 = k1 ? 1.0f : (float)M_SQRT1_2, c2 = k2 ? 1.0f : (float)M_SQRT1_2;
                acc += src[k0 * cc + k1 * c + k2] * c0 * c1 * c2 * cs_tab[n0 * c + k0] * cs_tab[n1 * c + k1] * cs_tab[n2 * c + k2];
            }
            acc *= scale;
            if (acc > 255.0f) acc = 255.0f; else if (acc < 0.0f) acc = 0.0f;
            dst[n0 * cc + n1 * c + n2] = acc;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* CPU baseline port: the Java encoder/decoder's work structure in C.         */
/* Forward: J/dct/DCT.java:77-163 groups, per output coefficient, the inputs   */
/* whose basis product is equal to 1e-9 ((long)(c*1e9), :115), drops zero      */
/* ones (:116), multiplies each group's memoised sum once (:41-59 with         */
/* J/dct/Sum.java:41-52).  Inverse: J/dct/InverseDCT.java:33-82 gathers the    */
/* non-zero inputs (|v|>1e-9) and accumulates a cs x cs table.  One task per   */
/* cube on `threads` workers (J/dct/Transform.java:63-104); quantise and       */
/* Exp-Golomb single-threaded as upstream.  Used only for the reported CPU     */
/* baseline (bench.py) and cross-checked against the separable form in tests.  */
/* ------------------------------------------------------------------------ */
typedef struct { int nsums; int *sum_start; int *sum_off; /* per sum: offsets (cube-local idx) */
                 int *out_start; int *mul_sum; double *mul_coef; } jplan_t;

static jplan_t *jplan_build(int c)
{
    const int cc = c * c, cs = cc * c;
    const double scale = sqrt(8.0) / sqrt((double)cs);
    jplan_t *p = (jplan_t *)calloc(1, sizeof(jplan_t));
    /* worst case: cs groups per output, each with its own sum */
    long long *keys = (long long *)malloc(sizeof(long long) * cs);
    unsigned char *member = (unsigned char *)malloc((size_t)cs * cs); /* group membership bitmaps */
    int maxmul = cs * cs;
    p->out_start = (int *)malloc(sizeof(int) * (cs + 1));
    p->mul_sum = (int *)malloc(sizeof(int) * maxmul);
    p->mul_coef = (double *)malloc(sizeof(double) * maxmul);
    /* distinct sums: store membership bitmap per sum for equality lookup */
    int sums_cap = 4096, nsums = 0;
    unsigned char *sum_maps = (unsigned char *)malloc((size_t)sums_cap * cs);
    int nmul = 0;
    for (int k0 = 0, o = 0; k0 < c; k0++) for (int k1 = 0; k1 < c; k1++) for (int k2 = 0; k2 < c; k2++, o++) {
        p->out_start[o] = nmul;
        int ng = 0;
        double *gcoef = p->mul_coef + nmul;
        for (int n0 = 0; n0 < c; n0++) for (int n1 = 0; n1 < c; n1++) for (int n2 = 0; n2 < c; n2++) {
            double co = scale * ck(k0) * ck(k1) * ck(k2) * basis_cos(c, n0, k0) * basis_cos(c, n1, k1) * basis_cos(c, n2, k2);
            long long key = (long long)(co * 1e9);
            if (key == 0) continue;
            int g = 0;
            while (g < ng && keys[g] != key) g++;
            if (g == ng) { keys[ng] = key; gcoef[ng] = co; memset(member + (size_t)ng * cs, 0, cs); ng++; }
            member[(size_t)g * cs + n0 * cc + n1 * c + n2] = 1;
        }
        for (int g = 0; g < ng; g++) {
            int s = 0;
            while (s < nsums && memcmp(sum_maps + (size_t)s * cs, member + (size_t)g * cs, cs)) s++;
            if (s == nsums) {
                if (nsums == sums_cap) { sums_cap *= 2; sum_maps = (unsigned char *)realloc(sum_maps, (size_t)sums_cap * cs); }
                memcpy(sum_maps + (size_t)nsums * cs, member + (size_t)g * cs, cs);
                nsums++;
            }
            p->mul_sum[nmul + g] = s;
        }
        nmul += ng;
    }
    p->out_start[cs] = nmul;
    p->nsums = nsums;
    p->sum_start = (int *)malloc(sizeof(int) * (nsums + 1));
    int tot = 0;
    for (int s = 0; s < nsums; s++) for (int i = 0; i < cs; i++) tot += sum_maps[(size_t)s * cs + i];
    p->sum_off = (int *)malloc(sizeof(int) * tot);
    tot = 0;
    for (int s = 0; s < nsums; s++) {
        p->sum_start[s] = tot;
        for (int i = 0; i < cs; i++) if (sum_maps[(size_t)s * cs + i]) p->sum_off[tot++] = i;
    }
    p->sum_start[nsums] = tot;
    free(keys); free(member); free(sum_maps);
    return p;
}

static void jplan_free(jplan_t *p)
{
    free(p->sum_start); free(p->sum_off); free(p->out_start); free(p->mul_sum); free(p->mul_coef); free(p);
}

/* Plan statistics, for checking against SURVEY.md App. D (11 567 multiplications, 2 319 sums for 8^3). */
void orc_java_plan_stats(int cube, int *nmul, int *nsums, int *nadds)
{
    jplan_t *p = jplan_build(cube);
    int cs = cube * cube * cube;
    *nmul = p->out_start[cs]; *nsums = p->nsums; *nadds = p->sum_start[p->nsums];
    jplan_free(p);
}

typedef struct {
    const double *in; double *out; int W, H, F, c; int inverse;
    const jplan_t *plan; const double *itab; /* inverse: [n][k] table */
    volatile long *next; long ncubes; int bx, by;
} jwork_t;

static void java_cube_forward(const jwork_t *w, size_t off)
{
    const int c = w->c, cc = c * c, cs = cc * c;
    const size_t fs = (size_t)w->W * w->H;
    const jplan_t *p = w->plan;
    double cache[4096]; unsigned char have[4096];
    double *cachep = p->nsums <= 4096 ? cache : (double *)malloc(sizeof(double) * p->nsums);
    unsigned char *havep = p->nsums <= 4096 ? have : (unsigned char *)malloc(p->nsums);
    memset(havep, 0, p->nsums);
    for (int o = 0; o < cs; o++) {
        double acc = 0.0;
        for (int m = p->out_start[o]; m < p->out_start[o + 1]; m++) {
            int s = p->mul_sum[m];
            if (!havep[s]) {
                double v = 0.0;
                for (int i = p->sum_start[s]; i < p->sum_start[s + 1]; i++) {
                    int l = p->sum_off[i];
                    v += w->in[off + (size_t)(l / cc) * fs + (size_t)((l % cc) / c) * w->W + (l % c)];
                }
                cachep[s] = v; havep[s] = 1;
            }
            acc += cachep[s] * p->mul_coef[m];
        }
        w->out[off + (size_t)(o / cc) * fs + (size_t)((o % cc) / c) * w->W + (o % c)] = acc;
    }
    if (cachep != cache) { free(cachep); free(havep); }
}

static void java_cube_inverse(const jwork_t *w, size_t off)
{
    const int c = w->c, cc = c * c, cs = cc * c;
    const size_t fs = (size_t)w->W * w->H;
    double nzv[512]; int nzi[512]; int nnz = 0;
    for (int k = 0; k < cs; k++) {
        double v = w->in[off + (size_t)(k / cc) * fs + (size_t)((k % cc) / c) * w->W + (k % c)];
        if (fabs(v) > 1e-9) { nzv[nnz] = v; nzi[nnz++] = k; }
    }
    for (int n = 0; n < cs; n++) {
        double acc = 0.0;
        const double *row = w->itab + (size_t)n * cs;
        for (int i = 0; i < nnz; i++) acc += nzv[i] * row[nzi[i]];
        if (acc > 255.0) acc = 255.0;
        if (acc < 0.0) acc = 0.0;
        w->out[off + (size_t)(n / cc) * fs + (size_t)((n % cc) / c) * w->W + (n % c)] = acc;
    }
}

static void *java_worker(void *arg)
{
    jwork_t *w = (jwork_t *)arg;
    const size_t fs = (size_t)w->W * w->H;
    for (;;) {
        long i = __sync_fetch_and_add(w->next, 1);
        if (i >= w->ncubes) break;
        long perslab = (long)w->bx * w->by;
        long z = i / perslab, r = i % perslab;
        size_t off = (size_t)z * w->c * fs + (size_t)(r / w->bx) * w->c * w->W + (size_t)(r % w->bx) * w->c;
        if (w->inverse) java_cube_inverse(w, off); else java_cube_forward(w, off);
    }
    return NULL;
}

static void java_transform(const double *in, double *out, int W, int H, int F, int cube, int inverse, int threads)
{
    const int c = cube, cs = c * c * c;
    jwork_t w; memset(&w, 0, sizeof w);
    volatile long next = 0;
    w.in = in; w.out = out; w.W = W; w.H = H; w.F = F; w.c = c; w.inverse = inverse;
    w.bx = W / c; w.by = H / c; w.ncubes = (long)w.bx * w.by * (F / c); w.next = &next;
    jplan_t *plan = NULL; double *itab = NULL;
    if (!inverse) { plan = jplan_build(c); w.plan = plan; }
    else {
        const double scale = sqrt(8.0) / sqrt((double)cs);
        itab = (double *)malloc(sizeof(double) * cs * cs);
        for (int n = 0; n < cs; n++) for (int k = 0; k < cs; k++) {
            int n0 = n / (c * c), n1 = (n / c) % c, n2 = n % c, k0 = k / (c * c), k1 = (k / c) % c, k2 = k % c;
            itab[(size_t)n * cs + k] = scale * ck(k0) * ck(k1) * ck(k2) * basis_cos(c, n0, k0) * basis_cos(c, n1, k1) * basis_cos(c, n2, k2);
        }
        w.itab = itab;
    }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, java_worker, &w);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    if (plan) jplan_free(plan);
    free(itab);
}

void orc_java_dct_f64(const double *in, double *out, int W, int H, int F, int cube, int threads)
{ java_transform(in, out, W, H, F, cube, 0, threads); }
void orc_java_idct_f64(const double *in, double *out, int W, int H, int F, int cube, int threads)
{ java_transform(in, out, W, H, F, cube, 1, threads); }

/* Java-structured encode of u8 frames (J/Encoder.java:47-111): u8->double,    */
/* threaded DCT, single-threaded quantise (+reshuffle) and Exp-Golomb.         */
uint64_t orc_java_encode_u8(const uint8_t *frames, int W, int H, int F, int cube, int threads, uint8_t *buf, size_t cap)
{
    int Fe = F - F % cube;
    size_t n = (size_t)W * H * Fe;
    double *px = (double *)malloc(n * sizeof(double));
    double *cf = (double *)calloc(n, sizeof(double));
    int32_t *q = (int32_t *)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; i++) px[i] = (double)frames[i];
    orc_java_dct_f64(px, cf, W, H, Fe, cube, threads);
    orc_quantize_planar(cf, q, W, H, Fe, cube, 0);
    uint64_t bits = orc_eg_encode_cubes(q, n / (size_t)(cube * cube * cube), cube, buf, cap, 0);
    free(px); free(cf); free(q);
    return bits;
}

/* Java-structured decode (J/Decoder.java:61-117). */
int orc_java_decode_u8(const uint8_t *buf, size_t nbytes, int W, int H, int F, int cube, int threads, uint8_t *frames)
{
    int Fe = F - F % cube;
    size_t n = (size_t)W * H * Fe;
    int32_t *q = (int32_t *)malloc(n * sizeof(int32_t));
    uint64_t end = orc_eg_decode_cubes(buf, nbytes, 0, n / (size_t)(cube * cube * cube), cube, q);
    if (end == UINT64_MAX) { free(q); return -1; }
    double *cf = (double *)malloc(n * sizeof(double));
    double *px = (double *)calloc(n, sizeof(double));
    orc_dequantize_planar(q, cf, W, H, Fe, cube);
    orc_java_idct_f64(cf, px, W, H, Fe, cube, threads);
    for (size_t i = 0; i < n; i++) frames[i] = (uint8_t)px[i];
    free(q); free(cf); free(px);
    return 0;
}
