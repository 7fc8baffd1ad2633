// dct_math.h -- separable orthonormal DCT-II / DCT-III butterflies and the codec's
// quantiser, as __host__ __device__ inline functions.
//
// What this replaces in the reference (paths relative to /root/reference):
//   * the O(512^2)-per-cube transforms of 3d-DCT-video-encoding-OpenCL/3dDCT.cl:43-143
//     (forward) and :164-265 (inverse), and Java's DCT.apply / InverseDCT.apply
//     (3d-DCT-video-encoding/src/br/jpiccoli/video/dct/DCT.java:41-59,
//     InverseDCT.java:33-82).  The 3D transform
//       X[k0,k1,k2] = s c(k0)c(k1)c(k2) sum x cos cos cos,  s = sqrt(8/(N^3)), c(0)=1/sqrt2
//     (DCT.java:81-85,112) factors per axis into sqrt(2/N) c(k) cos(pi/N (n+1/2) k),
//     i.e. the orthonormal DCT-II, applied along x, y and t.
//   * the quantisation function max(1, 5*(k0+k1+k2)) (Encoder.java:82, encoder.c:53;
//     inverse Decoder.java:89, decoder.c:54).
//
// Each 8-point transform is an even/odd butterfly: 36 FADD/FMUL/FFMA with
// compile-time constants (immediates in SASS) instead of 64 MACs.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DCT_HD __host__ __device__ __forceinline__
#else
#define DCT_HD inline
#endif

namespace dct3d {

template <typename T> DCT_HD T fma_t(T a, T b, T c);
template <> DCT_HD float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> DCT_HD double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// cos(k*pi/16) / 2 for k = 1..7 (the a_k = 1/2 normalisation folded in), and 1/sqrt(8).
#define DCT_H1 0.49039264020161522456
#define DCT_H2 0.46193976625564337806
#define DCT_H3 0.41573480615127261854
#define DCT_H5 0.27778511650980111237
#define DCT_H6 0.19134171618254488586
#define DCT_H7 0.09754516100806413392
#define DCT_R8 0.35355339059327376220
// 4-point: cos(pi/8)/sqrt2, cos(3pi/8)/sqrt2
#define DCT_G1 0.65328148243818826393
#define DCT_G3 0.27059805007309849220

// Forward orthonormal 8-point DCT-II, in place.
template <typename T>
DCT_HD void dct8_fwd(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7)
{
    const T s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
    const T d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
    const T e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
    const T r = (T)DCT_R8, a = (T)DCT_H2, b = (T)DCT_H6;
    const T h1 = (T)DCT_H1, h3 = (T)DCT_H3, h5 = (T)DCT_H5, h7 = (T)DCT_H7;
    const T t0 = e0 * r;
    x0 = fma_t<T>(e1, r, t0);
    x4 = fma_t<T>(e1, -r, t0);
    x2 = fma_t<T>(e3, b, e2 * a);
    x6 = fma_t<T>(e3, -a, e2 * b);
    x1 = fma_t<T>(d3, h7, fma_t<T>(d2, h5, fma_t<T>(d1, h3, d0 * h1)));
    x3 = fma_t<T>(d3, -h5, fma_t<T>(d2, -h1, fma_t<T>(d1, -h7, d0 * h3)));
    x5 = fma_t<T>(d3, h3, fma_t<T>(d2, h7, fma_t<T>(d1, -h1, d0 * h5)));
    x7 = fma_t<T>(d3, -h1, fma_t<T>(d2, h3, fma_t<T>(d1, -h5, d0 * h7)));
}

// Inverse (orthonormal DCT-III), in place: X0..X7 -> x0..x7.
template <typename T>
DCT_HD void dct8_inv(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7)
{
    const T r = (T)DCT_R8, a = (T)DCT_H2, b = (T)DCT_H6;
    const T h1 = (T)DCT_H1, h3 = (T)DCT_H3, h5 = (T)DCT_H5, h7 = (T)DCT_H7;
    const T t0 = x0 * r;
    const T u0 = fma_t<T>(x4, r, t0), u1 = fma_t<T>(x4, -r, t0);
    const T u2 = fma_t<T>(x6, b, x2 * a), u3 = fma_t<T>(x6, -a, x2 * b);
    const T e0 = u0 + u2, e3 = u0 - u2, e1 = u1 + u3, e2 = u1 - u3;
    const T o0 = fma_t<T>(x7, h7, fma_t<T>(x5, h5, fma_t<T>(x3, h3, x1 * h1)));
    const T o1 = fma_t<T>(x7, -h5, fma_t<T>(x5, -h1, fma_t<T>(x3, -h7, x1 * h3)));
    const T o2 = fma_t<T>(x7, h3, fma_t<T>(x5, h7, fma_t<T>(x3, -h1, x1 * h5)));
    const T o3 = fma_t<T>(x7, -h1, fma_t<T>(x5, h3, fma_t<T>(x3, -h5, x1 * h7)));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

// ---------------------------------------------------------------------------------------------
// Scaled variants.  The 3D transform is a product of three 1D transforms, so a per-frequency scale
// S[k] of one axis can be applied anywhere later.  The kernels use
//   dct8_fwd_n : X'_k = X_k / S[k]   (29 instead of 36 instructions), S = {r, h1, h2, h3, r, h5, h2, h7}
//   dct8_inv_n : takes P_k = S[k] X_k (31 instead of 36)
// along x and y, and fold the missing factors into (a) the constants of the t-axis transform
// (dct8_fwd_g / dct8_inv_g: all constants times a compile-time g = S[k2] of the column) and
// (b) the per-lane quantiser / dequantiser table (S[k1], k1 = lane).
// ---------------------------------------------------------------------------------------------
template <typename T> DCT_HD constexpr T dct8_scale(int k)
{
    return k == 0 || k == 4 ? (T)DCT_R8 : k == 1 ? (T)DCT_H1 : k == 2 || k == 6 ? (T)DCT_H2 : k == 3 ? (T)DCT_H3 : k == 5 ? (T)DCT_H5 : (T)DCT_H7;
}

template <typename T>
DCT_HD void dct8_fwd_n(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7)
{
    const T s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
    const T d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
    const T e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
    const T ba = (T)(DCT_H6 / DCT_H2);
    x0 = e0 + e1;
    x4 = e0 - e1;
    x2 = fma_t<T>(e3, ba, e2);
    x6 = fma_t<T>(e2, ba, -e3);
    x1 = fma_t<T>(d3, (T)(DCT_H7 / DCT_H1), fma_t<T>(d2, (T)(DCT_H5 / DCT_H1), fma_t<T>(d1, (T)(DCT_H3 / DCT_H1), d0)));
    x3 = fma_t<T>(d3, (T)(-DCT_H5 / DCT_H3), fma_t<T>(d2, (T)(-DCT_H1 / DCT_H3), fma_t<T>(d1, (T)(-DCT_H7 / DCT_H3), d0)));
    x5 = fma_t<T>(d3, (T)(DCT_H3 / DCT_H5), fma_t<T>(d2, (T)(DCT_H7 / DCT_H5), fma_t<T>(d1, (T)(-DCT_H1 / DCT_H5), d0)));
    x7 = fma_t<T>(d3, (T)(-DCT_H1 / DCT_H7), fma_t<T>(d2, (T)(DCT_H3 / DCT_H7), fma_t<T>(d1, (T)(-DCT_H5 / DCT_H7), d0)));
}

template <typename T>
DCT_HD void dct8_inv_n(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7)
{
    const T ba = (T)(DCT_H6 / DCT_H2);
    const T u0 = x0 + x4, u1 = x0 - x4;
    const T u2 = fma_t<T>(x6, ba, x2), u3 = fma_t<T>(x2, ba, -x6);
    const T e0 = u0 + u2, e3 = u0 - u2, e1 = u1 + u3, e2 = u1 - u3;
    const T o0 = (x1 + x3) + (x5 + x7);
    const T o1 = fma_t<T>(x7, (T)(-DCT_H5 / DCT_H7), fma_t<T>(x5, (T)(-DCT_H1 / DCT_H5), fma_t<T>(x3, (T)(-DCT_H7 / DCT_H3), x1 * (T)(DCT_H3 / DCT_H1))));
    const T o2 = fma_t<T>(x7, (T)(DCT_H3 / DCT_H7), fma_t<T>(x5, (T)(DCT_H7 / DCT_H5), fma_t<T>(x3, (T)(-DCT_H1 / DCT_H3), x1 * (T)(DCT_H5 / DCT_H1))));
    const T o3 = fma_t<T>(x7, (T)(-DCT_H1 / DCT_H7), fma_t<T>(x5, (T)(DCT_H3 / DCT_H5), fma_t<T>(x3, (T)(-DCT_H5 / DCT_H3), x1 * (T)(DCT_H7 / DCT_H1))));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

// true-constant transforms with every constant multiplied by g (a compile-time value at the call site)
template <typename T>
DCT_HD void dct8_fwd_g(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7, const T g)
{
    const T s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
    const T d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
    const T e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
    const T r = (T)DCT_R8 * g, a = (T)DCT_H2 * g, b = (T)DCT_H6 * g;
    const T h1 = (T)DCT_H1 * g, h3 = (T)DCT_H3 * g, h5 = (T)DCT_H5 * g, h7 = (T)DCT_H7 * g;
    const T t0 = e0 * r;
    x0 = fma_t<T>(e1, r, t0);
    x4 = fma_t<T>(e1, -r, t0);
    x2 = fma_t<T>(e3, b, e2 * a);
    x6 = fma_t<T>(e3, -a, e2 * b);
    x1 = fma_t<T>(d3, h7, fma_t<T>(d2, h5, fma_t<T>(d1, h3, d0 * h1)));
    x3 = fma_t<T>(d3, -h5, fma_t<T>(d2, -h1, fma_t<T>(d1, -h7, d0 * h3)));
    x5 = fma_t<T>(d3, h3, fma_t<T>(d2, h7, fma_t<T>(d1, -h1, d0 * h5)));
    x7 = fma_t<T>(d3, -h1, fma_t<T>(d2, h3, fma_t<T>(d1, -h5, d0 * h7)));
}

template <typename T>
DCT_HD void dct8_inv_g(T &x0, T &x1, T &x2, T &x3, T &x4, T &x5, T &x6, T &x7, const T g)
{
    const T r = (T)DCT_R8 * g, a = (T)DCT_H2 * g, b = (T)DCT_H6 * g;
    const T h1 = (T)DCT_H1 * g, h3 = (T)DCT_H3 * g, h5 = (T)DCT_H5 * g, h7 = (T)DCT_H7 * g;
    const T t0 = x0 * r;
    const T u0 = fma_t<T>(x4, r, t0), u1 = fma_t<T>(x4, -r, t0);
    const T u2 = fma_t<T>(x6, b, x2 * a), u3 = fma_t<T>(x6, -a, x2 * b);
    const T e0 = u0 + u2, e3 = u0 - u2, e1 = u1 + u3, e2 = u1 - u3;
    const T o0 = fma_t<T>(x7, h7, fma_t<T>(x5, h5, fma_t<T>(x3, h3, x1 * h1)));
    const T o1 = fma_t<T>(x7, -h5, fma_t<T>(x5, -h1, fma_t<T>(x3, -h7, x1 * h3)));
    const T o2 = fma_t<T>(x7, h3, fma_t<T>(x5, h7, fma_t<T>(x3, -h1, x1 * h5)));
    const T o3 = fma_t<T>(x7, -h1, fma_t<T>(x5, h3, fma_t<T>(x3, -h5, x1 * h7)));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

// 4-point orthonormal DCT-II / DCT-III (a_0 = 1/2, a_k = 1/sqrt2).
template <typename T>
DCT_HD void dct4_fwd(T &x0, T &x1, T &x2, T &x3)
{
    const T s0 = x0 + x3, s1 = x1 + x2, d0 = x0 - x3, d1 = x1 - x2;
    const T g1 = (T)DCT_G1, g3 = (T)DCT_G3, h = (T)0.5;
    const T t0 = s0 * h;
    x0 = fma_t<T>(s1, h, t0);
    x2 = fma_t<T>(s1, -h, t0);
    x1 = fma_t<T>(d1, g3, d0 * g1);
    x3 = fma_t<T>(d1, -g1, d0 * g3);
}

template <typename T>
DCT_HD void dct4_inv(T &x0, T &x1, T &x2, T &x3)
{
    const T g1 = (T)DCT_G1, g3 = (T)DCT_G3, h = (T)0.5;
    const T t0 = x0 * h;
    const T e0 = fma_t<T>(x2, h, t0), e1 = fma_t<T>(x2, -h, t0);
    const T o0 = fma_t<T>(x3, g3, x1 * g1), o1 = fma_t<T>(x3, -g1, x1 * g3);
    x0 = e0 + o0; x3 = e0 - o0;
    x1 = e1 + o1; x2 = e1 - o1;
}

// N-point dispatch on a strided register array (all indices compile-time once unrolled).
template <int N, typename T> struct Dct1D;
template <typename T> struct Dct1D<8, T> {
    template <int S> static DCT_HD void fwd(T *v) { dct8_fwd<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S]); }
    template <int S> static DCT_HD void inv(T *v) { dct8_inv<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S]); }
    // scaled variants (see above); scale(k) is what fwd_n leaves out / inv_n expects
    template <int S> static DCT_HD void fwd_n(T *v) { dct8_fwd_n<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S]); }
    template <int S> static DCT_HD void inv_n(T *v) { dct8_inv_n<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S]); }
    template <int S> static DCT_HD void fwd_g(T *v, const T g) { dct8_fwd_g<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S], g); }
    template <int S> static DCT_HD void inv_g(T *v, const T g) { dct8_inv_g<T>(v[0], v[S], v[2 * S], v[3 * S], v[4 * S], v[5 * S], v[6 * S], v[7 * S], g); }
    static DCT_HD constexpr T scale(int k) { return dct8_scale<T>(k); }
};
template <typename T> struct Dct1D<4, T> {
    template <int S> static DCT_HD void fwd(T *v) { dct4_fwd<T>(v[0], v[S], v[2 * S], v[3 * S]); }
    template <int S> static DCT_HD void inv(T *v) { dct4_inv<T>(v[0], v[S], v[2 * S], v[3 * S]); }
    // the 4-point butterflies are already minimal: the scaled interface maps to the plain one
    template <int S> static DCT_HD void fwd_n(T *v) { fwd<S>(v); }
    template <int S> static DCT_HD void inv_n(T *v) { inv<S>(v); }
    template <int S> static DCT_HD void fwd_g(T *v, const T) { fwd<S>(v); }
    template <int S> static DCT_HD void inv_g(T *v, const T) { inv<S>(v); }
    static DCT_HD constexpr T scale(int) { return (T)1; }
};

// Quantiser divisor max(1, 5*(k0+k1+k2)).
DCT_HD int quant_divisor(int ksum) { return ksum == 0 ? 1 : 5 * ksum; }

// Round-to-nearest of v = coef * recip via the 1.5*2^23 magic constant: one FFMA, the
// integer lands in the low mantissa bits.  Exact ties (measure zero for this transform, see
// DESIGN.md) round to even; the reference rounds them up (Java) or away from zero (C).
#define DCT_MAGIC 12582912.0f
DCT_HD int quantize_f32(float coef, float recip)
{
    const float r = fmaf(coef, recip, DCT_MAGIC);
    union { float f; int32_t i; } u;
    u.f = r;
    return u.i - 0x4B400000;
}

// The largest float t with t * recip <= 0.5 in exact arithmetic:  |x| <= t  <=>  fmaf(x, recip, DCT_MAGIC) == DCT_MAGIC,
// i.e. x quantises to zero (the FFMA rounds the exact sum once; the tie |x recip| = 0.5 goes to the even neighbour, which is
// DCT_MAGIC itself from both sides).  The quotient 0.5 / recip is rounded to nearest, so the bound is it or its predecessor;
// the product of two floats is exact in double.  Used by the fused encoder's zero-run test.
DCT_HD float zero_threshold(float recip)
{
    const float t = 0.5f / recip;
    union { float f; int32_t i; } u;
    u.f = t;
    if ((double)t * (double)recip > 0.5) u.i -= 1;
    return u.f;
}

}  // namespace dct3d
