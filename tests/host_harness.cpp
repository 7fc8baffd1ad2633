// host_harness.cpp -- TEST-ONLY: compiles the kernels' __host__ __device__ arithmetic
// (csrc/dct_math.h, csrc/eg_bits.h) with the host compiler so the butterflies, the quantiser
// and the per-thread Exp-Golomb writer/parser can be checked against the oracle in the
// CPU-only test tier.  It is NOT a CPU fallback: nothing in the package or in libdct3d.so
// links or loads it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../3ddctvideoencoding_b200/csrc/dct_math.h"
#include "../3ddctvideoencoding_b200/csrc/eg_bits.h"

using namespace dct3d;

template <int N, typename T>
static void cube_transform(const T *in, T *out, bool inverse)
{
    T a[N * N * N];
    memcpy(a, in, sizeof(a));
    // same axis order as the kernels: forward x, y, t ; inverse t, y, x.  a is [t][y][x].
    auto ax = [&](int which) {
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
            T v[N];
            for (int k = 0; k < N; k++) {
                int idx = which == 0 ? (i * N + j) * N + k : which == 1 ? (i * N + k) * N + j : (k * N + i) * N + j;
                v[k] = a[idx];
            }
            if (inverse) Dct1D<N, T>::template inv<1>(v); else Dct1D<N, T>::template fwd<1>(v);
            for (int k = 0; k < N; k++) {
                int idx = which == 0 ? (i * N + j) * N + k : which == 1 ? (i * N + k) * N + j : (k * N + i) * N + j;
                a[idx] = v[k];
            }
        }
    };
    if (!inverse) { ax(0); ax(1); ax(2); } else { ax(2); ax(1); ax(0); }
    memcpy(out, a, sizeof(a));
}

// the scaled pipeline the fused kernels use: x and y normalised, t with constants times S[k2],
// S[k1] applied at the end (the kernels fold it into the quantiser table)
template <int N, typename T>
static void cube_transform_scaled(const T *in, T *out, bool inverse)
{
    T a[N][N][N];   // [t][y][x]
    memcpy(a, in, sizeof(a));
    if (!inverse) {
        for (int t = 0; t < N; t++) for (int y = 0; y < N; y++) Dct1D<N, T>::template fwd_n<1>(&a[t][y][0]);
        for (int t = 0; t < N; t++) for (int x = 0; x < N; x++) Dct1D<N, T>::template fwd_n<N>(&a[t][0][x]);
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) Dct1D<N, T>::template fwd_g<N * N>(&a[0][y][x], Dct1D<N, T>::scale(x));
        for (int t = 0; t < N; t++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) a[t][y][x] *= Dct1D<N, T>::scale(y);
    } else {
        for (int t = 0; t < N; t++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) a[t][y][x] *= Dct1D<N, T>::scale(y);
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) Dct1D<N, T>::template inv_g<N * N>(&a[0][y][x], Dct1D<N, T>::scale(x));
        for (int t = 0; t < N; t++) for (int x = 0; x < N; x++) Dct1D<N, T>::template inv_n<N>(&a[t][0][x]);
        for (int t = 0; t < N; t++) for (int y = 0; y < N; y++) Dct1D<N, T>::template inv_n<1>(&a[t][y][0]);
    }
    memcpy(out, a, sizeof(a));
}

struct VecSink {
    std::vector<uint32_t> &w;
    explicit VecSink(std::vector<uint32_t> &v) : w(v) {}
    void put(uint64_t idx, uint32_t be, bool shared)
    {
        if (idx >= w.size()) w.resize(idx + 1, 0);
        if (shared) w[idx] |= be; else w[idx] = be;
    }
};
struct VecSource {
    const uint32_t *w; uint64_t n;
    uint32_t word(uint32_t i) const { return i < n ? bswap32(w[i]) : 0u; }
};
struct ArrOut { int16_t *o; void put(int idx, int16_t v) { o[idx] = v; } };

extern "C" {

void hh_cube_f32(const float *in, float *out, int n, int inverse)
{ if (n == 8) cube_transform<8, float>(in, out, inverse); else cube_transform<4, float>(in, out, inverse); }
void hh_cube_f64(const double *in, double *out, int n, int inverse)
{ if (n == 8) cube_transform<8, double>(in, out, inverse); else cube_transform<4, double>(in, out, inverse); }
void hh_cube_scaled_f32(const float *in, float *out, int n, int inverse)
{ if (n == 8) cube_transform_scaled<8, float>(in, out, inverse); else cube_transform_scaled<4, float>(in, out, inverse); }
void hh_cube_scaled_f64(const double *in, double *out, int n, int inverse)
{ if (n == 8) cube_transform_scaled<8, double>(in, out, inverse); else cube_transform_scaled<4, double>(in, out, inverse); }
int hh_quantize(float coef, int ksum) { return quantize_f32(coef, 1.0f / (float)quant_divisor(ksum)); }
// the fused encoder's zero-run test: threshold of a reciprocal, and the quantiser's own verdict for a coefficient
float hh_zero_threshold(float recip) { return zero_threshold(recip); }
int hh_quantize_recip(float coef, float recip) { return quantize_f32(coef, recip); }

// zz: ncubes x cs int16 in zig-zag order.  Writes the stream into out (cap bytes, zeroed by the
// caller) starting at start_bit; returns the end bit.  Uses count + write exactly as the kernel.
uint64_t hh_eg_write(const int16_t *zz, int ncubes, int cs, uint64_t start_bit, uint8_t *out, size_t cap)
{
    std::vector<uint32_t> words;
    VecSink sink(words);
    uint64_t pos = start_bit;
    for (int c = 0; c < ncubes; c++) {
        const int16_t *p = zz + (size_t)c * cs;
        uint32_t mask = cs == 512 ? eg_chunkmask<512>(p) : eg_chunkmask<64>(p);
        uint32_t nb = cs == 512 ? eg_count_cube<512>(p, mask) : eg_count_cube<64>(p, mask);
        if (cs == 512) eg_write_cube<512>(p, mask, pos, sink); else eg_write_cube<64>(p, mask, pos, sink);
        pos += nb;
    }
    size_t nbytes = words.size() * 4;
    if (nbytes > cap) nbytes = cap;
    for (size_t i = 0; i < nbytes; i++) out[i] |= reinterpret_cast<const uint8_t *>(words.data())[i];
    return pos;
}

// Parse ncubes cubes from bit `start`; out = ncubes x cs int16 (natural order via izz), zeroed by caller.
uint64_t hh_eg_parse(const uint8_t *buf, size_t nbytes, uint64_t start, int ncubes, int cs, const uint16_t *izz, int16_t *out)
{
    std::vector<uint32_t> words((nbytes + 3) / 4 + 1, 0);
    memcpy(words.data(), buf, nbytes);
    VecSource src{words.data(), (uint64_t)words.size()};
    uint32_t pos = (uint32_t)start;
    for (int c = 0; c < ncubes; c++) {
        ArrOut o{out + (size_t)c * cs};
        pos = cs == 512 ? eg_parse_cube<512>(src, pos, izz, o) : eg_parse_cube<64>(src, pos, izz, o);
        if (pos == ~0u) return ~0ull;
    }
    return pos;
}

int hh_eg_scan(const uint8_t *buf, size_t nbytes, uint64_t start, uint64_t limit, uint32_t *ncodes, uint64_t *next)
{
    std::vector<uint32_t> words((nbytes + 3) / 4 + 1, 0);
    memcpy(words.data(), buf, nbytes);
    VecSource src{words.data(), (uint64_t)words.size()};
    uint32_t nx = 0;
    const bool ok = eg_scan_segment(src, (uint32_t)start, (uint32_t)limit, (uint32_t)(nbytes * 8), *ncodes, nx);
    *next = nx;
    return ok ? 0 : -1;
}
}
