// eg_bits.h -- signed order-0 Exp-Golomb bit coding, one cube per thread, as
// __host__ __device__ inline functions (the device kernels and the host unit
// harness in tests/ compile the same code).
//
// Replaces, from the reference (/root/reference):
//   * ExpGolombWriter.writeValue / expGolomb_writeValue
//     (3d-DCT-video-encoding/src/br/jpiccoli/video/ExpGolombWriter.java:19-49,
//      3d-DCT-video-encoding-OpenCL/ExpGolomb.c:32-64) and the zig-zag driver loops
//     (Encoder.java:101-111, encoder.c:60-71);
//   * ExpGolombReader.readValue / expGolomb_readValue (ExpGolombReader.java:19-63,
//     ExpGolomb.c:66-110) and the scatter loops (Decoder.java:68-76, decoder.c:61-72).
//
// Code for v:  m = (v <= 0 ? -2v : 2v-1) + 1,  L = bitlength(m),  L-1 zero bits then m in
// L bits, MSB first.  v = 0 is the single bit '1', so a cube (97% zeros on natural content)
// is mostly runs of ones: both directions skip all-zero 16-coefficient chunks / runs of
// one-bits in O(1).
//
// Bit addressing: stream bit p lives in byte p>>3 at bit 7-(p&7).  Kernels move the stream
// as 32-bit words holding 4 stream bytes; bswap32() of such a word puts stream bit
// (32*w + i) at numeric bit 31-i.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EG_HD __host__ __device__ __forceinline__
#else
#define EG_HD inline
#endif

namespace dct3d {

EG_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
EG_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
EG_HD uint32_t bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// 16-byte vector: one LDS.128 / LDG.128 on the device.
struct alignas(16) Vec16 { uint32_t x, y, z, w; };
// the 16 int16 coefficients of chunk c of a zig-zag cube (32-byte aligned rows of 16 B)
EG_HD void load_chunk(const int16_t *zz, int c, uint32_t (&w)[8])
{
    const Vec16 *q = reinterpret_cast<const Vec16 *>(zz + 16 * c);
    const Vec16 a = q[0], b = q[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// v -> m (>= 1).  v > 0: 2v ; v <= 0: 1 - 2v.
EG_HD uint32_t eg_map(int v) { const int t = 2 * v; return (uint32_t)(v > 0 ? t : 1 - t); }
// m -> v
EG_HD int eg_unmap(uint32_t m) { return (m & 1u) ? -(int)(m >> 1) : (int)(m >> 1); }
// extra bits beyond the mandatory one: len - 1 = 2*(L-1)
EG_HD int eg_extra(int v) { return 2 * (31 - clz32(eg_map(v))); }

// ---------------------------------------------------------------------------------------
// Counting.  zz = the cube's coefficients in zig-zag order (int16), chunkmask bit i set iff
// coefficients [16i, 16i+16) contain a non-zero.  Returns the cube's code length in bits.
// ---------------------------------------------------------------------------------------
template <int CS>
EG_HD uint32_t eg_chunkmask(const int16_t *zz)
{
    uint32_t mask = 0;
    for (int c = 0; c < CS / 16; c++) {
        uint32_t w[8];
        load_chunk(zz, c, w);
        uint32_t any = 0;
        for (int i = 0; i < 8; i++) any |= w[i];
        if (any) mask |= 1u << c;
    }
    return mask;
}

template <int CS>
EG_HD uint32_t eg_count_cube(const int16_t *zz, uint32_t chunkmask)
{
    int extra = 0;
    for (uint32_t mk = chunkmask; mk; mk &= mk - 1) {
        const int c = 31 - clz32(mk & (0u - mk));
        uint32_t w[8];
        load_chunk(zz, c, w);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t p = w[i];
            extra += eg_extra((int)(int16_t)(p & 0xffffu));
            extra += eg_extra((int)(int16_t)(p >> 16));
        }
    }
    return (uint32_t)CS + (uint32_t)extra;
}

// ---------------------------------------------------------------------------------------
// Writing.  Sink is a policy with  void put(uint64_t word_index, uint32_t be_word, bool shared):
// `shared` words (the cube's first and last, which neighbouring cubes also touch) must be
// OR-merged; the others are plain stores.  Words are passed already byte-swapped for memory.
// ---------------------------------------------------------------------------------------
template <typename Sink>
struct BitWriter {
    Sink &sink;
    uint64_t acc;   // pending bits, left-aligned
    int nacc;       // number of pending bits (< 32 between calls)
    uint64_t widx;  // index of the word the pending bits start in
    bool first;

    EG_HD BitWriter(Sink &s, uint64_t start_bit) : sink(s), acc(0), nacc((int)(start_bit & 31)), widx(start_bit >> 5), first(true) {}

    // append the low `len` bits of m (1 <= len <= 33, nacc + len <= 64)
    EG_HD void put(uint64_t m, int len)
    {
        acc |= m << (64 - nacc - len);
        nacc += len;
        while (nacc >= 32) {  // at most twice (31 pending + a 33-bit code)
            sink.put(widx, bswap32((uint32_t)(acc >> 32)), first);
            first = false;
            widx++;
            acc <<= 32;
            nacc -= 32;
        }
    }
    EG_HD void put_code(int v)
    {
        const uint32_t m = eg_map(v);
        put(m, 2 * (32 - clz32(m)) - 1);
    }
    EG_HD void put_ones(int k)
    {
        while (k >= 32) { put(0xffffffffull, 32); k -= 32; }
        if (k > 0) put((1ull << k) - 1, k);
    }
    EG_HD void flush()
    {
        if (nacc > 0) sink.put(widx, bswap32((uint32_t)(acc >> 32)), true);
    }
};

template <int CS, typename Sink>
EG_HD void eg_write_cube(const int16_t *zz, uint32_t chunkmask, uint64_t start_bit, Sink &sink)
{
    BitWriter<Sink> bw(sink, start_bit);
    int pos = 0;
    for (uint32_t mk = chunkmask; mk; mk &= mk - 1) {
        const int c = 31 - clz32(mk & (0u - mk));
        bw.put_ones(16 * c - pos);
        uint32_t w[8];
        load_chunk(zz, c, w);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t p = w[i];
            bw.put_code((int)(int16_t)(p & 0xffffu));
            bw.put_code((int)(int16_t)(p >> 16));
        }
        pos = 16 * c + 16;
    }
    bw.put_ones(CS - pos);
    bw.flush();
}

// ---------------------------------------------------------------------------------------
// Reading.  Source is a policy with  uint32_t word(uint64_t word_index)  returning the
// byte-swapped (numeric MSB-first) word, zero past the end of the stream.
// ---------------------------------------------------------------------------------------
template <typename Source>
struct BitReader {
    const Source &src;
    uint64_t buf;    // next bits, left-aligned
    int navail;      // valid bits in buf
    uint64_t wnext;  // next word to load
    uint64_t pos;    // absolute position of the first bit in buf

    EG_HD BitReader(const Source &s, uint64_t start_bit) : src(s), pos(start_bit)
    {
        wnext = start_bit >> 5;
        const int sh = (int)(start_bit & 31);
        buf = (uint64_t)src.word(wnext++) << 32;
        buf |= (uint64_t)src.word(wnext++);
        buf <<= sh;
        navail = 64 - sh;
    }
    EG_HD void refill()
    {
        if (navail <= 32) {
            buf |= (uint64_t)src.word(wnext++) << (32 - navail);
            navail += 32;
        }
    }
    EG_HD void skip(int n) { buf = n >= 64 ? 0ull : buf << n; navail -= n; pos += (uint64_t)n; }
};

// Decode one cube (CS codes) starting at absolute bit `start`, scattering non-zero values to
// out[izz[i]] (out must be zero-filled).  Returns the end bit, or ~0 for a malformed code
// (more than 16 leading zeros: the value would not fit the codec's int16 range).
template <int CS, typename Source, typename Out>
EG_HD uint64_t eg_parse_cube(const Source &src, uint64_t start, const uint16_t *izz, Out &out)
{
    BitReader<Source> br(src, start);
    int i = 0;
    while (i < CS) {
        br.refill();
        const uint64_t inv = ~br.buf;
        int ones = inv ? clz64(inv) : 64;
        if (ones > br.navail) ones = br.navail;
        if (ones > 0) {
            if (ones > CS - i) ones = CS - i;
            i += ones;
            br.skip(ones);
            continue;
        }
        const int z = clz64(br.buf);  // leading bit is 0 here, so z >= 1
        if (z > 16) return ~0ull;
        const int len = 2 * z + 1;    // <= 33 <= navail
        const uint32_t m = (uint32_t)(br.buf >> (64 - len));
        out.put(izz[i], (int16_t)eg_unmap(m));
        i++;
        br.skip(len);
    }
    return br.pos;
}

// Count the codes that START in [start, limit) and report where the first code at or after
// `limit` starts (stream segments for index discovery).  Stops early at `maxcodes`.
// Returns false on a malformed code.
template <typename Source>
EG_HD bool eg_scan_segment(const Source &src, uint64_t start, uint64_t limit, uint64_t end_of_stream,
                           uint32_t &ncodes, uint64_t &next_start)
{
    BitReader<Source> br(src, start);
    uint32_t n = 0;
    while (br.pos < limit) {
        br.refill();
        const uint64_t inv = ~br.buf;
        int ones = inv ? clz64(inv) : 64;
        if (ones > br.navail) ones = br.navail;
        if (ones > 0) {
            const uint64_t room = limit - br.pos;
            if ((uint64_t)ones > room) ones = (int)room;
            n += (uint32_t)ones;
            br.skip(ones);
            continue;
        }
        const int z = clz64(br.buf);
        if (z > 16) {
            // zero padding after the last code of the stream is not an error
            if (br.pos + (uint64_t)z >= end_of_stream) { br.pos = limit > br.pos ? limit : br.pos; break; }
            return false;
        }
        n++;
        br.skip(2 * z + 1);
    }
    ncodes = n;
    next_start = br.pos;
    return true;
}

}  // namespace dct3d
