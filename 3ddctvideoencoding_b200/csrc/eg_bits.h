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

#ifndef DCT3D_SCAN_RUNLOOP
#define DCT3D_SCAN_RUNLOOP 0
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

// 32-bit funnel helpers (shift counts 0..32, clamped): one SHF each on the device.
EG_HD uint32_t fsl(uint32_t hi, uint32_t lo, int n)      // high word of (hi:lo) << n
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(lo, hi, n);
#else
    return n <= 0 ? hi : n >= 32 ? lo : (hi << n) | (lo >> (32 - n));
#endif
}
EG_HD uint32_t shl32c(uint32_t x, int n)                 // x << n, 0 when n == 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(0u, x, n);
#else
    return n >= 32 ? 0u : x << n;
#endif
}
EG_HD uint32_t shr32c(uint32_t x, int n)                 // x >> n, 0 when n == 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(x, 0u, n);
#else
    return n >= 32 ? 0u : x >> n;
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

// The largest code number an int16 coefficient maps to: v = -32768 -> m = 65537, the only 33-bit code.
// (m = 65536 would be v = +32768, every other 17-bit m lies beyond int16: such streams are malformed here;
// the reference reads them as full ints, ExpGolombReader.java:19-63.)
constexpr uint32_t kEgMaxCode = 65537u;

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

// DENSE: zz holds only the cube's non-zero chunks, one after the other in chunk order (the layout of the fused encoder's
// scratch); otherwise zz is the whole cube and chunk c sits at zz + 16 c.
template <int CS, bool DENSE = false>
EG_HD uint32_t eg_count_cube(const int16_t *zz, uint32_t chunkmask)
{
    int extra = 0, j = 0;
    for (uint32_t mk = chunkmask; mk; mk &= mk - 1, j++) {
        const int c = 31 - clz32(mk & (0u - mk));
        uint32_t w[8];
        load_chunk(zz, DENSE ? j : c, w);
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
    uint32_t acc;   // pending bits, left-aligned
    int nacc;       // number of pending bits (< 32 between calls)
    uint64_t widx;  // index of the word the pending bits belong to
    bool first;

    EG_HD BitWriter(Sink &s, uint64_t start_bit) : sink(s), acc(0), nacc((int)(start_bit & 31)), widx(start_bit >> 5), first(true) {}

    // append the low `len` bits of m, 1 <= len <= 32
    EG_HD void put(uint32_t m, int len)
    {
        const uint32_t ml = m << (32 - len);
        acc |= ml >> nacc;
        const uint32_t spill = shl32c(ml, 32 - nacc);   // bits that do not fit the current word
        nacc += len;
        if (nacc >= 32) {
            sink.put(widx, bswap32(acc), first);
            first = false;
            widx++;
            acc = spill;
            nacc -= 32;
        }
    }
    EG_HD void put_code(int v)
    {
        const uint32_t m = eg_map(v);
        const int L = 32 - clz32(m);
        if (L <= 16) {
            put(m, 2 * L - 1);
        } else {              // 33-bit code (v = -32768): 16 zeros, then m in 17 bits
            put(0u, 16);
            put(m, 17);
        }
    }
    EG_HD void put_ones(int k)
    {
        while (k >= 32) { put(0xffffffffu, 32); k -= 32; }
        if (k > 0) put(0xffffffffu >> (32 - k), k);
    }
    EG_HD void flush()
    {
        if (nacc > 0) sink.put(widx, bswap32(acc), true);
    }
};

template <int CS, typename Sink, bool DENSE = false>
EG_HD void eg_write_cube(const int16_t *zz, uint32_t chunkmask, uint64_t start_bit, Sink &sink)
{
    BitWriter<Sink> bw(sink, start_bit);
    int pos = 0, j = 0;
    for (uint32_t mk = chunkmask; mk; mk &= mk - 1, j++) {
        const int c = 31 - clz32(mk & (0u - mk));
        bw.put_ones(16 * c - pos);
        uint32_t w[8];
        load_chunk(zz, DENSE ? j : c, w);
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
// Reading.  Source is a policy with  uint32_t word(uint32_t j)  returning word j (relative to the
// source's own base) byte-swapped to numeric MSB-first order, zero past the end of the stream.
// Positions are 32-bit bit offsets relative to that base.
// ---------------------------------------------------------------------------------------
template <typename Source>
struct BitReader {
    const Source &src;
    uint32_t hi, lo;  // next bits, left-aligned in hi:lo
    int navail;       // valid bits in hi:lo (>= 33 after refill)
    uint32_t wnext;   // next word to load
    uint32_t pos;     // position of the first bit in hi

    EG_HD BitReader(const Source &s, uint32_t start_bit) : src(s), pos(start_bit)
    {
        wnext = start_bit >> 5;
        const int sh = (int)(start_bit & 31);
        const uint32_t w0 = src.word(wnext++), w1 = src.word(wnext++);
        hi = fsl(w0, w1, sh);
        lo = w1 << sh;
        navail = 64 - sh;
    }
    EG_HD void refill()                   // afterwards the 32 bits of hi are all valid
    {
        if (navail <= 32) {               // lo is empty here
            const uint32_t w = src.word(wnext++);
            hi |= shr32c(w, navail);
            lo = shl32c(w, 32 - navail);
            navail += 32;
        }
    }
    EG_HD void skip(int n)                // 0 <= n <= 32
    {
        hi = fsl(hi, lo, n);
        lo = shl32c(lo, n);
        navail -= n;
        pos += (uint32_t)n;
    }
    // Decode the code at the head (its first bit is 0).  Returns false for more than 16 leading zeros.
    EG_HD bool take_code(uint32_t &m)
    {
        const int z = clz32(hi);
        if (z > 16) return false;
        if (z < 16) {
            const int len = 2 * z + 1;
            m = hi >> (32 - len);
            skip(len);
        } else {                          // 33 bits: 16 zeros, a one, 16 more bits
            refill();                     // navail may be exactly 32 here
            m = (hi << 1) | (lo >> 31);
            if (m != kEgMaxCode) return false;   // the only 17-bit code number an int16 cube can hold (v = -32768)
            skip(32);
            skip(1);
        }
        return true;
    }
};

// Decode one cube (CS codes) starting at bit `start`, scattering non-zero values to
// out[izz[i]] (out must be zero-filled).  Returns the end bit, or ~0u for a malformed code
// (more than 16 leading zeros: the value would not fit the codec's int16 range).
template <int CS, typename Source, typename Out>
EG_HD uint32_t eg_parse_cube(const Source &src, uint32_t start, const uint16_t *izz, Out &out)
{
    BitReader<Source> br(src, start);
    int i = 0;
    while (i < CS) {
        // one iteration = a run of one-bits (zero coefficients) followed by one longer code
        br.refill();
        int ones = clz32(~br.hi);
        if (ones > CS - i) ones = CS - i;
        i += ones;
        br.skip(ones);
        if (i >= CS) break;
        br.refill();
        if (br.hi >> 31) continue;        // the run of ones goes on
        uint32_t m;
        if (!br.take_code(m)) return ~0u;
        out.put(izz[i], (int16_t)eg_unmap(m));
        i++;
    }
    return br.pos;
}

// Sink for the non-zero codes a segment scan meets: push(i, e) receives the i-th non-zero code of the
// segment as e = (code index relative to the segment's first code) << 17 | m, with m the 17-bit code
// number (value = eg_unmap(m)); the hot loop of the scan leaves the sign mapping to the reader.
struct NullNzSink {
    EG_HD void push(uint32_t, uint32_t) {}
    EG_HD void flush(uint32_t) {}
};

// Count the codes that START in [start, limit) and report where the first code at or after
// `limit` starts (stream segments for index discovery).  Returns false on a malformed code
// (zero padding after the last code of the stream is not an error).  With STOP the scan ends after
// stop_after codes instead (next_start = first bit after them).
// The 32 stream bits that start at bit `pos` (relative to the source's base), MSB first.  The scan keeps
// no window state: every look at the stream is two word reads and one funnel shift.
template <typename Source>
EG_HD uint32_t eg_fetch32(const Source &src, uint32_t pos)
{
    const uint32_t i = pos >> 5;
    return fsl(src.word(i), src.word(i + 1), (int)(pos & 31u));
}

template <typename Source, typename NzSink = NullNzSink, bool STOP = false>
EG_HD bool eg_scan_segment(const Source &src, uint32_t start, uint32_t limit, uint32_t end_of_stream,
                           uint32_t &ncodes, uint32_t &next_start, uint32_t *nonzero = nullptr, NzSink sink = NzSink(),
                           uint32_t stop_after = 0)
{
    uint32_t pos = start, n = 0, nz = 0;
    while (pos < limit) {
        // one iteration = a run of one-bits (zero coefficients) + one longer code
        // the run is measured on 64 bits at once: most runs of zero coefficients end inside them
        const uint32_t wi = pos >> 5;
        const int wb = (int)(pos & 31u);
        const uint32_t a1 = src.word(wi + 1);
        uint32_t w = fsl(src.word(wi), a1, wb);
        uint32_t ones = (uint32_t)clz32(~w);
        const uint32_t room = limit - pos;
#if DCT3D_SCAN_RUNLOOP
        // a long run (the zero tail of a cube is some 400 one-bits) is walked a word at a time: one new word, one funnel
        // shift and one count per 32 bits instead of a whole iteration of the code loop per 64
        if (ones == 32u) {
            uint32_t hi = a1, j = wi + 2;
            for (;;) {
                const uint32_t lo = src.word(j++);
                const uint32_t c = (uint32_t)clz32(~fsl(hi, lo, wb));
                ones += c;
                if (c < 32u || ones >= room) break;
                hi = lo;
            }
        }
#else
        if (ones == 32u) ones += (uint32_t)clz32(~fsl(a1, src.word(wi + 2), wb));
#endif
        if (ones > room) ones = room;
        if (STOP && ones > stop_after - n) ones = stop_after - n;
        n += ones;
        pos += ones;
        if (pos >= limit || (STOP && n >= stop_after)) break;
        w = eg_fetch32(src, pos);
        if (w >> 31) continue;                                  // the run goes on
        const int z = clz32(w);
        if (z > 16) {                                           // no such code: zero padding at the end of the stream, or damage
            if (pos + 17 >= end_of_stream || (uint32_t)z + pos >= end_of_stream) { if (pos < limit) pos = limit; break; }
            return false;
        }
        // a code that runs past the end of the stream is not a code yet: the caller sees fewer codes than it
        // needs (DCT3D_E_NEED_MORE / truncated) instead of a value read from the zero padding
        const uint32_t len = z < 16 ? 2u * (uint32_t)z + 1u : 33u;
        if (pos + len > end_of_stream) { pos = limit; break; }
        uint32_t m;
        if (z < 16) {
            m = w >> (32 - len);
        } else {                                                // 33 bits: 16 zeros, then m in 17 bits
            m = eg_fetch32(src, pos + 16) >> 15;
            if (m != kEgMaxCode) return false;                  // would not fit the codec's int16 cubes
        }
        pos += len;
        sink.push(nz, (n << 17) | m);
        n++;
        nz++;
        if (STOP && n >= stop_after) break;
    }
    sink.flush(nz);
    ncodes = n;
    next_start = pos;
    if (nonzero) *nonzero = nz;
    return true;
}

}  // namespace dct3d
