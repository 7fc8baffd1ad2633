// dct3d_kernels.cuh -- CUDA kernels (sm_100a) for the 3D-DCT codec hot path.
//
// Design (see DESIGN.md for the full account):
//   * encode_kernel: a warp transforms a "unit" of 32 px x C rows x C frames (32/C cubes) per pass.
//     The unit arrives by ONE TMA box {32, C frames, C rows} (SWIZZLE_32B) straight from the planar
//     frame stack into the warp's private double buffer, completing on the warp's own mbarrier: no
//     CTA-wide barrier in the loop.  This is what replaces the host-side u8 -> float cube reshuffle
//     readCubes / writeCubes (reference 3d-DCT-video-encoding-OpenCL/encoder.c:10-45, decoder.c:10-46).
//   * Transform: C threads own one cube.  Thread t first holds the CxC plane of frame t in
//     registers and runs the 1D butterflies along x and y, the cube is then transposed through
//     a swizzled shared-memory exchange so that thread j holds all (k0, k2) for row-frequency
//     k1 = j, and the butterflies along t follow.  ~12 flop-instructions per sample (un-normalised
//     butterflies, scales folded into the t pass and the quantiser) vs 24 for the matrix form.
//   * The quantiser is one FFMA per coefficient (reciprocal table per lane + magic rounding).
//   * In that layout the elements a thread owns on one diagonal k0+k2 are CONTIGUOUS in the
//     reference's zig-zag order (CubeUtils.c:17-42: slices of constant x+y+z, y outer, z middle),
//     so the zig-zag reorder is a predicated STS.U16 per non-zero with per-lane base registers and
//     immediate offsets; the non-zero 16-coefficient chunks go to a sparsely touched scratch.  Groups of
//     high diagonals are first tested against the lane's exact zero threshold and skipped when the whole warp
//     has nothing there (97% of the coefficients are zero).
//   * eg_pack_kernel / eg_pack_sorted_kernel: one thread per cube: count pass, block scan, decoupled look-back
//     across tiles for the global bit offset, write pass straight to the global stream (plain stores inside a
//     cube, OR-merge for the two boundary words).  The sorted kernel deals the cubes of a tile to the threads in
//     order of their chunk counts, so that the lanes of a warp carry like loads.  The bitstream never visits the
//     host (reference: ExpGolomb.c:32-64 on one host thread).
//   * Decode: index discovery over 1024-bit segments (seg_scan_kernel with a lead-in walk that
//     guesses every entry point, keeping the non-zero codes it decodes; seg_fix_kernel; two prefix
//     sums), seg_emit_kernel turns the segment lists into CSR rows of the cubes, and
//     reconstruct_coo_kernel scatters the dequantised non-zeros into a float cube, runs the inverse
//     butterflies (skipping all-zero columns), clamps, truncates and hands the warp's pixel tile to TMA
//     (one cp.async.bulk.tensor store per unit: the encoder's box read backwards; row stores for widths that
//     are not a multiple of 32).
//   * transform_kernel: the f32 / f64 transform seams; rgb_planes_kernel: RGBUtils split / mix;
//     quant_f64_kernel and friends: the fp64 mode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "dct_math.h"
#include "eg_bits.h"

namespace dct3d {

constexpr int kUnitW = 32;       // pixels (bytes) per warp unit = TMA box row
constexpr int kWarps = 4;        // transform warps per CTA
constexpr int kThreads = kWarps * 32;

template <int C>
struct Geo {
    static constexpr int CS = C * C * C;
    static constexpr int CPW = 32 / C;                       // cubes per warp pass
    static constexpr int UNIT_BYTES = kUnitW * C * C;        // one TMA op: 32 px x C frames x C rows
    static constexpr int ZZ_STRIDE = CS + 8;                 // int16 units; odd multiple of 16 B
    static constexpr int NDIAG = 2 * C - 1;
    static constexpr int CHUNKS = CS / 16;                   // 16-coefficient chunks per cube
};

// Frame / cube geometry shared by host and device.
struct Layout {
    int W, H, C;
    int bx, by;          // cubes per row / cube rows
    int bxu;             // warp units (32 px = 32/C cubes) per cube row
    int nslabs;
    long long nunits, ncubes;
};

struct EncParams {
    Layout L;
    const uint8_t *frames;
    uint32_t *out_words;            // stream as 32-bit words (memory byte order)
    unsigned long long cap_bits;    // capacity of out_words in bits
    unsigned long long start_bit;
    const unsigned long long *start_bit_dev;   // when set, the start bit is read from device memory instead (written by the
                                               // previous call's packer: slab ranges chained without a host round trip)
    unsigned long long *tile_status;  // [ntiles], zeroed: flag<<62 | inclusive bit count
    unsigned int *ticket;             // zeroed
    unsigned int *err;                // zeroed; bit0 = overflow
    unsigned long long *end_bit;      // out: start_bit + total bits
    unsigned long long *end_bit_host; // optional mirror in mapped page-locked host memory: the host learns the end bit of a
                                      // pipeline chunk without a copy that would queue behind other D2H traffic
    int16_t *zzg;                     // zig-zag chunk scratch: a region of CPW * CS int16 per warp unit (CPW cubes side by
                                      // side in a cube row), holding only the unit's NON-ZERO 16-coefficient chunks, one after
                                      // the other in (cube, chunk) order: contiguous for the packer, no 64-byte DRAM over-fetch
    uint32_t *cmask;                  // [cube] mask of non-zero 16-coefficient chunks
    int16_t *qcubes;                  // MODE_NAT: natural-order int16 cubes out
    const int16_t *qcubes_in;         // EG-only kernel: natural-order int16 cubes in
    int use_tma;
    int debug;                        // reserved for profiling experiments
};

// ------------------------------------------------------------------------------------------
// zig-zag tables (filled by the host from the CubeUtils.diagonalSlices order)
// ------------------------------------------------------------------------------------------
struct ZzTables {
    uint16_t base8[8][15];   // zig-zag index of the first element of thread j's run on diagonal s'
    uint16_t base4[4][7];
    uint16_t lin8[512];      // zig-zag position -> natural index k2 + 8*k1 + 64*k0
    uint16_t lin4[64];
    uint16_t slin8[512];     // the same through coo_swizzle(): index space of the decoder's lists
    uint16_t slin4[64];
};
__constant__ ZzTables c_zz;
__device__ ZzTables g_zz;        // the same tables in global memory, for lookups with a per-thread index

template <int C> __device__ __forceinline__ uint16_t zz_base(int j, int s) { return C == 8 ? c_zz.base8[j][s] : c_zz.base4[j][s]; }
template <int C> __device__ __forceinline__ const uint16_t *zz_lin() { return C == 8 ? c_zz.lin8 : c_zz.lin4; }
template <int C> __device__ __forceinline__ const uint16_t *zz_slin() { return C == 8 ? c_zz.slin8 : c_zz.slin4; }

// Index space of the decoder's non-zero lists: the natural index with the two 4-float halves of a row
// exchanged for k1 >= 4 (C = 8), so that the 8 lanes k1 = 0..7 of a cube read their 16-byte half rows
// of the float cube in shared memory (row stride 32 B) from 8 distinct bank groups.  An involution.
template <int C> __host__ __device__ __forceinline__ constexpr uint32_t coo_swizzle(uint32_t idx)
{
    return C == 8 ? idx ^ ((idx >> 3) & 4u) : idx;
}

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// TMA tile store (shared -> global) of one {32 px, C frames, C rows} box, bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tmap, const void *src, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
        ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(src)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the issuing thread waits until its bulk stores have READ their shared-memory source (the tile may be rewritten)
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Unit geometry.  A unit is what one warp transforms per pass: 32 px x C rows x C frames =
// 32/C cubes side by side.  Shared-memory image of a unit: [y][t][32 px], which is what one TMA
// box {32, C frames, C rows} over the tensor {W, F, H} delivers.  With SWIZZLE_32B the two 16-byte
// halves of a 32-byte row are exchanged when bit 7 of the address is set (rows t = 4..7 of every y for
// C = 8), so the 16 lanes of a half warp (2 cubes x 8 frames, 8 bytes each) cover all 32 banks once.
// ------------------------------------------------------------------------------------------
struct UnitPos { int slab, byi, bxu; };

// byte offset of pixel xb (0..31) of row (y, t) inside a unit buffer (256-byte aligned)
template <int C>
__device__ __forceinline__ int unit_offset(int y, int t, int xb)
{
    const int row = y * C + t;
    return row * kUnitW + (xb ^ (((row >> 2) & 1) << 4));
}

__device__ __forceinline__ UnitPos unit_pos(const Layout &L, long long u)
{
    UnitPos p;
    const int per_slab = L.by * L.bxu;
    p.slab = (int)(u / per_slab);
    const int rem = (int)(u - (long long)p.slab * per_slab);
    p.byi = rem / L.bxu;
    p.bxu = rem - p.byi * L.bxu;
    return p;
}

// advance by (ds, dy, dx) = the decomposition of the grid stride, with carries
__device__ __forceinline__ void unit_advance(const Layout &L, UnitPos &p, const UnitPos &d)
{
    p.bxu += d.bxu;
    p.byi += d.byi;
    p.slab += d.slab;
    if (p.bxu >= L.bxu) { p.bxu -= L.bxu; p.byi++; }
    if (p.byi >= L.by) { p.byi -= L.by; p.slab++; }
}

// ------------------------------------------------------------------------------------------
// Exchange (transpose between "thread = frame t" and "thread = row frequency k1") through a
// per-warp shared buffer of CPW cubes x C*C 16-byte vectors.  One round moves vector h of
// every row; chunk (t, k1) sits at t*C + (k1 ^ t) so both the 8 writers (t = 0..7, fixed k1)
// and the 8 readers (k1 = 0..7, fixed t) of a quarter warp hit 8 distinct bank groups.
// ------------------------------------------------------------------------------------------
template <int C, typename T>
struct Xch {
    static constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte vector
    static constexpr int VPR = C / VEC;                 // vectors per row (rounds)
    static constexpr int CUBE_BYTES = C * C * 16;
    static constexpr int WARP_BYTES = (32 / C) * CUBE_BYTES;

    // one round: vector h (elements h*VEC .. h*VEC+VEC-1) of every row
    static __device__ __forceinline__ void round(uint8_t *wbuf, int cl, int r, const T (&a)[C][C], T (&b)[C][C], int h)
    {
        uint4 *base = reinterpret_cast<uint4 *>(wbuf + cl * CUBE_BYTES);
#pragma unroll
        for (int k1 = 0; k1 < C; k1++) {
            uint4 v;
            T *pv = reinterpret_cast<T *>(&v);
#pragma unroll
            for (int e = 0; e < VEC; e++) pv[e] = a[k1][h * VEC + e];
            base[r * C + (k1 ^ r)] = v;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < C; t++) {
            const uint4 v = base[t * C + (r ^ t)];
            const T *pv = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int e = 0; e < VEC; e++) b[t][h * VEC + e] = pv[e];
        }
        __syncwarp();
    }
    // a[k1][k2] (thread = t)  ->  b[t][k2] (thread = j = k1)
    static __device__ __forceinline__ void transpose(uint8_t *wbuf, int cl, int r, const T (&a)[C][C], T (&b)[C][C])
    {
#pragma unroll
        for (int h = 0; h < VPR; h++) round(wbuf, cl, r, a, b, h);
    }
};

template <int C, typename T>
__device__ __forceinline__ void fwd_xy(T (&a)[C][C])
{
#pragma unroll
    for (int y = 0; y < C; y++) Dct1D<C, T>::template fwd<1>(&a[y][0]);
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template fwd<C>(&a[0][x]);
}
template <int C, typename T>
__device__ __forceinline__ void inv_yx(T (&a)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template inv<C>(&a[0][x]);
#pragma unroll
    for (int y = 0; y < C; y++) Dct1D<C, T>::template inv<1>(&a[y][0]);
}
// Scaled pipeline of the fused kernels (dct_math.h): x and y butterflies normalised, the t butterfly
// carries S[k2] in its constants, S[k1] (k1 = lane) lives in the quantiser / dequantiser tables.
template <int C, typename T>
__device__ __forceinline__ void fwd_xy_n(T (&a)[C][C])
{
#pragma unroll
    for (int y = 0; y < C; y++) Dct1D<C, T>::template fwd_n<1>(&a[y][0]);
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template fwd_n<C>(&a[0][x]);
}
template <int C, typename T>
__device__ __forceinline__ void inv_yx_n(T (&a)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template inv_n<C>(&a[0][x]);
#pragma unroll
    for (int y = 0; y < C; y++) Dct1D<C, T>::template inv_n<1>(&a[y][0]);
}
// The same with the column transforms of the y pass (and, in inv_t_n_masked, of the t pass) skipped for
// the k2 columns that are zero in every cube of the warp: bit x of colmask = column x holds a non-zero.
// The inverse transform of an all-zero column is all zero, so the result is bit-identical; the branch is
// warp-uniform.
template <int C, typename T>
__device__ __forceinline__ void inv_yx_n_masked(T (&a)[C][C], uint32_t colmask)
{
#pragma unroll
    for (int x = 0; x < C; x++)
        if ((colmask >> x) & 1u) Dct1D<C, T>::template inv_n<C>(&a[0][x]);
#pragma unroll
    for (int y = 0; y < C; y++) Dct1D<C, T>::template inv_n<1>(&a[y][0]);
}
template <int C, typename T>
__device__ __forceinline__ void inv_t_n_masked(T (&b)[C][C], uint32_t colmask)
{
#pragma unroll
    for (int x = 0; x < C; x++)
        if ((colmask >> x) & 1u) Dct1D<C, T>::template inv_n<C>(&b[0][x]);
}
template <int C, typename T>
__device__ __forceinline__ void fwd_t_g(T (&b)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template fwd_g<C>(&b[0][x], Dct1D<C, T>::scale(x));
}
template <int C, typename T>
__device__ __forceinline__ void inv_t_g(T (&b)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template inv_g<C>(&b[0][x], Dct1D<C, T>::scale(x));
}
template <int C, typename T>
__device__ __forceinline__ void inv_t_n(T (&b)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template inv_n<C>(&b[0][x]);
}
// S[k1] for a run-time k1 (the lane)
template <int C>
__device__ __forceinline__ float lane_scale(int k1)
{
    float v = Dct1D<C, float>::scale(0);
#pragma unroll
    for (int k = 1; k < C; k++) v = k1 == k ? Dct1D<C, float>::scale(k) : v;
    return v;
}

template <int C, typename T>
__device__ __forceinline__ void fwd_t(T (&b)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template fwd<C>(&b[0][x]);
}
template <int C, typename T>
__device__ __forceinline__ void inv_t(T (&b)[C][C])
{
#pragma unroll
    for (int x = 0; x < C; x++) Dct1D<C, T>::template inv<C>(&b[0][x]);
}

// clamp to [0,255] + truncate toward zero (Decoder.java:74-80,112; decoder.c:29; 3dDCT.cl:255-262)
__device__ __forceinline__ uint32_t f2u8_sat(float v)
{
    uint32_t r;
    asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// u8 -> f32 without the conversion pipe: PRMT the byte into the mantissa of 2^23, subtract.
__device__ __forceinline__ float byte_to_float(uint32_t word, int i)
{
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u + i)) - 8388608.0f;
}

// u8 -> 2^15 + value in ONE PRMT (the byte becomes mantissa bits 8..15 of 2^15).  The bias is not
// subtracted per sample: the first butterfly level forms exact sums and differences, every output
// but DC is built from differences (the bias cancels exactly), and the DC term carries C*2^15 per
// axis, all still exact in fp32 (integers below 2^24 * ulp).  One FADD per plane removes it after
// the x and y passes (kBiasXY), instead of one per sample.
__device__ __forceinline__ float byte_to_float_biased(uint32_t word, int i)
{
    return __uint_as_float(__byte_perm(word, 0x47000000u, 0x7404u + (i << 4)));
}
template <int C> struct BiasXY { static constexpr float value = (C == 8 ? 64.0f : 4.0f) * 32768.0f; };

// ------------------------------------------------------------------------------------------
// Entropy stage, shared by the fused encoder and the int16-cube encoder.
// ------------------------------------------------------------------------------------------
struct GlobalSink {
    uint32_t *words;
    __device__ __forceinline__ void put(uint64_t idx, uint32_t be, bool shared)
    {
        if (shared) atomicOr(words + idx, be); else words[idx] = be;
    }
};

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

// A status word carries its flag AND its value in one 64-bit word and publishes nothing else, so relaxed gpu-scope accesses
// are enough.  (Round 2: they were acquire / release; every acquire poll of the look-back is a CCTL.IVALL, an invalidation
// of the SM's whole L1 -- 279 k of them per launch of the packer, whose write pass re-reads the chunks its count pass loaded.)
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE full warp.  total = this tile's bit count.  Returns the absolute bit offset of
// the tile (decoupled look-back over the predecessors' status words).
__device__ __forceinline__ unsigned long long tile_lookback(unsigned long long *status, long long tile,
                                                            unsigned long long total, unsigned long long start_bit, int lane,
                                                            unsigned int *err)
{
    if (tile == 0) {
        if (lane == 0) st_status(status, kFlagPrefix | (start_bit + total));
        return start_bit;
    }
    if (lane == 0) st_status(status + tile, kFlagAgg | total);
    unsigned long long excl = 0;
    long long look = tile - 1;
    for (;;) {
        const long long idx = look - lane;
        unsigned long long st;
        if (idx >= 0) {
            unsigned spins = 0;
            do {
                st = ld_status(status + idx);
                // a predecessor always holds an earlier ticket and is therefore resident; the bound only
                // turns a logic error into a reported failure instead of a hung GPU
                if ((st >> 62) == 0 && ++spins > (1u << 24)) { atomicOr(err, 8u); st = kFlagPrefix; }
            } while ((st >> 62) == 0);
        } else {
            st = (idx == -1) ? (kFlagPrefix | start_bit) : kFlagAgg;  // virtual tile -1 carries start_bit
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = pmask ? (__ffs((int)pmask) - 1) : 32;
        unsigned long long v = (lane <= first) ? (st & kValMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (pmask) break;
        look -= 32;
    }
    if (lane == 0) st_status(status + tile, kFlagPrefix | (excl + total));
    return excl;
}

// ------------------------------------------------------------------------------------------
// Encoder kernel 1: u8 frames -> quantised coefficients.
//   MODE_ZZ : zig-zag ordered int16 cubes in a dense-addressed, SPARSELY TOUCHED scratch
//             [cube][CS] plus one chunk mask per cube; only the 32-byte chunks (16 coefficients)
//             that hold a non-zero value are ever written (3.3 of 32 on natural content), and
//             kernel 2 only reads those.  This is the one place a coefficient is written.
//   MODE_NAT: natural-order int16 cubes (the dct3d_quantize_u8 stage entry point).
// Every warp runs its own pipeline: a unit (32 px x C rows x C frames = 32/C cubes) arrives by ONE
// TMA box into the warp's private double buffer and completes on the warp's own mbarrier, so the
// warps of a CTA never wait for each other (no CTA-wide barrier in the loop).
// ------------------------------------------------------------------------------------------
constexpr int MODE_ZZ = 0, MODE_NAT = 1;

template <int C>
struct EncSmem {
    using G = Geo<C>;
    static constexpr int IN_BYTES = 2 * G::UNIT_BYTES;                              // double buffer
    static constexpr int XCH_BYTES = Xch<C, float>::WARP_BYTES;
    static constexpr int ZZ_BYTES = G::CPW * G::ZZ_STRIDE * 2;
    static constexpr int WARP_BYTES = (IN_BYTES + XCH_BYTES + ZZ_BYTES + 255) / 256 * 256;   // swizzle phase = address bit 7
    static constexpr int BAR_OFF = kWarps * WARP_BYTES;
    static constexpr int TOTAL = BAR_OFF + kWarps * 16;
};

// Resident CTAs per SM (measured, round 2, 1920x1080x256): 4 CTAs = 128 registers (20 B of spills, tid-derived values re-read
// from S2R all over the loop) 418 us; 3 CTAs = 168 registers, no spills, 5% fewer instructions: 394 us; 5 CTAs = 96 registers:
// 607 us.  The kernel is issue-bound, so instructions count for more than resident warps.  Storing every quantised value
// instead of only the non-zero ones (no compare, 64 unconditional STS.U16 per thread) loads the shared-memory pipe: 450-459 us.
#ifndef DCT3D_ENC_CTAS
#define DCT3D_ENC_CTAS 3
#endif

// SKIP (C = 8, MODE_ZZ): zero-run skipping in the zig-zag stage.  97% of the quantised coefficients are zero, and the
// high diagonals k0 + k2 of a unit are zero in nearly every lane (measured on the benchmark clip: k0 + k2 >= 8 holds a
// non-zero in 4% of the units, 7 in 15%, 6 in 38%).  One FMNMX3 per three coefficients finds the largest magnitude of a
// group of diagonals, one compare against the lane's zero threshold and a warp vote decide whether the warp runs the
// group's quantiser (FFMA + compare + predicated STS.U16 per coefficient) at all.  Bit-identical: the skipped
// coefficients are exactly those the quantiser would have found zero.
template <int C, int MODE, bool SKIP = false>
__global__ void __launch_bounds__(kThreads, DCT3D_ENC_CTAS)
encode_kernel(const __grid_constant__ CUtensorMap tmap, const EncParams P)
{
    using G = Geo<C>;
    using S = EncSmem<C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *s_in = smem + warp * S::WARP_BYTES;
    uint8_t *s_xch = s_in + S::IN_BYTES;
    int16_t *s_zz = reinterpret_cast<int16_t *>(s_xch + S::XCH_BYTES);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + S::BAR_OFF) + warp * 2;

    const int cl = lane / C, r = lane % C;   // cube within the warp; t (stage 1) or k1 (stage 2)
    const Layout &L = P.L;
    const bool tma = P.use_tma != 0;

    float rq[G::NDIAG];
    uint32_t zb[G::NDIAG];
#pragma unroll
    for (int s = 0; s < G::NDIAG; s++) {
        rq[s] = lane_scale<C>(r) / (float)quant_divisor(s + r);   // S[k1] of the scaled butterflies folded in
        zb[s] = zz_base<C>(r, s);
    }
    // SKIP: zero thresholds of diagonals 6, 7 and 8; rq falls with s, so the threshold of 8 is a (sufficient) bound for
    // every later diagonal
    constexpr int SK0 = C == 8 ? 6 : G::NDIAG;                    // first diagonal that is tested before it is quantised
    constexpr int SKG = C == 8 ? 8 : G::NDIAG;                    // diagonals from here on are tested as one group
    const float thr_a = SKIP ? zero_threshold(rq[SK0 % G::NDIAG]) : 0.f;
    const float thr_b = SKIP ? zero_threshold(rq[(SK0 + 1) % G::NDIAG]) : 0.f;
    const float thr_g = SKIP ? zero_threshold(rq[SKG % G::NDIAG]) : 0.f;
    // unit counts fit 31 bits (checked by the host)
    const int nw = (int)gridDim.x * kWarps, nunits = (int)L.nunits;
    int u = (int)blockIdx.x * kWarps + warp;
    UnitPos pos = unit_pos(L, u);
    const UnitPos step = unit_pos(L, nw);

    auto issue_tma = [&](const UnitPos &q, int h) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&s_bar[h], G::UNIT_BYTES);
        tma_load_3d(s_in + h * G::UNIT_BYTES, &tmap, q.bxu * kUnitW, q.slab * C, q.byi * C, &s_bar[h]);
    };
    if (lane == 0 && tma) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        if (u < nunits) issue_tma(pos, 0);
    }
    // The zig-zag buffer is kept all-zero between units: only non-zero coefficients are scattered
    // into it (97% are zero), and the lanes that find a non-zero chunk wipe it after storing it.
    for (int i = lane; i < S::ZZ_BYTES / 16; i += 32) reinterpret_cast<uint4 *>(s_zz)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();

    for (int it = 0; u < nunits; u += nw, it++) {
        const int h = it & 1;
        uint8_t *box = s_in + h * G::UNIT_BYTES;
        // every lane finished reading in[h^1] before the exchange barriers of the previous pass
        if (lane == 0 && tma && u + nw < nunits) {
            UnitPos nx = pos;
            unit_advance(L, nx, step);
            issue_tma(nx, h ^ 1);
        }
        const int bxi0 = pos.bxu * G::CPW;
        const int nvalid = min(G::CPW, L.bx - bxi0);
        const long long cube0 = ((long long)pos.slab * L.by + pos.byi) * L.bx + bxi0;
        if (!tma) {
            // plain loader: C-byte pieces (always aligned since W % C == 0), same layout
            constexpr int PIECES = G::UNIT_BYTES / C;
#pragma unroll
            for (int i = lane; i < PIECES; i += 32) {
                const int row = i / G::CPW;                 // y*C + t
                const int cb = i - row * G::CPW;            // cube in unit
                const int y = row / C, t = row - y * C;
                uint32_t lo = 0, hi = 0;
                if (cb < nvalid) {
                    const uint8_t *src = P.frames + ((size_t)(pos.slab * C + t) * L.H + (pos.byi * C + y)) * L.W +
                                         (size_t)(bxi0 + cb) * C;
                    if (C == 8) { const uint2 v = *reinterpret_cast<const uint2 *>(src); lo = v.x; hi = v.y; }
                    else lo = *reinterpret_cast<const uint32_t *>(src);
                }
                uint8_t *dst = box + unit_offset<C>(y, t, cb * C);
                if (C == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
                else *reinterpret_cast<uint32_t *>(dst) = lo;
            }
            __syncwarp();
        } else {
            const uint32_t parity = (uint32_t)(it >> 1) & 1u;
            unsigned spins = 0;
            while (!mbar_try_wait(&s_bar[h], parity)) {
                if (++spins > (1u << 22)) { if (lane == 0) atomicOr(P.err, 16u); break; }   // never hang the GPU
            }
        }
        unit_advance(L, pos, step);

        // ---- transform: one cube per C threads ---------------------------------------------
        float a[C][C], bq[C][C];
#pragma unroll
        for (int y = 0; y < C; y++) {
            const uint8_t *src = box + unit_offset<C>(y, r, cl * C);
            if (C == 8) {
                const uint2 v = *reinterpret_cast<const uint2 *>(src);
#pragma unroll
                for (int x = 0; x < 4; x++) { a[y][x] = byte_to_float_biased(v.x, x); a[y][(x + 4) % C] = byte_to_float_biased(v.y, x); }
            } else {
                const uint32_t v = *reinterpret_cast<const uint32_t *>(src);
#pragma unroll
                for (int x = 0; x < 4; x++) a[y][x % C] = byte_to_float_biased(v, x);
            }
        }
        fwd_xy_n<C, float>(a);
        a[0][0] -= BiasXY<C>::value;
        Xch<C, float>::transpose(s_xch, cl, r, a, bq);
        fwd_t_g<C, float>(bq);
        // bq[k0][k2] is coefficient (k0, k1 = r, k2)
        if (MODE == MODE_NAT) {
            if (cl < nvalid) {
                int16_t *dst = P.qcubes + (size_t)(cube0 + cl) * G::CS + r * C;
#pragma unroll
                for (int k0 = 0; k0 < C; k0++) {
                    uint32_t w[C / 2];
#pragma unroll
                    for (int k2 = 0; k2 < C; k2 += 2) {
                        const uint32_t q0 = (uint32_t)quantize_f32(bq[k0][k2], rq[k0 + k2]) & 0xffffu;
                        const uint32_t q1 = (uint32_t)quantize_f32(bq[k0][k2 + 1], rq[k0 + k2 + 1]) & 0xffffu;
                        w[k2 / 2] = q0 | (q1 << 16);
                    }
                    if (C == 8) *reinterpret_cast<uint4 *>(dst + k0 * C * C) = make_uint4(w[0], w[1], w[2 % (C / 2)], w[3 % (C / 2)]);
                    else *reinterpret_cast<uint2 *>(dst + k0 * C * C) = make_uint2(w[0], w[1]);
                }
            }
            continue;
        }
        // zig-zag scatter into this warp's private cubes (runs of a diagonal are contiguous)
        int16_t *zz = s_zz + cl * G::ZZ_STRIDE;
        // quantise + scatter the diagonals [s_lo, s_hi] of this thread's plane
        auto scatter_diagonals = [&](int s_lo, int s_hi) {
#pragma unroll
            for (int k0 = 0; k0 < C; k0++) {
#pragma unroll
                for (int k2 = 0; k2 < C; k2++) {
                    const int s = k0 + k2;
                    if (s < s_lo || s > s_hi) continue;
                    const int k0min = s > C - 1 ? s - (C - 1) : 0;
                    // one FFMA quantises (magic rounding); a zero result is exactly the magic constant
                    const int bits = __float_as_int(fmaf(bq[k0][k2], rq[s], DCT_MAGIC));
                    if (bits != 0x4B400000) zz[zb[s] + (k0 - k0min)] = (int16_t)bits;
                }
            }
        };
        // largest magnitude on the diagonals [s_lo, s_hi]
        auto max_abs = [&](int s_lo, int s_hi) {
            float m = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < C; k0++) {
#pragma unroll
                for (int k2 = 0; k2 < C; k2++) {
                    const int s = k0 + k2;
                    if (s >= s_lo && s <= s_hi) m = fmaxf(m, fabsf(bq[k0][k2]));
                }
            }
            return m;
        };
        if (!SKIP) {
            scatter_diagonals(0, G::NDIAG - 1);
        } else {
            scatter_diagonals(0, SK0 - 1);
            if (__any_sync(0xffffffffu, max_abs(SK0, SK0) > thr_a)) scatter_diagonals(SK0, SK0);
            if (__any_sync(0xffffffffu, max_abs(SK0 + 1, SK0 + 1) > thr_b)) scatter_diagonals(SK0 + 1, SK0 + 1);
            if (__any_sync(0xffffffffu, max_abs(SKG, G::NDIAG - 1) > thr_g)) scatter_diagonals(SKG, G::NDIAG - 1);
        }
        __syncwarp();
        // chunk masks + sparse store: lane <-> 16-coefficient chunk
        constexpr int ITER = (G::CPW * G::CHUNKS) / 32;   // 4 (C=8) / 1 (C=4)
        uint32_t run = 0;                                 // non-zero chunks of this unit stored so far
        int16_t *const unit_zz = P.zzg + (size_t)cube0 * G::CS;
#pragma unroll
        for (int k = 0; k < ITER; k++) {
            const int ci = k * 32 + lane;
            const int cube = ci / G::CHUNKS, chunk = ci % G::CHUNKS;
            uint4 *q = reinterpret_cast<uint4 *>(s_zz + cube * G::ZZ_STRIDE + chunk * 16);
            // lanes 4..7 of every 8 read their two 16-byte halves in the other order: the quarter
            // warp then touches 8 distinct bank groups instead of 4 twice
            const int hsel = (lane >> 2) & 1;
            const uint4 va = q[hsel], vb = q[hsel ^ 1];
            const uint32_t any = va.x | va.y | va.z | va.w | vb.x | vb.y | vb.z | vb.w;
            const bool ok = cube < nvalid;
            const uint32_t bal = __ballot_sync(0xffffffffu, any != 0);
            const long long gc = cube0 + cube;
            if (any != 0) {
                if (ok) {
                    const uint32_t at = run + (uint32_t)__popc(bal & ((1u << lane) - 1u));   // dense: rank among the unit's non-zero chunks
                    uint4 *dst = reinterpret_cast<uint4 *>(unit_zz + (size_t)at * 16);
                    dst[hsel] = va;
                    dst[hsel ^ 1] = vb;
                }
                q[0] = make_uint4(0, 0, 0, 0);      // leave the buffer clean for the next unit
                q[1] = make_uint4(0, 0, 0, 0);
            }
            run += (uint32_t)__popc(bal);
            if (G::CHUNKS == 32) { if (lane == 0 && ok) P.cmask[gc] = bal; }
            else { if (chunk == 0 && ok) P.cmask[gc] = (bal >> (cube * G::CHUNKS)) & ((1u << (G::CHUNKS & 31)) - 1u); }
        }
        __syncwarp();                           // zz is rewritten by the next unit
    }
}

// Where the packer finds a cube's chunks in the scratch: the unit's region starts at its first cube's slot, the cube's chunks
// follow those of the cubes before it in the unit.
template <int C>
__device__ __forceinline__ const int16_t *cube_chunks(const EncParams &P, long long cube)
{
    using G = Geo<C>;
    const int idx = (int)(cube % P.L.bx) % G::CPW;                 // position of the cube inside its warp unit
    const long long first = cube - idx;
    uint32_t before = 0;
    for (int j = 0; j < idx; j++) before += (uint32_t)__popc(P.cmask[first + j]);
    return P.zzg + (size_t)first * G::CS + (size_t)before * 16;
}

// Natural-order int16 cubes -> the same zig-zag chunk scratch + masks (dct3d_eg_encode_i16).
// One warp per unit (the CPW cubes a warp of the fused encoder would hold), lane <-> chunk.
template <int C>
__global__ void __launch_bounds__(kThreads)
zz_gather_kernel(const EncParams P)
{
    using G = Geo<C>;
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarps;
    const uint16_t *lin = zz_lin<C>();
    const int bx = P.L.bx, bxu = P.L.bxu;
    const long long nrows = (P.L.ncubes + bx - 1) / bx;
    for (long long u = wid; u < nrows * bxu; u += nw) {
        const long long row = u / bxu;
        const int ux = (int)(u - row * bxu);
        const long long cube0 = row * bx + (long long)ux * G::CPW;
        const long long row_end = min(P.L.ncubes, (row + 1) * bx);
        const int nvalid = (int)min((long long)G::CPW, row_end - cube0);
        int16_t *const unit_zz = P.zzg + (size_t)cube0 * G::CS;
        uint32_t run = 0;
        for (int k = 0; k < nvalid; k++) {
            const int16_t *src = P.qcubes_in + (size_t)(cube0 + k) * G::CS;
            const int chunk = lane;
            uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t any = 0;
            if (chunk < G::CHUNKS) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t v = (uint16_t)src[lin[chunk * 16 + i]];
                    w[i >> 1] |= v << ((i & 1) * 16);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) any |= w[i];
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, any != 0);
            if (any) {
                uint4 *dst = reinterpret_cast<uint4 *>(unit_zz + (size_t)(run + (uint32_t)__popc(bal & ((1u << lane) - 1u))) * 16);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            run += (uint32_t)__popc(bal);
            if (lane == 0) P.cmask[cube0 + k] = bal;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Encoder kernel 2: Exp-Golomb bit packing, one thread per cube, 256 cubes per CTA tile.
//   count pass (only the non-zero chunks) -> block scan -> decoupled look-back across tiles
//   (tiles taken from a ticket, in stream order) -> write pass straight into the stream.
// The serial bit append per thread is latency-bound; it runs at full occupancy so that other
// warps cover it.
// ------------------------------------------------------------------------------------------
// Measured and dropped in round 2: a software-pipelined CTA (a ninth "scanner" warp resolves the look-back of tile i while
// the eight worker warps count tile i+1, aggregates published early): 340 us instead of 173 -- two barriers per tile, a
// lower occupancy (48 registers) and an idle warp cost more than the 28% barrier stall it was meant to remove.
#ifndef DCT3D_PACK_WORKERS
#define DCT3D_PACK_WORKERS 256
#endif
constexpr int kPackWorkers = DCT3D_PACK_WORKERS;    // one cube each
constexpr int kPackThreads = kPackWorkers;

// The CTA is software-pipelined over its tiles: while the scanner warp resolves the bit offset of tile i (decoupled
// look-back: a chain of global-memory round trips), the 8 worker warps already run the count pass of tile i+1; the write
// pass of tile i follows the barrier.  (First version: all 8 warps waited at a barrier around the look-back, 28% of the
// kernel's warp time.)
// 8 CTAs per SM = 32 registers = full occupancy: the serial bit append is latency-bound (39 registers at 6 CTAs: +6%)
template <int C>
__global__ void __launch_bounds__(kPackThreads, 2048 / kPackThreads)
eg_pack_kernel(const EncParams P)
{
    using G = Geo<C>;
    __shared__ uint32_t s_wsum[kPackThreads / 32];
    __shared__ unsigned long long s_off;
    __shared__ long long s_tile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (P.L.ncubes + kPackWorkers - 1) / kPackWorkers;
    const unsigned long long start_bit = P.start_bit_dev ? *P.start_bit_dev : P.start_bit;
    for (;;) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.ticket, 1u);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        const long long cube = tile * kPackWorkers + tid;
        const bool valid = cube < P.L.ncubes;
        const int16_t *zz = valid ? cube_chunks<C>(P, cube) : P.zzg;
        const uint32_t cm = valid ? P.cmask[cube] : 0u;
        const uint32_t nb = valid ? eg_count_cube<G::CS, true>(zz, cm) : 0u;
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = lane < kPackThreads / 32 ? s_wsum[lane] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            if (lane < kPackThreads / 32) s_wsum[lane] = wi - w;          // exclusive warp offsets
            const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
            const unsigned long long off = tile_lookback(P.tile_status, tile, total, start_bit, lane, P.err);
            if (lane == 0) {
                s_off = off;
                if (tile == ntiles - 1) {
                    *P.end_bit = off + total;
                    if (P.end_bit_host) *P.end_bit_host = off + total;
                }
            }
        }
        __syncthreads();
        if (valid) {
            const unsigned long long off = s_off + s_wsum[warp] + (incl - nb);
            if (off + nb + 64 > P.cap_bits) {
                atomicOr(P.err, 1u);
            } else {
                GlobalSink sink{P.out_words};
                eg_write_cube<G::CS, GlobalSink, true>(zz, cm, off, sink);
            }
        }
        __syncthreads();                        // s_tile / s_wsum / s_off are reused
    }
}

// The same packer with the cubes of a tile dealt to the threads in order of their chunk counts.  A thread's work is
// proportional to the number of non-zero chunks of its cube, and a warp runs as long as its longest lane: in stream order a
// warp of the benchmark clip runs 5.3 chunk iterations for a mean of 3.3 per cube (62% lane efficiency); sorted inside the
// tile it runs 3.6 (91%).  Which thread packs which cube is free -- only the prefix sum of the bit counts has to follow
// the stream order -- so the tile is counting-sorted by popc(chunk mask) in shared memory (33 keys, warp-aggregated
// histogram), thread t counts and later writes cube perm[t], and the bit counts meet in stream order in shared memory for
// the scan.  Same stream, bit for bit.
template <int C>
__global__ void __launch_bounds__(kPackThreads, 2048 / kPackThreads)
eg_pack_sorted_kernel(const EncParams P)
{
    using G = Geo<C>;
    constexpr int NW = kPackThreads / 32;
    static_assert(NW <= 32 && kPackWorkers <= 65536, "tile geometry");
    __shared__ uint32_t s_wsum[NW];
    __shared__ uint32_t s_bits[kPackWorkers];       // bit counts of the tile's cubes in stream order, then their exclusive offsets
    __shared__ uint16_t s_perm[kPackWorkers];       // sorted position -> cube of the tile
    __shared__ uint32_t s_hist[33];                 // cubes per chunk count
    __shared__ unsigned long long s_off;
    __shared__ long long s_tile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (P.L.ncubes + kPackWorkers - 1) / kPackWorkers;
    const unsigned long long start_bit = P.start_bit_dev ? *P.start_bit_dev : P.start_bit;
    for (;;) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.ticket, 1u);
        if (tid < 33) s_hist[tid] = 0u;
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        // ---- counting sort of the tile by chunk count ------------------------------------------------------------
        {
            const long long cube_t = tile * kPackWorkers + tid;
            const uint32_t key = cube_t < P.L.ncubes ? (uint32_t)__popc(P.cmask[cube_t]) : 0u;     // 0..32
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs((int)peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&s_hist[key], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            const uint32_t rank = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            __syncthreads();                            // the histogram is complete
            // every warp forms the exclusive prefix of keys 0..31 for itself (key 32 starts at their total)
            const uint32_t h = s_hist[lane];
            uint32_t incl = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t first = __shfl_sync(0xffffffffu, incl - h, (int)(key & 31u));
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            s_perm[(key < 32u ? first : total) + rank] = (uint16_t)tid;
        }
        __syncthreads();
        // ---- count pass, on the cube this thread was dealt -----------------------------------------------------------
        const int j = s_perm[tid];
        const long long cube = tile * kPackWorkers + j;
        const bool valid = cube < P.L.ncubes;
        const int16_t *zz = valid ? cube_chunks<C>(P, cube) : P.zzg;
        const uint32_t cm = valid ? P.cmask[cube] : 0u;
        const uint32_t nbj = valid ? eg_count_cube<G::CS, true>(zz, cm) : 0u;
        s_bits[j] = nbj;
        __syncthreads();
        // ---- scan in stream order ------------------------------------------------------------------------------------
        const uint32_t nb = s_bits[tid];
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        {
            // every warp scans the warp sums for itself; warp 0 also resolves the tile's bit offset
            const uint32_t w = lane < NW ? s_wsum[lane] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            const uint32_t wexcl = __shfl_sync(0xffffffffu, wi - w, warp);
            s_bits[tid] = wexcl + (incl - nb);          // exclusive offset of cube `tid` inside the tile (own slot: no reader yet)
            if (warp == 0) {
                const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
                const unsigned long long off = tile_lookback(P.tile_status, tile, total, start_bit, lane, P.err);
                if (lane == 0) {
                    s_off = off;
                    if (tile == ntiles - 1) {
                        *P.end_bit = off + total;
                        if (P.end_bit_host) *P.end_bit_host = off + total;
                    }
                }
            }
        }
        __syncthreads();
        // ---- write pass ----------------------------------------------------------------------------------------------
        if (valid) {
            const unsigned long long off = s_off + s_bits[j];
            if (off + nbj + 64 > P.cap_bits) {
                atomicOr(P.err, 1u);
            } else {
                GlobalSink sink{P.out_words};
                eg_write_cube<G::CS, GlobalSink, true>(zz, cm, off, sink);
            }
        }
        __syncthreads();                                // s_tile, s_hist, s_perm, s_bits, s_wsum and s_off are reused
    }
}

// The sorted deal with BALANCED warps.  With one cube per thread the sorted order gives warp 0 the lightest 32 cubes of a
// tile and the last warp the heaviest: the light warps wait at the tile's barriers (barrier stalls rose from 33% to 44% of
// the warp samples with the sort).  Here a tile is 2 x kPackThreads cubes, cut into 2 NW sorted blocks of 32, and warp w
// packs blocks w and 2 NW - 1 - w -- a light one and a heavy one: every warp's lanes still run cubes of like size, and
// every warp carries about the same load.  Between the passes a thread keeps nothing but its place: what it needs of a
// cube (mask, chunk pointer) is found again from the permutation, which costs a few L1 hits and keeps the kernel at 32
// registers.
// Measured (profiles/r2_runs/ab_variants_s2_call6.jsonl), against the sorted deal: 4^3 cubes -45 us per 256 frames, 8^3
// natural content +6 us, 8^3 noise +230 us (two dense cubes in a row per thread, half as many tiles in flight).  It is
// option pack_sort = 2, not the default.
constexpr int kPackBalTile = 2 * kPackThreads;

template <int C>
__global__ void __launch_bounds__(kPackThreads, 2048 / kPackThreads)
eg_pack_balanced_kernel(const EncParams P)
{
    using G = Geo<C>;
    constexpr int NW = kPackThreads / 32, TILE = kPackBalTile;
    static_assert(NW <= 32 && TILE <= 65536, "tile geometry");
    __shared__ uint32_t s_wsum[NW];
    __shared__ uint32_t s_bits[TILE];               // bit counts of the tile's cubes in stream order, then their exclusive offsets
    __shared__ uint16_t s_perm[TILE];               // sorted position -> cube of the tile
    __shared__ uint32_t s_hist[33];                 // cubes per chunk count
    __shared__ unsigned long long s_off;
    __shared__ long long s_tile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (P.L.ncubes + TILE - 1) / TILE;
    const unsigned long long start_bit = P.start_bit_dev ? *P.start_bit_dev : P.start_bit;
    // this thread's two places in the sorted order: a block from the light end and its mirror from the heavy end
    const int place_light = warp * 32 + lane, place_heavy = (2 * NW - 1 - warp) * 32 + lane;
    for (;;) {
        if (tid == 0) s_tile = (long long)atomicAdd(P.ticket, 1u);
        if (tid < 33) s_hist[tid] = 0u;
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        const long long cube0 = tile * TILE;
        // ---- counting sort of the tile by chunk count (two cubes per thread: tid and kPackThreads + tid) ----------
        {
            uint32_t key[2], rank[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const long long c = cube0 + q * kPackThreads + tid;
                key[q] = c < P.L.ncubes ? (uint32_t)__popc(P.cmask[c]) : 0u;              // 0..32
                const uint32_t peers = __match_any_sync(0xffffffffu, key[q]);
                const int leader = __ffs((int)peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&s_hist[key[q]], (uint32_t)__popc(peers));
                base = __shfl_sync(0xffffffffu, base, leader);
                rank[q] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            }
            __syncthreads();                            // the histogram is complete
            const uint32_t h = s_hist[lane];
            uint32_t incl = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t first = __shfl_sync(0xffffffffu, incl - h, (int)(key[q] & 31u));
                s_perm[(key[q] < 32u ? first : total) + rank[q]] = (uint16_t)(q * kPackThreads + tid);
            }
        }
        __syncthreads();
        // ---- count pass ------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            const int j = s_perm[q ? place_heavy : place_light];
            const long long cube = cube0 + j;
            uint32_t nb = 0;
            if (cube < P.L.ncubes) nb = eg_count_cube<G::CS, true>(cube_chunks<C>(P, cube), P.cmask[cube]);
            s_bits[j] = nb;
        }
        __syncthreads();
        // ---- scan in stream order: thread t owns cubes 2t and 2t + 1 of the tile --------------------------------------
        {
            const uint32_t b0 = s_bits[2 * tid], b1 = s_bits[2 * tid + 1];
            const uint32_t sum = b0 + b1;
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            const uint32_t w = lane < NW ? s_wsum[lane] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            const uint32_t excl = __shfl_sync(0xffffffffu, wi - w, warp) + (incl - sum);
            s_bits[2 * tid] = excl;                     // own slots: nobody else reads them before the next barrier
            s_bits[2 * tid + 1] = excl + b0;
            if (warp == 0) {
                const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
                const unsigned long long off = tile_lookback(P.tile_status, tile, total, start_bit, lane, P.err);
                if (lane == 0) {
                    s_off = off;
                    if (tile == ntiles - 1) {
                        *P.end_bit = off + total;
                        if (P.end_bit_host) *P.end_bit_host = off + total;
                    }
                }
            }
        }
        __syncthreads();
        // ---- write pass ----------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            const int j = s_perm[q ? place_heavy : place_light];
            const long long cube = cube0 + j;
            if (cube < P.L.ncubes) {
                const unsigned long long off = s_off + s_bits[j];
                // the cube's own bit count: the next cube's offset minus this one's (the tile total for the last cube)
                const int16_t *zz = cube_chunks<C>(P, cube);
                const uint32_t cm = P.cmask[cube];
                if (off + (uint32_t)G::CS * 33u + 64 > P.cap_bits && off + eg_count_cube<G::CS, true>(zz, cm) + 64 > P.cap_bits) {
                    atomicOr(P.err, 1u);
                } else {
                    GlobalSink sink{P.out_words};
                    eg_write_cube<G::CS, GlobalSink, true>(zz, cm, off, sink);
                }
            }
        }
        __syncthreads();                                // s_tile, s_hist, s_perm, s_bits, s_wsum and s_off are reused
    }
}

// Placement of a slab range's stream inside the clip's one stream (the rule of ExpGolomb.c:112-130 with
// encoder.c:263-271): a GPU codes its range from bit 0 of its own buffer; once the bit counts of the ranges before it are
// known, its bits move to phase = (global start bit) % 8, so that the host only has to copy whole bytes to byte
// (global start bit) / 8 and OR the one byte the range shares with its predecessor.  dst bit (phase + i) = src bit i;
// the first `phase` bits of dst are zero.  src must be zero beyond its last bit (it is: the packer's buffer is wiped).
__global__ void __launch_bounds__(256)
stream_shift_kernel(const uint32_t *__restrict__ src, unsigned long long src_words, uint32_t *__restrict__ dst,
                    unsigned long long dst_words, int phase)
{
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < dst_words;
         j += gridDim.x * (unsigned long long)blockDim.x) {
        const uint32_t cur = j < src_words ? bswap32(__ldg(src + j)) : 0u;
        const uint32_t prev = (j > 0 && j - 1 < src_words) ? bswap32(__ldg(src + j - 1)) : 0u;
        dst[j] = bswap32(__funnelshift_r(cur, prev, phase));      // (prev:cur) >> phase, low word
    }
}

// ------------------------------------------------------------------------------------------
// Decode side
// ------------------------------------------------------------------------------------------
// The segment kernels give every thread 1024 consecutive stream bits.  Read word by word from
// global memory that is 32 different cache lines per warp load (8x over-fetch out of L2), so the
// CTA first stages its contiguous part of the stream in shared memory with coalesced loads,
// byte-swapped, one padding word per 32 so that threads 32 words apart hit different banks.
constexpr int kSegWords = 32;                       // 1024-bit segments
constexpr int kSegThreads = 128;
constexpr unsigned kLeadBits = 96;                  // lead-in walk of the speculative scan
constexpr int kStageN = kSegThreads * kSegWords + 32;   // + margin for codes running past the last segment
constexpr int kStageSmem = kStageN + kStageN / 32 + 1;

// Relative addressing shared by both word sources: word j / bit positions are relative to word w0.
struct WordBase {
    unsigned long long w0;
    __device__ __forceinline__ uint32_t rel(unsigned long long abs_bit) const
    {
        const unsigned long long base = w0 * 32ull;
        const unsigned long long d = abs_bit > base ? abs_bit - base : 0ull;
        return d > 0xffffffffull ? 0xffffffffu : (uint32_t)d;
    }
};

// The CTA's staged window.  The scan only looks at bit positions below its segment's end (plus 16 for the
// second half of a 33-bit code), and a look reads two words: with the 4 lead words in front and 28 words of
// margin behind the last segment every index is inside the window, whatever the stream holds.
struct StagedSource : WordBase {
    const uint32_t *s;
    __device__ __forceinline__ uint32_t word(uint32_t j) const { return s[j + (j >> 5)]; }
};

// Plain global reads (fix-up kernel: scattered segments).
struct GlobalSource : WordBase {
    const uint32_t *words; unsigned long long nwords;
    __device__ __forceinline__ uint32_t word(uint32_t j) const
    {
        const unsigned long long i = w0 + j;
        return i < nwords ? bswap32(__ldg(words + i)) : 0u;
    }
};

// One segment (plus the reader's look-ahead) copied into the thread's local memory with independent
// loads: for the single thread that re-walks a segment serially.
constexpr int kLocalWords = kSegWords + 8;
struct LocalSource : WordBase {
    uint32_t w[kLocalWords];
    __device__ __forceinline__ void fill(const uint32_t *words, unsigned long long nwords)
    {
#pragma unroll
        for (int j = 0; j < kLocalWords; j++) w[j] = w0 + j < nwords ? bswap32(__ldg(words + w0 + j)) : 0u;
    }
    __device__ __forceinline__ uint32_t word(uint32_t j) const { return w[min(j, (uint32_t)kLocalWords - 1u)]; }
};

__device__ __forceinline__ StagedSource stage_stream(uint32_t *s_words, const uint32_t *words, unsigned long long nwords,
                                                     unsigned long long start_bit, unsigned long long first_seg,
                                                     unsigned int lead_words = 0)
{
    const unsigned long long w0 = (start_bit >> 5) + first_seg * kSegWords - lead_words;
    for (int j = threadIdx.x; j < kStageN; j += blockDim.x) {
        const unsigned long long i = w0 + j;
        s_words[j + (j >> 5)] = i < nwords ? bswap32(__ldg(words + i)) : 0u;
    }
    __syncthreads();
    StagedSource src;
    src.w0 = w0;
    src.s = s_words;
    return src;
}

// A 1024-bit segment holds at most 342 non-zero codes (3 bits each at least).  The lists live in a
// dense-addressed scratch of which only the heads (about 24 entries per segment on natural content) are
// ever touched, interleaved over the 32 segments of a warp: vector v (4 entries) of segment k sits at
// [k / 32][v][k % 32], so a warp that walks 32 consecutive segments reads and writes whole 512-byte
// lines.  Entries leave the scanning thread four at a time as one 16-byte store.
constexpr int kSegListVec = 86;                      // uint4 per segment: 344 entries >= 342
__device__ __forceinline__ unsigned long long seg_list_base(unsigned long long k)
{
    return (k >> 5) * (unsigned long long)(kSegListVec * 32) + (k & 31);
}
struct SegListSink {
    uint4 *dst;                                      // next vector of this segment (stride 32 vectors)
    uint32_t e0, e1, e2, e3;                         // shift register: branch-free in the scan's hot loop
    __device__ __forceinline__ void push(uint32_t i, uint32_t e)
    {
        e0 = e1; e1 = e2; e2 = e3; e3 = e;
        if ((i & 3u) == 3u) { *dst = make_uint4(e0, e1, e2, e3); dst += 32; }
    }
    __device__ __forceinline__ void flush(uint32_t n)
    {
        while (n & 3u) push(n++, 0u);                // pad the last vector
    }
};

struct DecParams {
    Layout L;
    const uint32_t *words; unsigned long long nwords; unsigned long long nbits_total;  // stream
    unsigned long long start_bit;
    unsigned long long count_end_bit;   // codes that START before this bit are counted (= nbits_total, except for a part of a stream
                                        // whose successor part is counted by another GPU: the window then extends past it)
    unsigned long long code_base, nz_base;   // codes / non-zero codes of the clip in front of start_bit (a later piece of a stream
                                             // that is parsed piece by piece: the prefix sums continue, lists and row pointers are global)
    int first_entry;                    // entry point of segment 0: 0 = the stream's first code starts at start_bit; -1 = unknown,
                                        // guess it like any other segment's (a part in the middle of a stream; needs 128 bits of
                                        // stream in front of start_bit); > 0 = that many bits + 1 (a verified overhang)
    // index discovery
    unsigned int seg_bits; unsigned long long nseg;
    unsigned int *seg_count;            // [nseg] codes starting in the segment | non-zero codes << 16 | malformed << 31
    unsigned int *seg_over;             // [nseg+1] overhang INTO segment k (seg_over[0] = 0)
    unsigned int *seg_used;             // [nseg] entry overhang used for the current count
    uint4 *seg_list;                    // [nseg][kSegListVec] non-zero codes of each segment (sparsely touched)
    unsigned long long *seg_first;      // [nseg+1] exclusive prefix of the code counts
    unsigned long long *seg_nzfirst;    // [nseg+1] exclusive prefix of the non-zero code counts
    unsigned int *changed;              // number of the last fix-up round that moved an overhang
    unsigned int *err;                  // bit1 = malformed (a truncated stream shows as too few codes: DCT3D_E_NEED_MORE)
    unsigned long long *end_bit;        // out: first bit after the last code
    uint32_t *coo;                      // non-zero coefficients of the whole stream, in stream order: natural index << 16 | value
    unsigned long long *coo_start;      // [ncubes+1] first entry of every cube (CSR row pointers)
    int locate_only;                    // seg_emit_kernel: only report where the clip ends (dct3d_eg_locate)
    int16_t *qcubes;                    // natural-order cubes out (coo_scatter_kernel)
    uint8_t *frames;
};

// Pass 1: every thread scans one segment from a guessed entry point (see the lead-in walk), keeps the
// non-zero codes it meets, and records how far its last code runs into the next segment (seg_over[k+1]).
__global__ void __launch_bounds__(kSegThreads)
seg_scan_kernel(const DecParams P)
{
    __shared__ uint32_t s_words[kStageSmem];
    const unsigned long long k = blockIdx.x * (unsigned long long)kSegThreads + threadIdx.x;
    // the window starts 4 words early (not for the first CTA): room for the lead-in walk below
    const StagedSource src = stage_stream(s_words, P.words, P.nwords, P.start_bit, blockIdx.x * (unsigned long long)kSegThreads,
                                          (blockIdx.x || P.first_entry < 0) ? 4u : 0u);
    if (k >= P.nseg) return;
    const uint32_t seg0 = src.rel(P.start_bit + k * (unsigned long long)P.seg_bits);
    const uint32_t eos = src.rel(P.nbits_total);
    uint32_t lim = min(seg0 + P.seg_bits, src.rel(P.count_end_bit));
    if (lim > eos) lim = eos;
    uint32_t n = 0, next = 0, nz = 0;
    // Guess the entry point: walk the last kLeadBits bits of the previous segment from an arbitrary
    // phase.  Exp-Golomb streams resynchronise within a few codes (every run of one-bits is a run of
    // complete codes), so the walk usually arrives at the segment's first code; the guess is verified
    // against the predecessor's overhang by seg_fix_kernel like any other.
    uint32_t entry = (k == 0 && P.first_entry > 0) ? (uint32_t)P.first_entry - 1u : 0u;
    if ((k > 0 || P.first_entry < 0) && seg0 < lim) {
        uint32_t nd, nx;
        if (eg_scan_segment(src, seg0 - kLeadBits, seg0, eos, nd, nx) && nx - seg0 <= 33u) entry = nx - seg0;
    }
    // A scan from a wrongly assumed entry point may run into an impossible code; that is only an
    // error if it is still there once the entry points have converged, so it is recorded per segment.
    unsigned int bad = 0;
    SegListSink sink;
    sink.dst = P.seg_list + seg_list_base(k);
    sink.e0 = sink.e1 = sink.e2 = sink.e3 = 0u;
    if (seg0 + entry >= lim) { n = 0; next = seg0 + entry; }
    else if (!eg_scan_segment(src, seg0 + entry, lim, eos, n, next, &nz, sink)) { bad = 0x80000000u; n = 0; nz = 0; next = lim; }
    P.seg_count[k] = n | (nz << 16) | bad;
    P.seg_used[k] = entry;
    P.seg_over[k + 1] = next > lim ? next - lim : 0u;
    if (k == 0) P.seg_over[0] = entry;           // 0 for a whole stream: its first code starts at its first bit
}

// Fix-up rounds: every thread compares the entry point its segment was scanned with against the
// predecessor's overhang and re-scans the segment on a mismatch (one in seven without the lead-in walk,
// nearly none with it), until a round changes nothing.  `round` is stored in *changed by a thread that
// moves an overhang, so the flag needs no reset between rounds.
__global__ void __launch_bounds__(128)
seg_fix_kernel(const DecParams P, unsigned int round)
{
    const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k >= P.nseg) return;
    const unsigned int entry = P.seg_over[k];
    if (entry == P.seg_used[k]) return;
    LocalSource src;                                           // the segment is fetched with independent loads, then walked
    src.w0 = (P.start_bit >> 5) + k * kSegWords;
    src.fill(P.words, P.nwords);
    const uint32_t seg0 = src.rel(P.start_bit + k * (unsigned long long)P.seg_bits);
    const uint32_t eos = src.rel(P.nbits_total);
    uint32_t lim = min(seg0 + P.seg_bits, src.rel(P.count_end_bit));
    if (lim > eos) lim = eos;
    uint32_t n = 0, next = 0, nz = 0;
    unsigned int bad = 0;
    SegListSink sink;
    sink.dst = P.seg_list + seg_list_base(k);
    sink.e0 = sink.e1 = sink.e2 = sink.e3 = 0u;
    if (seg0 + entry >= lim) { n = 0; next = seg0 + entry; }
    else if (!eg_scan_segment(src, seg0 + entry, lim, eos, n, next, &nz, sink)) { bad = 0x80000000u; n = 0; nz = 0; next = lim; }
    P.seg_count[k] = n | (nz << 16) | bad;
    P.seg_used[k] = entry;
    const unsigned int over = next > lim ? next - lim : 0u;
    if (P.seg_over[k + 1] != over) { P.seg_over[k + 1] = over; *P.changed = round; }
}

// Exclusive prefix sums over all segments of (a) the code counts and (b) the non-zero code counts:
// 4096 segments per CTA tile (16 per thread; round 2: 4 per thread left the kernel a chain of look-back hops, 19 -> 16 us), block scan, decoupled look-back across tiles (same machinery as the
// bit packer; one status array per quantity).  seg_first[nseg] / seg_nzfirst[nseg] = totals.
#ifndef DCT3D_SCAN_ITEMS
#define DCT3D_SCAN_ITEMS 16
#endif
constexpr int kScanThreads = 256, kScanItems = DCT3D_SCAN_ITEMS;

__global__ void __launch_bounds__(kScanThreads)
seg_prefix_kernel(const DecParams P, unsigned long long *status_codes, unsigned long long *status_nz, unsigned int *ticket)
{
    __shared__ unsigned long long s_wsum[2][kScanThreads / 32];
    __shared__ unsigned long long s_off[2];
    __shared__ long long s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int TILE = kScanThreads * kScanItems;
    const long long ntiles = (long long)((P.nseg + TILE - 1) / TILE);
    for (;;) {
        if (tid == 0) s_tile = (long long)atomicAdd(ticket, 1u);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        const unsigned long long k0 = (unsigned long long)tile * TILE + (unsigned long long)tid * kScanItems;
        unsigned long long v[2][kScanItems], sum[2] = {0, 0}, incl[2];
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            unsigned int x = k0 + i < P.nseg ? P.seg_count[k0 + i] : 0u;
            if (x & 0x80000000u) { atomicOr(P.err, 2u); x = 0; }                  // still malformed after convergence
            v[0][i] = x & 0xffffu;
            v[1][i] = (x >> 16) & 0x7fffu;
            sum[0] += v[0][i];
            sum[1] += v[1][i];
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            incl[q] = sum[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, incl[q], d); if (lane >= d) incl[q] += o; }
            if (lane == 31) s_wsum[q][warp] = incl[q];
        }
        __syncthreads();
        if (warp < 2) {
            const int q = warp;
            const unsigned long long w = lane < kScanThreads / 32 ? s_wsum[q][lane] : 0ull;
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
            if (lane < kScanThreads / 32) s_wsum[q][lane] = wi - w;
            const unsigned long long total = __shfl_sync(0xffffffffu, wi, 31);
            const unsigned long long off = tile_lookback(q == 0 ? status_codes : status_nz, tile, total, q == 0 ? P.code_base : P.nz_base, lane, P.err);
            if (lane == 0) {
                s_off[q] = off;
                if (tile == ntiles - 1) (q == 0 ? P.seg_first : P.seg_nzfirst)[P.nseg] = off + total;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; q++) {
            unsigned long long run = s_off[q] + s_wsum[q][warp] + incl[q] - sum[q];
            unsigned long long *dst = q == 0 ? P.seg_first : P.seg_nzfirst;
#pragma unroll
            for (int i = 0; i < kScanItems; i++) {
                if (k0 + i < P.nseg) dst[k0 + i] = run;
                run += v[q][i];
            }
        }
        __syncthreads();
    }
}

// Segment lists -> rows of a CSR matrix of the cubes.  A warp owns 32 consecutive segments; every lane
// knows its segment's first code index and the rank of its first non-zero code in the whole stream, so
// entry i of segment s goes to coo[rank_s + i] as (list index space of the coefficient << 16 | value).
// The copy is ENTRY-parallel: the warp takes the 32 lists one after the other (four in flight), a lane per
// entry, with the segment's parameters broadcast by shuffle, so the stores are contiguous and sparse and
// dense content keep the lanes busy alike.  Row pointers: each lane bisects its own (sorted) list for every
// cube whose first code lies in its segment.  No bit is parsed a second time; only the one thread that
// holds the clip's last code re-walks its segment to report where the stream ends.
constexpr int kEmitThreads = 128;

template <int C>
__global__ void __launch_bounds__(kEmitThreads, 12)
seg_emit_kernel(const DecParams P)
{
    using G = Geo<C>;
    __shared__ uint16_t s_lin[G::CS];
    for (int i = threadIdx.x; i < G::CS; i += kEmitThreads) s_lin[i] = (C == 8 ? g_zz.slin8 : g_zz.slin4)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t *lists = reinterpret_cast<const uint32_t *>(P.seg_list);
    const unsigned long long ncodes = (unsigned long long)P.L.ncubes * G::CS;
    for (unsigned long long blk = blockIdx.x; blk * kEmitThreads < P.nseg; blk += gridDim.x) {
        const unsigned long long k = blk * (unsigned long long)kEmitThreads + threadIdx.x;
        const unsigned long long kw = k - lane;                  // the warp's first segment
        if (kw >= P.nseg) break;                                 // whole warp beyond the stream
        const bool inside = k < P.nseg;
        const unsigned long long cur = inside ? P.seg_first[k] : ncodes;
        unsigned long long hi = inside ? P.seg_first[k + 1] : ncodes;
        const bool last = cur < ncodes && hi >= ncodes;          // this segment holds the clip's last code
        if (hi > ncodes) hi = ncodes;
        const bool active = cur < hi;
        const uint32_t span = active ? (uint32_t)(hi - cur) : 0u;   // codes of this segment that belong to the clip
        const uint32_t cnt = active ? (P.seg_count[k] >> 16) & 0x7fffu : 0u;
        const unsigned long long zr = inside ? P.seg_nzfirst[k] : 0ull;
        const uint32_t pos0 = (uint32_t)(cur % G::CS);           // position of the first code inside its cube
        // word address of entry i of segment kw + s:  wbase + s * 4 + (i >> 2) * 128 + (i & 3)
        const unsigned long long wbase = (kw >> 5) * (unsigned long long)(kSegListVec * 32 * 4);
        auto entry = [&](int s, uint32_t i) { return __ldg(lists + wbase + (unsigned)s * 4u + (i >> 2) * 128u + (i & 3u)); };

        // first entry with rel >= bound (the list is sorted by rel)
        auto lower_bound = [&](uint32_t bound) {
            uint32_t lo = 0, up = cnt;
            while (lo < up) {
                const uint32_t mid = (lo + up) >> 1;
                if ((entry(lane, mid) >> 17) < bound) lo = mid + 1; else up = mid;
            }
            return lo;
        };
        if (!P.locate_only) {
        // ---- entries: the warp copies one segment's list after the other, a lane per entry, four segments
        // in flight; contiguous stores, and no search for the segment an entry belongs to
        auto emit_one = [&](uint32_t e, uint32_t sp, unsigned long long dst) {
            const uint32_t rel = e >> 17;
            P.coo[dst] = ((uint32_t)s_lin[(sp + rel) & (G::CS - 1)] << 16) | ((uint32_t)eg_unmap(e & 0x1ffffu) & 0xffffu);
        };
#ifndef DCT3D_EMIT_FLIGHT
#define DCT3D_EMIT_FLIGHT 4
#endif
        constexpr int NF = DCT3D_EMIT_FLIGHT;                        // lists in flight per warp
        for (int s0 = 0; s0 < 32; s0 += NF) {
            uint32_t sc[NF], sp[NF], e[NF];
            unsigned long long sz[NF];
#pragma unroll
            for (int q = 0; q < NF; q++) {
                sc[q] = __shfl_sync(0xffffffffu, cnt, s0 + q);
                sp[q] = __shfl_sync(0xffffffffu, pos0, s0 + q);
                sz[q] = __shfl_sync(0xffffffffu, zr, s0 + q);
            }
#pragma unroll
            for (int q = 0; q < NF; q++) e[q] = (uint32_t)lane < sc[q] ? entry(s0 + q, lane) : 0u;
#pragma unroll
            for (int q = 0; q < NF; q++)
                if ((uint32_t)lane < sc[q]) emit_one(e[q], sp[q], sz[q] + lane);
#pragma unroll
            for (int q = 0; q < NF; q++)                             // lists longer than a warp: dense content
                for (uint32_t i = 32 + lane; i < sc[q]; i += 32) emit_one(entry(s0 + q, i), sp[q], sz[q] + i);
        }

        // ---- row pointers: this lane's cubes ------------------------------------------------------
        if (active) {
            unsigned long long cube = cur / G::CS + (pos0 ? 1 : 0);       // next cube to start at or after `cur`
            for (uint32_t nb = pos0 ? (uint32_t)G::CS - pos0 : 0u; nb < span; nb += G::CS)
                P.coo_start[cube++] = zr + (nb ? lower_bound(nb) : 0u);
        }
        }
        if (last) {
            if (!P.locate_only) P.coo_start[P.L.ncubes] = zr + lower_bound(span);
            // where does the clip end?  re-walk this one segment up to its last code
            LocalSource src;
            src.w0 = (P.start_bit >> 5) + k * kSegWords;
            src.fill(P.words, P.nwords);
            const uint32_t seg0 = src.rel(P.start_bit + k * (unsigned long long)P.seg_bits);
            const uint32_t eos = src.rel(P.nbits_total);
            uint32_t lim = min(seg0 + P.seg_bits, src.rel(P.count_end_bit));
            if (lim > eos) lim = eos;
            uint32_t n = 0, next = seg0 + P.seg_over[k];
            if (!eg_scan_segment<LocalSource, NullNzSink, true>(src, seg0 + P.seg_over[k], lim, eos, n, next, nullptr, NullNzSink(), span)) atomicOr(P.err, 2u);
            *P.end_bit = src.w0 * 32ull + next;
        }
    }
}

// Where does code number `target` (counted from the first code of the scanned range) start?  One thread: bisect the
// prefix of the segments' code counts, then walk that one segment.  Used after a count-only pass over a part of a stream
// (distributed index discovery, dct3d_multi_locate); target must be below seg_first[nseg].
__global__ void seg_locate_kernel(const DecParams P, unsigned long long target, unsigned long long *out_bit)
{
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long lo = 0, hi = P.nseg;              // seg_first[lo] <= target < seg_first[hi]
    while (hi - lo > 1) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (P.seg_first[mid] <= target) lo = mid; else hi = mid;
    }
    const unsigned long long k = lo;
    LocalSource src;
    src.w0 = (P.start_bit >> 5) + k * kSegWords;
    src.fill(P.words, P.nwords);
    const uint32_t seg0 = src.rel(P.start_bit + k * (unsigned long long)P.seg_bits);
    const uint32_t eos = src.rel(P.nbits_total);
    uint32_t lim = min(seg0 + P.seg_bits, src.rel(P.count_end_bit));
    if (lim > eos) lim = eos;
    uint32_t n = 0, next = seg0 + P.seg_over[k];
    if (!eg_scan_segment<LocalSource, NullNzSink, true>(src, seg0 + P.seg_over[k], lim, eos, n, next, nullptr, NullNzSink(),
                                                        (uint32_t)(target - P.seg_first[k])))
        atomicOr(P.err, 2u);
    *out_bit = src.w0 * 32ull + next;
}

// non-zero lists -> dense natural-order int16 cubes (dct3d_eg_decode_i16; qcubes zeroed beforehand).
template <int C>
__global__ void __launch_bounds__(kThreads)
coo_scatter_kernel(const DecParams P)
{
    using G = Geo<C>;
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * kWarps;
    for (long long cube = wid; cube < P.L.ncubes; cube += nw) {
        const unsigned long long z0 = P.coo_start[cube];
        const uint32_t n = (uint32_t)min(P.coo_start[cube + 1] - z0, (unsigned long long)G::CS);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t e = P.coo[z0 + i];
            P.qcubes[(size_t)cube * G::CS + coo_swizzle<C>((e >> 16) & (G::CS - 1))] = (int16_t)(e & 0xffffu);
        }
    }
}

// Dequantiser table of both reconstruct kernels, in coo_swizzle() index space:
// max(1, 5(k0+k1+k2)) (Decoder.java:89, decoder.c:54) times S[k0] S[k1] S[k2], the factors the
// un-normalised inverse butterflies expect on their inputs.
template <int C>
__device__ __forceinline__ void build_dequant_table(float *tab, int tid)
{
    for (int i = tid; i < C * C * C; i += kThreads) {
        const int k2 = i % C, k1 = (i / C) % C, k0 = i / (C * C);
        tab[coo_swizzle<C>(i)] = (float)quant_divisor(k0 + k1 + k2) * (lane_scale<C>(k0) * lane_scale<C>(k1) * lane_scale<C>(k2));
    }
}

// Inverse tail shared by both reconstruct kernels: b[k0][k2] holds the dequantised coefficients of
// row-frequency k1 = r.  Inverse butterflies along t, exchange, along y and x, clamp to [0,255],
// truncate (Decoder.java:112, decoder.c:29), store the thread's frame plane.
template <int C, bool ALL_SCALED = false, bool T_DONE = false>
__device__ __forceinline__ void idct_store(float (&b)[C][C], uint8_t *xbuf, int cl, int r, bool valid, const Layout &L,
                                           long long cube, uint8_t *__restrict__ frames, uint32_t colmask = 0xffffffffu)
{
    float a[C][C];
    // ALL_SCALED: b already carries S[k0] S[k1] S[k2]; otherwise S[k1] only and the t stage folds in S[k2]
    if (!T_DONE) { if (ALL_SCALED) inv_t_n<C, float>(b); else inv_t_g<C, float>(b); }
    Xch<C, float>::transpose(xbuf, cl, r, b, a);   // the exchange is its own inverse
    if (T_DONE) inv_yx_n_masked<C, float>(a, colmask); else inv_yx_n<C, float>(a);
    if (!valid) return;
    const int per_slab = L.by * L.bx;
    const int slab = (int)(cube / per_slab);
    const int rem = (int)(cube - (long long)slab * per_slab);
    const int byi = rem / L.bx, bxi = rem - byi * L.bx;
    uint8_t *dst = frames + ((size_t)(slab * C + r) * L.H + byi * C) * L.W + bxi * C;
#pragma unroll
    for (int y = 0; y < C; y++) {
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int x = 0; x < C; x += 4) {
            // clamp to [0,255] and truncate in one saturating conversion each (the otherwise idle
            // conversion pipe), then PRMT the four bytes together
            const uint32_t b0 = f2u8_sat(a[y][x]), b1 = f2u8_sat(a[y][x + 1]);
            const uint32_t b2 = f2u8_sat(a[y][x + 2]), b3 = f2u8_sat(a[y][x + 3]);
            w[x / 4] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
        }
        if (C == 8) *reinterpret_cast<uint2 *>(dst + (size_t)y * L.W) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.W) = w[0];
    }
}

// The same tail with the pixels leaving by TMA (reconstruct_coo_kernel<C, true>): after the y and x passes thread
// (cube cl, frame r) packs its C rows into the warp's unit tile [y][t][32 px] in shared memory (the image of ONE box
// {32 px, C frames, C rows} of the tensor {W, F, H}, the encoder's unit read backwards), and lane 0 hands the tile to the
// TMA unit.  Why: the row stores of idct_store are 8 bytes per lane to 8 different frames, 16 data-pipe wavefronts per
// STG.64 and 128 per unit, a third of all wavefronts of a kernel whose LSU data pipe is 87% busy; the tile costs 16
// (C = 8: 8 STS.64 of 256 contiguous bytes) and the TMA unit reads shared memory beside the LSU.  The tile lives in the
// exchange buffer, which is free between the exchange of this unit and the exchange of the next one; the store's read
// of it is awaited (long finished) just before that next exchange.  Needs groups = units, i.e. bx % CPW == 0.
// a[y][x]: the finished plane of thread (cube cl, frame r).  Clamp, truncate, pack into the warp's tile, hand it to TMA.
template <int C>
__device__ __forceinline__ void tile_store(float (&a)[C][C], uint8_t *xbuf, int cl, int r, int lane, const CUtensorMap *tmap,
                                           const UnitPos &pos)
{
    // the last exchange round ended with __syncwarp(): every lane holds its vectors, the buffer is free.
    // The tile has the encoder's SWIZZLE_32B image (unit_offset): the 16-byte halves of rows 4..7 of every 8 are exchanged,
    // so the 16 lanes a 64-bit store is served for (two cubes x 8 frames: half a row each) cover all 32 banks once.  (Linear,
    // rows 0..3 and 4..7 met in the same banks: 4 wavefronts per STS.64 instead of 2, measured.)
#pragma unroll
    for (int y = 0; y < C; y++) {
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int x = 0; x < C; x += 4) {
            const uint32_t b0 = f2u8_sat(a[y][x]), b1 = f2u8_sat(a[y][x + 1]);
            const uint32_t b2 = f2u8_sat(a[y][x + 2]), b3 = f2u8_sat(a[y][x + 3]);
            w[x / 4] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
        }
        uint8_t *dst = xbuf + unit_offset<C>(y, r, cl * C);        // [y][t = r][x = cl * C ..]
        if (C == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint32_t *>(dst) = w[0];
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) tma_store_3d(tmap, xbuf, pos.bxu * kUnitW, pos.slab * C, pos.byi * C);
}

template <int C>
__device__ __forceinline__ void idct_tile_store(float (&b)[C][C], uint8_t *xbuf, int cl, int r, int lane, uint32_t colmask,
                                                const CUtensorMap *tmap, const UnitPos &pos)
{
    float a[C][C];
    if (lane == 0) tma_store_wait_read();          // the previous unit's tile has been read
    __syncwarp();
    Xch<C, float>::transpose(xbuf, cl, r, b, a);
    inv_yx_n_masked<C, float>(a, colmask);
    tile_store<C>(a, xbuf, cl, r, lane, tmap, pos);
}

// Non-zero lists -> u8 frames (the decoder's inverse kernel).  Per warp and group of CPW cubes:
// C lanes per cube scatter the cube's entries, ALREADY DEQUANTISED (one multiply by the table entry
// max(1,5(k0+k1+k2)) * S[k0] S[k1] S[k2], so that all three butterfly passes run un-normalised), into
// a natural-order float cube in shared memory that is kept all-zero (the same lanes wipe their
// entries afterwards); every thread then reads the C rows (k0, k1 = lane) it needs with LDS.128,
// conflict-free.  Only the 3% non-zero coefficients are ever converted or multiplied.  Counts and the
// first 3*C entries of the NEXT group are prefetched into registers while the current group is
// transformed.
// Occupancy note (measured, round 2): the kernel runs 4 CTAs = 16 warps per SM at 128 registers.  Buying a fifth or sixth
// CTA with fewer registers (96: 16 bytes of spills, 80: 144 bytes) and with the exchange buffer aliased onto the cubes made
// it SLOWER (388 -> 417 us for the aliasing alone, 442 us at 5 CTAs, 528 us at 6): the shared-memory pipe and the issue slots
// are co-limiters, so extra wipes and spill traffic cost more than the extra warps hide.  Allowing fewer CTAs changes nothing:
// the compiler does not want more than 128 registers here.
#ifndef DCT3D_COO_PRE
#define DCT3D_COO_PRE 3
#endif
// reconstruct_coo_kernel: a cube with more than PRE * C + kWipeAllOver list entries is wiped as a whole after the t pass
// instead of entry by entry (one pass of the entry loop serves 4 * C entries)
constexpr int kWipeAllOver = 32;
#ifndef DCT3D_NAT_PAD
#define DCT3D_NAT_PAD 4
#endif
template <int C>
struct CooSmem {
    using G = Geo<C>;
    // The cubes of a warp sit DCT3D_NAT_PAD floats further apart than their size: the same coefficient of the CPW cubes (the
    // lanes of the cubes scatter their lists in step: DC first) then lies in CPW different banks instead of one (round 2:
    // 29 of the 38 wavefronts of the scatter and wipe stores of a unit were bank conflicts).  A multiple of 4 floats keeps
    // the rows 16-byte aligned and the row loads of a quarter warp (one cube) conflict-free.
    static constexpr int CUBE_STRIDE = G::CS + DCT3D_NAT_PAD;     // floats
    static constexpr int NAT_WARP = (G::CPW * CUBE_STRIDE * 4 + 255) / 256 * 256;   // bytes: CPW natural-order float cubes; the exchange
                                                                  // buffer behind them stays 256-byte aligned (TMA tile)
    static constexpr int WARP_BYTES = NAT_WARP + Xch<C, float>::WARP_BYTES;
    static constexpr int TAB_OFF = kWarps * WARP_BYTES;            // float[CS] dequantiser table
    static constexpr int TOTAL = TAB_OFF + G::CS * 4;
};

// TAIL: 0 = row stores (idct_store), 1 = TMA tile store (idct_tile_store).
// Measured and dropped in round 2 (profiles/r2_runs/ab_variants_s2.jsonl): "column classes" -- units whose non-zero k2
// columns all lie in 0..3 (25% of the benchmark clip) or 0..5 (65%) taking tails that neither load, exchange nor transform
// the columns known to be zero (dct8_inv_n with four / six live inputs, an exchange round of 8-byte vectors).  Three copies
// of the tail made the kernel 3248 instructions = 52 KB, beyond the 32 KB instruction cache: 405 us against 348; one tail
// with a warp-uniform branch around the upper half: 352 us against 346 (the branches cost what the skipped work saves).
constexpr int TAIL_ROWS = 0, TAIL_TMA = 1;

template <int C, int TAIL = TAIL_ROWS>
__global__ void __launch_bounds__(kThreads, 4)
reconstruct_coo_kernel(const __grid_constant__ CUtensorMap tmap_out, const Layout L, const uint32_t *__restrict__ coo,
                       const unsigned long long *__restrict__ coo_start_all, uint8_t *__restrict__ frames, const long long cube_base,
                       const unsigned long long coo_limit)
{
    // TMA_OUT: the pixels leave as one TMA box per unit (idct_tile_store; tmap_out is the tensor map of `frames`), else by
    // row stores (idct_store; tmap_out is not looked at)
    constexpr bool TMA_OUT = TAIL != TAIL_ROWS;
    // coo_limit: row pointers are clamped to it, so that a launch that runs ahead of the host's look at the control
    // block (lists of a stream whose index discovery has not converged, or of a damaged stream) stays inside the buffer
    // cube_base: the launch reconstructs cubes [cube_base, cube_base + L.ncubes) of the parsed stream into a frame buffer
    // that starts at the first of them (a whole number of slabs): the pipelined decoder's chunks
    const unsigned long long *__restrict__ coo_start = coo_start_all + cube_base;
    using G = Geo<C>;
    using S = CooSmem<C>;
    constexpr int PRE = DCT3D_COO_PRE;                              // prefetched entries per lane
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cl = lane / C, r = lane % C;
    float *nat = reinterpret_cast<float *>(smem + warp * S::WARP_BYTES) + cl * S::CUBE_STRIDE;   // this thread's cube
    uint8_t *xbuf = smem + warp * S::WARP_BYTES + S::NAT_WARP;
    float *tab = reinterpret_cast<float *>(smem + S::TAB_OFF);
    build_dequant_table<C>(tab, tid);
    for (int i = lane; i < S::NAT_WARP / 16; i += 32) reinterpret_cast<uint4 *>(smem + warp * S::WARP_BYTES)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // group counts fit 31 bits with room for two strides (checked by the host): 32-bit loop state, three registers less
    const int ngroups = (int)((L.ncubes + G::CPW - 1) / G::CPW);
    const int stride = (int)gridDim.x * kWarps;
    int g = (int)blockIdx.x * kWarps + warp;
    // pipeline: row pointers two groups ahead, the first PRE*C entries one group ahead
    unsigned long long z_nn = 0, z_n = 0;      // first entry of this thread's cube, groups g+2 and g+1
    uint32_t zend_nn = 0, c_n = 0, e_n[PRE];   // low word of the next row pointer: the count is formed a pass later
    auto fetch_rows = [&](int grp) {
        const long long cube = (long long)grp * G::CPW + cl;
        const bool ok = grp < ngroups && cube < L.ncubes;
        // raw values: nothing may consume the loads here, a pass before they are needed (round 2: the clamp sat right behind
        // the load and cost a full L2 round trip per pass, 3% of the kernel's samples)
        z_nn = ok ? __ldg(coo_start + cube) : 0ull;
        zend_nn = ok ? __ldg(reinterpret_cast<const uint32_t *>(coo_start + cube + 1)) : 0u;
    };
    auto rotate_rows = [&]() {
        z_n = min(z_nn, coo_limit);
        c_n = min(zend_nn - (uint32_t)z_nn, (uint32_t)G::CS);
    };
    auto fetch_entries = [&]() {               // for the group whose row pointers are in z_n / c_n
#pragma unroll
        for (int k = 0; k < PRE; k++) e_n[k] = (uint32_t)(r + k * C) < c_n ? __ldg(coo + z_n + r + k * C) : 0u;
    };
    auto put = [&](uint32_t x) {
        const uint32_t idx = (x >> 16) & (G::CS - 1);
        nat[idx] = (float)(int)(int16_t)(x & 0xffffu) * tab[idx];
    };
    fetch_rows(g);
    rotate_rows();
    fetch_entries();
    fetch_rows(g + stride);
    // TMA_OUT: group g is unit g of the encoder's numbering (the host checked bx % CPW == 0); the unit counts fit 31 bits
    UnitPos upos = unit_pos(L, TMA_OUT ? (long long)min(g, ngroups) : 0ll);
    const UnitPos ustep = unit_pos(L, TMA_OUT ? (long long)stride : 0ll);
    for (; g < ngroups; g += stride) {
        const uint32_t cnt = c_n;
        const unsigned long long z0 = z_n;
        uint32_t e[PRE];
#pragma unroll
        for (int k = 0; k < PRE; k++) e[k] = e_n[k];
        const long long cube = (long long)g * G::CPW + cl;
        rotate_rows();
        fetch_entries();                        // group g + stride
        fetch_rows(g + 2 * stride);
        // scatter this cube's entries (lane r of the cube takes entries r, r+C, r+2C, ...)
        uint32_t colbits = 0;                   // k2 columns this lane scattered into (physical = swizzled position)
        {
            // the table loads of the prefetched entries are issued together (index 0 for an absent entry)
            uint32_t ix[PRE];
            float tv[PRE];
#pragma unroll
            for (int k = 0; k < PRE; k++) { ix[k] = (e[k] >> 16) & (G::CS - 1); tv[k] = tab[ix[k]]; }
#pragma unroll
            for (int k = 0; k < PRE; k++) tv[k] *= (float)(int)(int16_t)(e[k] & 0xffffu);
#pragma unroll
            for (int k = 0; k < PRE; k++)
                if ((uint32_t)(r + k * C) < cnt) { nat[ix[k]] = tv[k]; colbits |= 1u << (coo_swizzle<C>(ix[k]) & (C - 1)); }
        }
        // dense cubes: the rest of the list, DENSE_LD independent loads at a time (the transform's registers are not live
        // here; on noise content, 64 entries per lane, the loop is a chain of L2 round trips: 8 in flight instead of 4
        // halves their number)
        constexpr int DENSE_LD = 8;
        for (uint32_t i = r + PRE * C; i < cnt; i += DENSE_LD * C) {
            uint32_t x[DENSE_LD];
#pragma unroll
            for (int k = 0; k < DENSE_LD; k++) x[k] = i + k * C < cnt ? __ldg(coo + z0 + i + k * C) : 0u;
#pragma unroll
            for (int k = 0; k < DENSE_LD; k++)
                if (i + k * C < cnt) put(x[k]);
        }
        if (cnt > (uint32_t)(PRE * C)) colbits = (1u << C) - 1u;   // a dense cube: no pruning, and no bookkeeping in its loop
        // columns that are zero in all cubes of this pass skip their t and y transforms
        const uint32_t colmask = __reduce_or_sync(0xffffffffu, colbits);
        __syncwarp();
        // wipe what was scattered (after the t pass: by now every lane's row loads have long landed)
        auto wipe = [&]() {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < PRE; k++)
                if ((uint32_t)(r + k * C) < cnt) nat[(e[k] >> 16) & (G::CS - 1)] = 0.0f;
            if (cnt > (uint32_t)(PRE * C + kWipeAllOver)) {
                // a dense cube is wiped as a whole -- CS / C floats per lane as interleaved 16-byte stores, no list re-read
                // (round 2, measured on noise: the re-read of 64 entries per lane was a second chain of L2 round trips)
#pragma unroll
                for (int i = 0; i < G::CS / (4 * C); i++) reinterpret_cast<float4 *>(nat)[i * C + r] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (uint32_t i = r + PRE * C; i < cnt; i += 4 * C) {
                    uint32_t x[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) x[k] = i + k * C < cnt ? __ldg(coo + z0 + i + k * C) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (i + k * C < cnt) nat[(x[k] >> 16) & (G::CS - 1)] = 0.0f;
                }
            }
        };
        float b[C][C];
        // first halves of all rows, then second halves: the t pass of columns 0..3 starts while the
        // second halves are still in flight
#pragma unroll
        for (int k2 = 0; k2 < C; k2 += 4) {
#pragma unroll
            for (int k0 = 0; k0 < C; k0++) {
                const float *row = nat + (k0 * C + r) * C;
                const float4 v = *reinterpret_cast<const float4 *>(row + (C == 8 ? k2 ^ (r & 4) : k2));   // coo_swizzle
                b[k0][k2] = v.x; b[k0][k2 + 1] = v.y; b[k0][k2 + 2] = v.z; b[k0][k2 + 3] = v.w;
            }
        }
        inv_t_n_masked<C, float>(b, colmask);
        wipe();
        if (TMA_OUT) {
            idct_tile_store<C>(b, xbuf, cl, r, lane, colmask, &tmap_out, upos);
            unit_advance(L, upos, ustep);
        } else {
            idct_store<C, true, true>(b, xbuf, cl, r, cube < L.ncubes, L, cube, frames, colmask);
        }
    }
    if (TMA_OUT && lane == 0) tma_store_wait_read();   // the last tile must have been read before the CTA's shared memory goes
}

// int16 natural-order cubes -> u8 frames: dequantise, inverse butterflies, clamp, truncate.
// (reference Decoder.java:78-117, decoder.c:48-59 + 3dDCT.cl:164-265 + decoder.c:29)
template <int C>
__global__ void __launch_bounds__(kThreads, 4)
reconstruct_kernel(const Layout L, const int16_t *__restrict__ qcubes, uint8_t *__restrict__ frames)
{
    using G = Geo<C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cl = lane / C, r = lane % C;
    float *tab = reinterpret_cast<float *>(smem + kWarps * Xch<C, float>::WARP_BYTES);   // float[CS]
    build_dequant_table<C>(tab, tid);
    __syncthreads();
    const long long ngroups = (L.ncubes + G::CPW - 1) / G::CPW;
    for (long long g = (long long)blockIdx.x * kWarps + warp; g < ngroups; g += (long long)gridDim.x * kWarps) {
        const long long cube = g * G::CPW + cl;
        const bool valid = cube < L.ncubes;
        float b[C][C];
        // thread = k1 = r: rows (k0, k1) of the cube, C int16 each
        const int16_t *src = qcubes + (size_t)(valid ? cube : 0) * G::CS + r * C;
#pragma unroll
        for (int k0 = 0; k0 < C; k0++) {
            uint32_t w[4] = {0, 0, 0, 0};
            if (valid) {
                if (C == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + k0 * C * C)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
                else { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src + k0 * C * C)); w[0] = v.x; w[1] = v.y; }
            }
            const float *trow = tab + (k0 * C + r) * C;
#pragma unroll
            for (int k2 = 0; k2 < C; k2 += 4) {
                const float4 m = *reinterpret_cast<const float4 *>(trow + (C == 8 ? k2 ^ (r & 4) : k2));   // coo_swizzle
                const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int q = (int)(int16_t)((w[(k2 + e) / 2] >> (((k2 + e) & 1) * 16)) & 0xffffu);
                    b[k0][k2 + e] = (float)q * mm[e];     // the same product reconstruct_coo_kernel forms at scatter time
                }
            }
        }
        idct_store<C, true>(b, smem + warp * Xch<C, float>::WARP_BYTES, cl, r, valid, L, cube, frames);
    }
}

// ------------------------------------------------------------------------------------------
// Transform-only seams.
//   CUBEMAJOR = true : T cube-major in, T cube-major out (the C codec's device boundary)
//   CUBEMAJOR = false: T planar [F][H][W] in and out      (Java's Transform boundary)
// ------------------------------------------------------------------------------------------
// C contiguous elements of T as 16-byte vectors (rows are 16-byte aligned in both layouts).
__device__ __forceinline__ void ld_global_256(const void *p, uint4 &a, uint4 &b)
{
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

template <int C, typename T>
__device__ __forceinline__ void load_row(const T *p, T (&row)[C], bool valid)
{
    constexpr int NV = C * sizeof(T) / 16, VEC = 16 / sizeof(T);
    if (NV % 2 == 0 && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; i += 2) {
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
            if (valid) ld_global_256(reinterpret_cast<const uint4 *>(p) + i, v0, v1);
            const T *p0 = reinterpret_cast<const T *>(&v0), *p1 = reinterpret_cast<const T *>(&v1);
#pragma unroll
            for (int e = 0; e < VEC; e++) { row[i * VEC + e] = p0[e]; row[(i + 1) * VEC + e] = p1[e]; }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid) v = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        const T *pv = reinterpret_cast<const T *>(&v);
#pragma unroll
        for (int e = 0; e < VEC; e++) row[i * VEC + e] = pv[e];
    }
}
// one 256-bit store (sm_100: st.global.v8.b32) writes a whole 32-byte sector from a single lane
__device__ __forceinline__ void st_global_256(void *p, const uint4 &a, const uint4 &b)
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

template <int C, typename T>
__device__ __forceinline__ void store_row(T *p, const T (&row)[C])
{
    constexpr int NV = C * sizeof(T) / 16, VEC = 16 / sizeof(T);
    if (NV % 2 == 0 && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; i += 2) {
            uint4 v0, v1;
            T *p0 = reinterpret_cast<T *>(&v0), *p1 = reinterpret_cast<T *>(&v1);
#pragma unroll
            for (int e = 0; e < VEC; e++) { p0[e] = row[i * VEC + e]; p1[e] = row[(i + 1) * VEC + e]; }
            st_global_256(reinterpret_cast<uint4 *>(p) + i, v0, v1);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
        uint4 v;
        T *pv = reinterpret_cast<T *>(&v);
#pragma unroll
        for (int e = 0; e < VEC; e++) pv[e] = row[i * VEC + e];
        reinterpret_cast<uint4 *>(p)[i] = v;
    }
}

template <int C, typename T, bool CUBEMAJOR, bool INVERSE>
__global__ void __launch_bounds__(kThreads)
transform_kernel(const Layout L, const T *__restrict__ in, T *__restrict__ out)
{
    using G = Geo<C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cl = lane / C, r = lane % C;
    const long long ngroups = (L.ncubes + G::CPW - 1) / G::CPW;
    const size_t fs = (size_t)L.W * L.H;
    for (long long g = (long long)blockIdx.x * kWarps + warp; g < ngroups; g += (long long)gridDim.x * kWarps) {
        const long long cube = g * G::CPW + cl;
        const bool valid = cube < L.ncubes;
        const long long cc = valid ? cube : 0;
        const int per_slab = L.by * L.bx;
        const int slab = (int)(cc / per_slab);
        const int rem = (int)(cc - (long long)slab * per_slab);
        const int byi = rem / L.bx, bxi = rem - byi * L.bx;
        // element (i0 = frame / k0, i1 = row / k1, i2 = col / k2) of this cube
        auto addr = [&](int i0, int i1, int i2) -> size_t {
            return CUBEMAJOR ? (size_t)cc * G::CS + (size_t)(i0 * C + i1) * C + i2
                             : (size_t)(slab * C + i0) * fs + (size_t)(byi * C + i1) * L.W + bxi * C + i2;
        };
        T a[C][C], b[C][C];
        if (!INVERSE) {
            // t pass first (thread = row y holds [t][x]): the C threads of a cube then load one contiguous
            // run per frame, the same access shape as the inverse direction, which scattered row loads by
            // frame do not reach (5.5 vs 6.0 TB/s); the transform is separable, so the order is free
#pragma unroll
            for (int t = 0; t < C; t++) load_row<C, T>(in + addr(t, r, 0), b[t], valid);
            fwd_t<C, T>(b);
            Xch<C, T>::transpose(smem + warp * Xch<C, T>::WARP_BYTES, cl, r, b, a);   // thread = k0 holds [y][x]
            fwd_xy<C, T>(a);
            if (valid) {
#pragma unroll
                for (int k1 = 0; k1 < C; k1++) store_row<C, T>(out + addr(r, k1, 0), a[k1]);
            }
        } else {
#pragma unroll
            for (int k0 = 0; k0 < C; k0++) load_row<C, T>(in + addr(k0, r, 0), b[k0], valid);
            inv_t<C, T>(b);
            Xch<C, T>::transpose(smem + warp * Xch<C, T>::WARP_BYTES, cl, r, b, a);
            inv_yx<C, T>(a);
            if (valid) {
#pragma unroll
                for (int y = 0; y < C; y++) {
#pragma unroll
                    for (int x = 0; x < C; x++) {
                        const T v = a[y][x];
                        a[y][x] = v > (T)255 ? (T)255 : (v < (T)0 ? (T)0 : v);   // InverseDCT.java:74-80, 3dDCT.cl:255-262
                    }
                    store_row<C, T>(out + addr(r, y, 0), a[y]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp64 mode (option "precision" = 64): the reference's Java flavour computes in double (DCT.java:41-59,
// Encoder.java:82, Decoder.java:89,112).  One fused kernel per direction: u8 frames -> double butterflies ->
// the reference's quantiser -> natural-order int16 cubes, and back; no planar doubles in HBM (3 bytes of traffic
// per sample instead of 34).  Same operations in the same order as the f64 transform seam (transform_kernel<C,
// double, false, .>), so the quantised cubes match the fp64 oracle without rounding-tie flips.
//   forward: q = Math.round(c / d) = floor(c / d + 0.5) (Encoder.java:82; rounding 0) or C round(), ties away from
//            zero (encoder.c:53; rounding 1), d = max(1, 5(k0+k1+k2));
//   inverse: c = q * d (Decoder.java:89), clamp to [0,255] (InverseDCT.java:74-80), (byte)(double) truncates
//            (Decoder.java:112).
// ------------------------------------------------------------------------------------------
// fp64 <-> integer without the conversion pipe (I2F.F64 / F2I.F64 / FRND.F64 run at a fraction of the DADD rate and were
// 60% of the first version of these kernels): integers travel through the mantissa of 2^52-scaled doubles.
constexpr double kMagic52 = 4503599627370496.0;            // 2^52
constexpr double kMagicRn = 6755399441055744.0;            // 1.5 * 2^52: (x + kMagicRn) has round-to-nearest-even(x) in its low word
__device__ __forceinline__ double u32_to_f64(uint32_t x) { return __hiloint2double(0x43300000, (int)x) - kMagic52; }           // exact
__device__ __forceinline__ double i32_to_f64(int x) { return __hiloint2double(0x43300000, x ^ (int)0x80000000) - (kMagic52 + 2147483648.0); }
// floor(t) for |t| < 2^31 as an int, and as a double in *fl
__device__ __forceinline__ int floor_f64(double t, double *fl)
{
    const double qq = __dadd_rn(t, kMagicRn);
    const double rn = __dadd_rn(qq, -kMagicRn);            // nearest integer, ties to even
    const bool over = rn > t;
    *fl = over ? rn - 1.0 : rn;
    return __double2loint(qq) - (over ? 1 : 0);
}

// The reference's quantiser as written: round(c / d) with a true division (Encoder.java:82 / encoder.c:53).  Out of line on
// purpose: it is only taken next to a rounding tie, and 64 inlined copies of the division sequence made the kernel
// three times larger than the instruction cache likes.
__device__ __noinline__ int quant_exact_f64(double c, int s, int rounding)
{
    const double v = c / (double)quant_divisor(s);
    return (int)(rounding ? (v < 0 ? -floor(-v + 0.5) : floor(v + 0.5)) : floor(v + 0.5));
}

template <int C, bool INVERSE>
__global__ void __launch_bounds__(kThreads)
codec_f64_kernel(const Layout L, const uint8_t *__restrict__ frames_in, int16_t *__restrict__ q_out,
                 const int16_t *__restrict__ q_in, uint8_t *__restrict__ frames_out, int rounding)
{
    using G = Geo<C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cl = lane / C, r = lane % C;
    const long long ngroups = (L.ncubes + G::CPW - 1) / G::CPW;
    const size_t fs = (size_t)L.W * L.H;
    uint8_t *xbuf = smem + warp * Xch<C, double>::WARP_BYTES;
    // forward: 1 / max(1, 5 s); inverse: max(1, 5 s); s = k0 + k1 + k2
    double *s_inv = reinterpret_cast<double *>(smem + kWarps * Xch<C, double>::WARP_BYTES);
    if (tid < 3 * C - 2) s_inv[tid] = INVERSE ? (double)quant_divisor(tid) : 1.0 / (double)quant_divisor(tid);
    __syncthreads();
    for (long long g = (long long)blockIdx.x * kWarps + warp; g < ngroups; g += (long long)gridDim.x * kWarps) {
        const long long cube = g * G::CPW + cl;
        const bool valid = cube < L.ncubes;
        const long long cc = valid ? cube : 0;
        const int per_slab = L.by * L.bx;
        const int slab = (int)(cc / per_slab);
        const int rem = (int)(cc - (long long)slab * per_slab);
        const int byi = rem / L.bx, bxi = rem - byi * L.bx;
        const size_t pix0 = (size_t)(slab * C) * fs + (size_t)(byi * C) * L.W + (size_t)bxi * C;   // pixel (t = 0, y = 0, x = 0) of the cube
        double a[C][C], b[C][C];
        if (!INVERSE) {
            // thread = row y: b[t][x]; t pass, exchange, thread = k0: a[y][x]; y and x passes; quantise
#pragma unroll
            for (int t = 0; t < C; t++) {
                const uint8_t *src = frames_in + pix0 + (size_t)t * fs + (size_t)r * L.W;
                uint32_t w[2] = {0, 0};
                if (valid) {
                    if (C == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src)); w[0] = v.x; w[1] = v.y; }
                    else w[0] = __ldg(reinterpret_cast<const uint32_t *>(src));
                }
#pragma unroll
                for (int x = 0; x < C; x++) b[t][x] = u32_to_f64((w[x / 4] >> (8 * (x % 4))) & 0xffu);
            }
            fwd_t<C, double>(b);
            Xch<C, double>::transpose(xbuf, cl, r, b, a);
            fwd_xy<C, double>(a);
            if (valid) {
                int16_t *dst = q_out + (size_t)cube * G::CS + (size_t)r * C * C;
#pragma unroll
                for (int k1 = 0; k1 < C; k1++) {
                    uint32_t w[C / 2];
#pragma unroll
                    for (int k2 = 0; k2 < C; k2++) {
                        // Division in double is a long software sequence: multiply by the tabulated inverse instead, and
                        // let the reference's own quotient decide only where the product is within 1e-9 of a rounding tie
                        // (the product is within 1e-11 of the quotient, so everywhere else both round the same way)
                        const int s = r + k1 + k2;
                        const double v = a[k1][k2] * s_inv[s];
                        double q;
                        int qi = floor_f64(v + 0.5, &q);
                        const double fr = (v + 0.5) - q;
                        if (fr < 1e-9 || fr > 1.0 - 1e-9) qi = quant_exact_f64(a[k1][k2], s, rounding);
                        // |q| <= 255 * sqrt(C^3) = 5770 for u8 input: int16 holds it without a clamp
                        const uint32_t h = (uint32_t)qi & 0xffffu;
                        if (k2 & 1) w[k2 / 2] |= h << 16; else w[k2 / 2] = h;
                    }
                    if (C == 8) *reinterpret_cast<uint4 *>(dst + k1 * C) = make_uint4(w[0], w[1], w[2 % (C / 2)], w[3 % (C / 2)]);
                    else *reinterpret_cast<uint2 *>(dst + k1 * C) = make_uint2(w[0], w[1]);
                }
            }
        } else {
            // thread = k1: b[k0][k2] = q * d; t pass, exchange, thread = frame t: a[y][x]; y and x passes; clamp, truncate
            const int16_t *src = q_in + (size_t)cc * G::CS + (size_t)r * C;
#pragma unroll
            for (int k0 = 0; k0 < C; k0++) {
                uint32_t w[4] = {0, 0, 0, 0};
                if (valid) {
                    if (C == 8) { const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + k0 * C * C)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
                    else { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src + k0 * C * C)); w[0] = v.x; w[1] = v.y; }
                }
#pragma unroll
                for (int k2 = 0; k2 < C; k2++) {
                    const int q = (int)(int16_t)((w[k2 / 2] >> ((k2 & 1) * 16)) & 0xffffu);
                    b[k0][k2] = i32_to_f64(q) * s_inv[k0 + r + k2];
                }
            }
            inv_t<C, double>(b);
            Xch<C, double>::transpose(xbuf, cl, r, b, a);
            inv_yx<C, double>(a);
            if (valid) {
                uint8_t *dst = frames_out + pix0 + (size_t)r * fs;
#pragma unroll
                for (int y = 0; y < C; y++) {
                    uint32_t w[2] = {0, 0};
#pragma unroll
                    for (int x = 0; x < C; x++) {
                        const double v = a[y][x];
                        const double c = v > 255.0 ? 255.0 : (v < 0.0 ? 0.0 : v);
                        double unused;
                        w[x / 4] |= ((uint32_t)floor_f64(c, &unused) & 0xffu) << (8 * (x % 4));   // c >= 0: truncation is floor
                    }
                    if (C == 8) *reinterpret_cast<uint2 *>(dst + (size_t)y * L.W) = make_uint2(w[0], w[1]);
                    else *reinterpret_cast<uint32_t *>(dst + (size_t)y * L.W) = w[0];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Colour planes: the reference codes colour video as three gray streams and
// converts with RGBUtils (RGBUtils.java:39-92 split, :94-131 mix): byte i of the raw RGB24 file
// belongs to plane i % 3.  One thread moves 16 pixels: 3 x 16 bytes one way, 48 contiguous bytes the
// other, shuffled with PRMT; pure HBM traffic (2 B moved per byte).  Tails and unaligned buffers take
// the byte path.
// ------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256)
rgb_planes_kernel(uint8_t *__restrict__ rgb, uint8_t *__restrict__ p0, uint8_t *__restrict__ p1, uint8_t *__restrict__ p2,
                  unsigned long long nbytes, int vec_ok)
{
    const unsigned long long ngroups = vec_ok ? nbytes / 48 : 0;     // 16 pixels each
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long nthr = gridDim.x * (unsigned long long)blockDim.x;
    for (unsigned long long g = tid; g < ngroups; g += nthr) {
        uint32_t w[12], o[3][4];
        if (SPLIT) {
            const uint4 *src = reinterpret_cast<const uint4 *>(rgb + g * 48);
#pragma unroll
            for (int i = 0; i < 3; i++) { const uint4 v = __ldg(src + i); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    // output word q of plane c = bytes 12q + c + {0, 3, 6, 9} of the group
                    const int b0 = 12 * q + c;
                    const uint32_t lo = __byte_perm(w[b0 / 4], w[(b0 + 3) / 4], (b0 % 4) | ((((b0 + 3) % 4) + 4) << 4));
                    const uint32_t hi = __byte_perm(w[(b0 + 6) / 4], w[(b0 + 9) / 4], ((b0 + 6) % 4) | ((((b0 + 9) % 4) + 4) << 4));
                    o[c][q] = __byte_perm(lo, hi, 0x5410);
                }
            uint8_t *dst[3] = {p0, p1, p2};
#pragma unroll
            for (int c = 0; c < 3; c++) *reinterpret_cast<uint4 *>(dst[c] + g * 16) = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
        } else {
            const uint8_t *src[3] = {p0, p1, p2};
#pragma unroll
            for (int c = 0; c < 3; c++) { const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src[c] + g * 16)); o[c][0] = v.x; o[c][1] = v.y; o[c][2] = v.z; o[c][3] = v.w; }
#pragma unroll
            for (int j = 0; j < 12; j++) {
                // output word j = bytes 4j .. 4j+3 of the group; byte b is pixel b / 3 of plane b % 3
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int b = 4 * j + k, px = b / 3, c = b % 3;
                    v |= ((o[c][px / 4] >> (8 * (px % 4))) & 0xffu) << (8 * k);
                }
                w[j] = v;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(rgb + g * 48);
#pragma unroll
            for (int i = 0; i < 3; i++) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
    }
    // bytes the vector path did not cover
    uint8_t *pl[3] = {p0, p1, p2};
    for (unsigned long long i = ngroups * 48 + tid; i < nbytes; i += nthr) {
        if (SPLIT) pl[i % 3][i / 3] = rgb[i]; else rgb[i] = pl[i % 3][i / 3];
    }
}

}  // namespace dct3d
