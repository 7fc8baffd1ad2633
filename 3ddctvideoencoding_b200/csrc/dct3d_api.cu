// dct3d_api.cu -- the C ABI of libdct3d.so (include/dct3d.h) over the kernels in
// dct3d_kernels.cuh.  CUDA only: there is no CPU code path behind these entry points.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dct3d.h"
#include "dct3d_kernels.cuh"

using namespace dct3d;

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kKevRing = 32;

struct Ctrl {               // small device control block, zeroed before every launch
    unsigned int ticket;
    unsigned int err;
    unsigned int changed;
    unsigned int pad_;
    unsigned long long end_bit;
};

}  // namespace

struct dct3d_ctx {
    int device = 0, W = 0, H = 0, C = 8;
    int num_sms = 0;
    int use_tma = 1;
    int zero_skip = 1;               // option: the fused 8^3 encoder tests groups of high diagonals for zero before quantising them
    int tma_store = 1;               // option: the inverse kernel stores its pixel tiles by TMA (needs width % 32 == 0)
    int pack_sort = 1;               // option: the bit packer deals the cubes of a tile to its threads in order of their chunk counts
    int precision = 32;              // option: 64 = the fused entry points compute in fp64 (Java parity)
    int rounding = 0;                // option, fp64 mode: 0 = Math.round (floor(v+0.5)), 1 = C round()
    int debug = 0;
    int tma_store_used = 0;          // statistic: the last inverse-kernel launch stored by TMA
    int reuse_zeroed = 0;            // option: see dct3d_set_option
    const void *clean_ptr = nullptr; // stream buffer known to be zero beyond clean_dirty bytes
    size_t clean_cap = 0, clean_dirty = 0;
    int occ_cache[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // cached occupancy / attribute set-up per kernel variant
    long launches = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;      // side stream: the stream wipe of the fused encoder runs beside kernel 1
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    DevBuf frames, bits, q, ctrl, seg, seglist, fa, fb, zz, cmask, coo, coocnt;
    // CUDA events around the last kKevRing launches of encode_kernel [0] / reconstruct_coo_kernel [1]: the kernels' device
    // times can be read after a run of calls without synchronising inside it (statistics ns_*_kernel)
    cudaEvent_t kev[2][kKevRing][2] = {};
    unsigned long kcalls[2] = {0, 0};
    cudaEvent_t ev_ctrl = nullptr;   // the control block of a decode has reached the host
    Ctrl *h_ctrl = nullptr;          // pinned
    unsigned long long *h_u64 = nullptr;  // pinned scratch (4 entries)
    // streaming state
    uint8_t carry_byte = 0;
    int carry_bits = 0;
    // pipelined host-buffer paths (dct3d_encode_u8 / dct3d_decode_u8 and the range calls)
    int chunk_frames = 0;            // option: frames per pipeline chunk (0 = about 32 MB of pixels)
    long piece_bytes = 0;            // option: stream bytes per upload piece of the pipelined decoder (0 = 32 MiB)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> pev;    // event pool (no timing)
    DevBuf ring[3];                  // frame chunks in flight
    DevBuf chain;                    // u64 bit positions between chunks + the chain's error word
    DevBuf bits2;                    // a range's stream moved to its phase (dct3d_encode_u8_place)
    unsigned long long *h_chain = nullptr;   // mapped page-locked mirror of the chain slots, written by the packer itself
    unsigned long long *h_chain_dev = nullptr;   // its device address
    size_t h_chain_n = 0;
    uint8_t *h_byte = nullptr;       // pinned scratch (the first byte of a placed range)
    uint64_t range_bits = 0;         // bit count of the range held in `bits` (dct3d_encode_u8_range)
    bool range_valid = false;
    long chunks_last = 0;            // statistics: pipeline chunks of the last host-buffer call
    DecParams part_P;                // segment arrays of the last count-only pass over a stream part (part_locate)
    bool part_valid = false;
};

namespace {

int fail(dct3d_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CU_CHECK(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, DCT3D_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

Layout make_layout(int W, int H, int C, int nslabs)
{
    Layout L;
    L.W = W; L.H = H; L.C = C;
    L.bx = W / C; L.by = H / C;
    const int cpu = kUnitW / C;
    L.bxu = (L.bx + cpu - 1) / cpu;
    L.nslabs = nslabs;
    L.nunits = (long long)nslabs * L.by * L.bxu;   // fused encoder: one unit = one TMA box = one warp pass
    L.ncubes = (long long)nslabs * L.by * L.bx;
    return L;
}

// The reference's zig-zag order (CubeUtils.java:15-37 / CubeUtils.c:17-42), built here from its
// definition: slices of constant x+y+z ascending; inside a slice y outer, z middle, x inner.
void build_zz(int C, std::vector<int> &lin)
{
    lin.clear();
    for (int s = 0; s <= 3 * (C - 1); s++)
        for (int y = 0; y < C; y++)
            for (int z = 0; z < C; z++) {
                const int x = s - y - z;
                if (x >= 0 && x < C) lin.push_back(x + y * C + z * C * C);
            }
}

bool g_tables_ready[64] = {false};
std::mutex g_tables_mu;

int upload_tables(dct3d_ctx *ctx)
{
    std::lock_guard<std::mutex> lock(g_tables_mu);          // contexts may be created concurrently (one per host thread)
    if (ctx->device < 64 && g_tables_ready[ctx->device]) return DCT3D_OK;
    ZzTables t;
    memset(&t, 0, sizeof t);
    for (int C : {8, 4}) {
        std::vector<int> lin;
        build_zz(C, lin);
        std::vector<int> pos(C * C * C);
        for (int i = 0; i < (int)lin.size(); i++) pos[lin[i]] = i;
        for (int i = 0; i < C * C * C; i++) {
            (C == 8 ? t.lin8[i] : t.lin4[i]) = (uint16_t)lin[i];
            (C == 8 ? t.slin8[i] : t.slin4[i]) = (uint16_t)(C == 8 ? coo_swizzle<8>(lin[i]) : coo_swizzle<4>(lin[i]));
        }
        for (int j = 0; j < C; j++)
            for (int s = 0; s < 2 * C - 1; s++) {
                const int k0min = s > C - 1 ? s - (C - 1) : 0, k0max = std::min(s, C - 1);
                const int p0 = pos[(s - k0min) + j * C + k0min * C * C];
                // the kernel relies on the run being contiguous and ordered by k0
                for (int k0 = k0min; k0 <= k0max; k0++)
                    if (pos[(s - k0) + j * C + k0 * C * C] != p0 + (k0 - k0min))
                        return fail(ctx, DCT3D_E_INVALID, "zig-zag run property violated (C=%d j=%d s=%d)", C, j, s);
                (C == 8 ? t.base8[j][s] : t.base4[j][s]) = (uint16_t)p0;
            }
    }
    CU_CHECK(ctx, cudaMemcpyToSymbol(c_zz, &t, sizeof t));
    CU_CHECK(ctx, cudaMemcpyToSymbol(g_zz, &t, sizeof t));
    if (ctx->device < 64) g_tables_ready[ctx->device] = true;
    return DCT3D_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled()
{
    // function-local static: initialised once, thread-safe (contexts are created from several host threads)
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)p;
        return nullptr;
    }();
    return fn;
}

// Tensor map over the u8 frame stack [F][H][W], presented as {W, F, H} (frames before rows) so that one
// box {32 px, C frames, C rows} lands in shared memory as [y][t][32 px] (SWIZZLE_32B): the image a warp
// unit wants (unit_offset() in dct3d_kernels.cuh).
// The inverse kernel's tile store uses the same map over its output frames.
bool make_tmap(CUtensorMap *tm, const void *frames, int W, int H, int F, int C)
{
    EncodeTiledFn fn = get_encode_tiled();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)F, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)W * H, (cuuint64_t)W};
    cuuint32_t box[3] = {(cuuint32_t)kUnitW, (cuuint32_t)C, (cuuint32_t)C};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(frames), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int bind(dct3d_ctx *ctx)
{
    if (!ctx) return fail(nullptr, DCT3D_E_INVALID, "null context");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    return DCT3D_OK;
}

cudaStream_t pick(dct3d_ctx *ctx, void *s) { return s ? (cudaStream_t)s : ctx->stream; }

int check_frames(dct3d_ctx *ctx, int nframes)
{
    if (nframes < 0) return fail(ctx, DCT3D_E_INVALID, "negative frame count");
    return DCT3D_OK;
}

template <int C, int MODE, bool SKIP = false>
int launch_encode(dct3d_ctx *ctx, const EncParams &P, const CUtensorMap &tm, cudaStream_t st)
{
    auto kern = encode_kernel<C, MODE, SKIP>;
    const int smem = EncSmem<C>::TOTAL;
    int &occ = ctx->occ_cache[SKIP ? 6 : (C == 8 ? 0 : 2) + MODE];
    if (occ == 0) {
        CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    }
    if (occ < 1) return fail(ctx, DCT3D_E_CUDA, "encode kernel does not fit on an SM");
    const long long grid = std::min<long long>((P.L.nunits + kWarps - 1) / kWarps, (long long)ctx->num_sms * occ);
    cudaEvent_t *kev = ctx->kev[0][ctx->kcalls[0] % kKevRing];
    if (MODE == MODE_ZZ) cudaEventRecord(kev[0], st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(tm, P);
    if (MODE == MODE_ZZ) { cudaEventRecord(kev[1], st); ctx->kcalls[0]++; }
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

// the fused (zig-zag mode) encoder of the context's cube size and zero-skip option
int launch_encode_zz(dct3d_ctx *ctx, const EncParams &P, const CUtensorMap &tm, cudaStream_t st)
{
    if (ctx->C == 4) return launch_encode<4, MODE_ZZ>(ctx, P, tm, st);
    return ctx->zero_skip ? launch_encode<8, MODE_ZZ, true>(ctx, P, tm, st) : launch_encode<8, MODE_ZZ>(ctx, P, tm, st);
}

// zero the control block and the look-back status array
// The control block and the look-back status words share one allocation (control block first), so that
// one memset resets both.
constexpr size_t kCtrlBytes = 64;
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block grew");

unsigned long long *status_words(dct3d_ctx *ctx) { return (unsigned long long *)((uint8_t *)ctx->ctrl.p + kCtrlBytes); }

int reset_ctrl(dct3d_ctx *ctx, long long ntiles, cudaStream_t st)
{
    const size_t bytes = kCtrlBytes + (size_t)std::max<long long>(ntiles, 1) * 8;
    CU_CHECK(ctx, ctx->ctrl.reserve(bytes));
    CU_CHECK(ctx, cudaMemsetAsync(ctx->ctrl.p, 0, bytes, st));
    return DCT3D_OK;
}

int fetch_ctrl(dct3d_ctx *ctx, cudaStream_t st)
{
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaStreamSynchronize(st));
    return DCT3D_OK;
}

// zero-fill the stream from the byte after start_bit's byte (the partial byte is kept).  With option
// "reuse_zeroed" a buffer this context packed into before (same pointer and capacity) is only wiped up to
// where that call wrote: everything beyond is still zero from the previous wipe.
int zero_stream(dct3d_ctx *ctx, void *d_stream, size_t cap, uint64_t start_bit, cudaStream_t st)
{
    const size_t first = (size_t)((start_bit + 7) / 8);
    size_t upto = cap;
    if (ctx->reuse_zeroed && d_stream == ctx->clean_ptr && cap == ctx->clean_cap && ctx->clean_dirty)
        upto = std::min(cap, ctx->clean_dirty + 64);
    if (upto > first) CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)d_stream + first, 0, upto - first, st));
    ctx->clean_ptr = d_stream;
    ctx->clean_cap = cap;
    ctx->clean_dirty = 0;            // unknown until the end bit of this call is read back
    return DCT3D_OK;
}

#define H2D(ctx, dst, src, n) CU_CHECK(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream))
#define D2H(ctx, dst, src, n) CU_CHECK(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream))
#define SYNC(ctx) CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream))


template <int C>
static int launch_reconstruct_coo(dct3d_ctx *ctx, const Layout &L, void *d_frames, cudaStream_t st, long long cube_base = 0)
{
    // the lists hold at most one entry per coefficient of the parsed cubes; ctx->coo has room for Geo<C>::CS more
    const unsigned long long coo_limit = ctx->coo.cap >= (size_t)Geo<C>::CS * 8 ? (unsigned long long)(ctx->coo.cap / 4 - Geo<C>::CS) : 0ull;
    // TMA tile store: a group of CPW cubes must be one unit of a cube row (width % 32 == 0) and the frames 16-byte aligned
    const bool tma_out = ctx->tma_store && ctx->W % kUnitW == 0 && !((uintptr_t)d_frames & 15) && L.nslabs > 0 && get_encode_tiled();
    ctx->tma_store_used = tma_out ? 1 : 0;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (tma_out && !make_tmap(&tm, d_frames, ctx->W, ctx->H, L.nslabs * C, C))
        return fail(ctx, DCT3D_E_CUDA, "cuTensorMapEncodeTiled failed (set option tma_store=0 to use row stores)");
    auto kern = tma_out ? reconstruct_coo_kernel<C, TAIL_TMA> : reconstruct_coo_kernel<C, TAIL_ROWS>;
    const int smem = CooSmem<C>::TOTAL;
    int &occ = ctx->occ_cache[tma_out ? 7 : 5];
    if (occ == 0) {
        CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    }
    const long long groups = (L.ncubes + Geo<C>::CPW - 1) / Geo<C>::CPW;
    if (groups > 0x7ff00000ll) return fail(ctx, DCT3D_E_INVALID, "too many cubes for one call");   // the kernel counts groups in 31 bits
    const long long grid = std::min<long long>((groups + kWarps - 1) / kWarps, (long long)ctx->num_sms * std::max(occ, 1));
    cudaEvent_t *kev = ctx->kev[1][ctx->kcalls[1] % kKevRing];
    cudaEventRecord(kev[0], st);
    kern<<<(unsigned)grid, kThreads, smem, st>>>(tm, L, (const uint32_t *)ctx->coo.p, (const unsigned long long *)ctx->coocnt.p, (uint8_t *)d_frames, cube_base, coo_limit);
    cudaEventRecord(kev[1], st);
    ctx->kcalls[1]++;
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

template <int C>
static int launch_reconstruct(dct3d_ctx *ctx, const Layout &L, const void *d_q, void *d_frames, cudaStream_t st)
{
    const int smem = kWarps * Xch<C, float>::WARP_BYTES + C * C * C * 4;   // exchange buffers + dequantiser table
    const long long groups = (L.ncubes + Geo<C>::CPW - 1) / Geo<C>::CPW;
    const long long grid = std::min<long long>((groups + kWarps - 1) / kWarps, (long long)ctx->num_sms * 8);
    reconstruct_kernel<C><<<(unsigned)grid, kThreads, smem, st>>>(L, (const int16_t *)d_q, (uint8_t *)d_frames);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

template <typename T, bool CUBEMAJOR, bool INVERSE>
static int transform_dev(dct3d_ctx *ctx, const void *d_in, void *d_out, int nslabs, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (nslabs < 0) return fail(ctx, DCT3D_E_INVALID, "negative slab count");
    if (nslabs == 0) return DCT3D_OK;
    if (!d_in || !d_out) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(ctx, DCT3D_E_INVALID, "buffers must be 16-byte aligned");
    if (!CUBEMAJOR && (ctx->W * sizeof(T)) % 16) return fail(ctx, DCT3D_E_INVALID, "planar rows must be 16-byte multiples");
    const Layout L = make_layout(ctx->W, ctx->H, ctx->C, nslabs);
    cudaStream_t st = pick(ctx, cuda_stream);
    const int cpw = 32 / ctx->C;
    const long long groups = (L.ncubes + cpw - 1) / cpw;
    const long long grid = std::min<long long>((groups + kWarps - 1) / kWarps, (long long)ctx->num_sms * 8);
    if (ctx->C == 8) {
        const int smem = kWarps * Xch<8, T>::WARP_BYTES;
        transform_kernel<8, T, CUBEMAJOR, INVERSE><<<(unsigned)grid, kThreads, smem, st>>>(L, (const T *)d_in, (T *)d_out);
    } else {
        const int smem = kWarps * Xch<4, T>::WARP_BYTES;
        transform_kernel<4, T, CUBEMAJOR, INVERSE><<<(unsigned)grid, kThreads, smem, st>>>(L, (const T *)d_in, (T *)d_out);
    }
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

template <typename T>
static int transform_host(dct3d_ctx *ctx, const T *in, T *out, size_t count, int arg,
                          int (*dev)(dct3d_ctx *, const void *, void *, int, void *))
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (count == 0) return DCT3D_OK;
    if (!in || !out) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    CU_CHECK(ctx, ctx->fa.reserve(count * sizeof(T)));
    CU_CHECK(ctx, ctx->fb.reserve(count * sizeof(T)));
    H2D(ctx, ctx->fa.p, in, count * sizeof(T));
    if ((rc = dev(ctx, ctx->fa.p, ctx->fb.p, arg, nullptr))) return rc;
    D2H(ctx, out, ctx->fb.p, count * sizeof(T));
    SYNC(ctx);
    return DCT3D_OK;
}


}  // namespace

// ============================================================================================
extern "C" {

int dct3d_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DCT3D_E_CUDA, "no CUDA device"); }
    return n;
}

int dct3d_list_devices(char *buf, size_t cap)
{
    int n = dct3d_device_count();
    if (n < 0) return n;
    std::string s;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) != cudaSuccess) continue;
        char line[256];
        snprintf(line, sizeof line, "%d - %s (sm_%d%d, %d SMs, %.1f GiB)\n", i, p.name, p.major, p.minor,
                 p.multiProcessorCount, (double)p.totalGlobalMem / (1 << 30));
        s += line;
    }
    if (buf && cap) { snprintf(buf, cap, "%s", s.c_str()); }
    return n;
}

int dct3d_create(dct3d_ctx **out, int device, int width, int height, int cube)
{
    if (!out) return fail(nullptr, DCT3D_E_INVALID, "null output pointer");
    *out = nullptr;
    if (cube != 8 && cube != 4) return fail(nullptr, DCT3D_E_INVALID, "cube edge must be 8 or 4, got %d", cube);
    if (width <= 0 || height <= 0 || width % cube || height % cube)
        return fail(nullptr, DCT3D_E_INVALID, "frame %dx%d is not a positive multiple of the cube edge %d", width, height, cube);
    int n = dct3d_device_count();
    if (n <= 0) return fail(nullptr, DCT3D_E_CUDA, "no CUDA device available (libdct3d has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, DCT3D_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
    dct3d_ctx *ctx = new dct3d_ctx();
    ctx->device = device; ctx->W = width; ctx->H = height; ctx->C = cube;
    ctx->use_tma = (width % 16 == 0) ? 1 : 0;
    // defaults of the kernel-variant options can be overridden from the environment (A/B runs of a whole test suite)
    if (const char *e = getenv("DCT3D_ZERO_SKIP")) ctx->zero_skip = atoi(e) ? 1 : 0;
    if (const char *e = getenv("DCT3D_TMA_STORE")) ctx->tma_store = atoi(e) ? 1 : 0;
    if (const char *e = getenv("DCT3D_PACK_SORT")) ctx->pack_sort = std::min(std::max(atoi(e), 0), 2);
    int rc = bind(ctx);
    if (rc == DCT3D_OK) {
        cudaError_t e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
        for (int k = 0; k < 2; k++)
            for (int i = 0; i < kKevRing; i++)
                for (int j = 0; j < 2 && e == cudaSuccess; j++) e = cudaEventCreate(&ctx->kev[k][i][j]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_ctrl, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_ctrl, sizeof(Ctrl));
        if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_u64, 4 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_byte, 16);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
        if (e != cudaSuccess) rc = fail(ctx, DCT3D_E_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    }
    if (rc == DCT3D_OK) rc = upload_tables(ctx);
    if (rc != DCT3D_OK) { std::string keep = ctx->err; dct3d_destroy(ctx); g_last_error = keep; return rc; }
    *out = ctx;
    return DCT3D_OK;
}

void dct3d_destroy(dct3d_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (DevBuf *b : {&ctx->frames, &ctx->bits, &ctx->q, &ctx->ctrl, &ctx->seg, &ctx->seglist, &ctx->fa, &ctx->fb, &ctx->zz, &ctx->cmask, &ctx->coo, &ctx->coocnt,
                      &ctx->ring[0], &ctx->ring[1], &ctx->ring[2], &ctx->chain, &ctx->bits2}) b->release();
    for (cudaEvent_t e : ctx->pev) cudaEventDestroy(e);
    if (ctx->h_chain) cudaFreeHost(ctx->h_chain);
    if (ctx->h_byte) cudaFreeHost(ctx->h_byte);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < kKevRing; i++)
            for (int j = 0; j < 2; j++) if (ctx->kev[k][i][j]) cudaEventDestroy(ctx->kev[k][i][j]);
    if (ctx->ev_ctrl) cudaEventDestroy(ctx->ev_ctrl);
    if (ctx->h_ctrl) cudaFreeHost(ctx->h_ctrl);
    if (ctx->h_u64) cudaFreeHost(ctx->h_u64);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *dct3d_last_error(const dct3d_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

void *dct3d_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void dct3d_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int dct3d_set_option(dct3d_ctx *ctx, const char *key, long value)
{
    if (!ctx || !key) return fail(ctx, DCT3D_E_INVALID, "null argument");
    if (!strcmp(key, "tma")) {
        if (value && ctx->W % 16) return fail(ctx, DCT3D_E_INVALID, "TMA loads need width %% 16 == 0");
        if (value && !get_encode_tiled()) return fail(ctx, DCT3D_E_CUDA, "cuTensorMapEncodeTiled unavailable");
        ctx->use_tma = value ? 1 : 0;
        return DCT3D_OK;
    }
    if (!strcmp(key, "debug")) { ctx->debug = (int)value; return DCT3D_OK; }
    if (!strcmp(key, "zero_skip")) { ctx->zero_skip = value ? 1 : 0; return DCT3D_OK; }
    if (!strcmp(key, "pack_sort")) {
        if (value < 0 || value > 2) return fail(ctx, DCT3D_E_INVALID, "pack_sort must be 0, 1 or 2");
        ctx->pack_sort = (int)value;
        return DCT3D_OK;
    }
    if (!strcmp(key, "tma_store")) {
        if (value && !get_encode_tiled()) return fail(ctx, DCT3D_E_CUDA, "cuTensorMapEncodeTiled unavailable");
        ctx->tma_store = value ? 1 : 0;
        return DCT3D_OK;
    }
    if (!strcmp(key, "precision")) {
        if (value != 32 && value != 64) return fail(ctx, DCT3D_E_INVALID, "precision must be 32 or 64");
        ctx->precision = (int)value;
        return DCT3D_OK;
    }
    if (!strcmp(key, "rounding")) { ctx->rounding = value ? 1 : 0; return DCT3D_OK; }
    if (!strcmp(key, "kernel_times_reset")) { ctx->kcalls[0] = ctx->kcalls[1] = 0; return DCT3D_OK; }
    if (!strcmp(key, "piece_bytes")) {
        if (value < 0) return fail(ctx, DCT3D_E_INVALID, "piece_bytes must not be negative");
        ctx->piece_bytes = value;
        return DCT3D_OK;
    }
    if (!strcmp(key, "chunk_frames")) {
        if (value < 0 || value % ctx->C) return fail(ctx, DCT3D_E_INVALID, "chunk_frames must be a non-negative multiple of the cube edge");
        ctx->chunk_frames = (int)value;
        return DCT3D_OK;
    }
    if (!strcmp(key, "reuse_zeroed")) { ctx->reuse_zeroed = value ? 1 : 0; ctx->clean_ptr = nullptr; return DCT3D_OK; }
    return fail(ctx, DCT3D_E_INVALID, "unknown option '%s'", key);
}

long dct3d_get_stat(const dct3d_ctx *ctx, const char *key)
{
    if (!ctx || !key) return -1;
    if (!strcmp(key, "launches")) return ctx->launches;
    if (!strcmp(key, "tma")) return ctx->use_tma;
    if (!strcmp(key, "zero_skip")) return ctx->zero_skip;
    if (!strcmp(key, "pack_sort")) return ctx->pack_sort;
    if (!strcmp(key, "tma_store")) return ctx->tma_store;
    if (!strcmp(key, "tma_store_used")) return ctx->tma_store_used;
    if (!strcmp(key, "num_sms")) return ctx->num_sms;
    if (!strcmp(key, "chunks")) return ctx->chunks_last;
    // device time of the last encode_kernel / reconstruct_coo_kernel launch, nanoseconds (CUDA events on
    // the launching stream; waits for that launch to finish)
    for (int k = 0; k < 2; k++) {
        const bool last = !strcmp(key, k == 0 ? "ns_encode_kernel" : "ns_reconstruct_kernel");
        const bool avg = !strcmp(key, k == 0 ? "ns_encode_kernel_avg" : "ns_reconstruct_kernel_avg");
        if (!last && !avg) continue;
        // duration of the last launch, or the mean over the launches since "kernel_times_reset" (at most kKevRing)
        const unsigned long n = ctx->kcalls[k];
        if (n == 0) return -1;
        const unsigned long cnt = last ? 1 : std::min<unsigned long>(n, kKevRing);
        double sum = 0;
        for (unsigned long i = n - cnt; i < n; i++) {
            cudaEvent_t const *e = ctx->kev[k][i % kKevRing];
            float ms = 0.f;
            if (cudaEventSynchronize(e[1]) != cudaSuccess) return -1;
            if (cudaEventElapsedTime(&ms, e[0], e[1]) != cudaSuccess) return -1;
            sum += ms;
        }
        return (long)(sum / cnt * 1e6);
    }
    if (!strcmp(key, "kernel_launches_timed")) return (long)std::min<unsigned long>(ctx->kcalls[0], kKevRing);
    return -1;
}

// ---- device-resident entry points -----------------------------------------------------------

static unsigned ew_grid(dct3d_ctx *ctx, unsigned long long n)
{
    return (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)ctx->num_sms * 32);
}

// fp64 mode: one fused kernel per direction (codec_f64_kernel), u8 frames <-> natural-order int16 cubes
extern "C++" template <bool INVERSE>
int codec_f64(dct3d_ctx *ctx, const void *d_in, int nslabs, void *d_out, cudaStream_t st)
{
    const int C = ctx->C;
    const Layout L = make_layout(ctx->W, ctx->H, C, nslabs);
    const int cpw = 32 / C;
    const long long groups = (L.ncubes + cpw - 1) / cpw;
    const long long grid = std::min<long long>((groups + kWarps - 1) / kWarps, (long long)ctx->num_sms * 8);
    const uint8_t *fin = INVERSE ? nullptr : (const uint8_t *)d_in;
    const int16_t *qin = INVERSE ? (const int16_t *)d_in : nullptr;
    int16_t *qout = INVERSE ? nullptr : (int16_t *)d_out;
    uint8_t *fout = INVERSE ? (uint8_t *)d_out : nullptr;
    if (C == 8) codec_f64_kernel<8, INVERSE><<<(unsigned)grid, kThreads, kWarps * Xch<8, double>::WARP_BYTES + 256, st>>>(L, fin, qout, qin, fout, ctx->rounding);
    else codec_f64_kernel<4, INVERSE><<<(unsigned)grid, kThreads, kWarps * Xch<4, double>::WARP_BYTES + 256, st>>>(L, fin, qout, qin, fout, ctx->rounding);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

static int quantize_f64(dct3d_ctx *ctx, const void *d_frames, int nslabs, void *d_q, cudaStream_t st)
{
    return codec_f64<false>(ctx, d_frames, nslabs, d_q, st);
}

static int reconstruct_f64(dct3d_ctx *ctx, const void *d_q, int nslabs, void *d_frames, cudaStream_t st)
{
    return codec_f64<true>(ctx, d_q, nslabs, d_frames, st);
}

// Bit positions chained on the device: the pipelined host-buffer encoder codes a clip as consecutive slab ranges
// without visiting the host in between (the rule of expGolomb_freeBuffer, C/ExpGolomb.c:112-122, kept on the GPU).
struct Chain {
    const unsigned long long *d_start = nullptr;   // the packer reads its start bit here ...
    unsigned long long *d_end = nullptr;           // ... and leaves its end bit here
    unsigned long long *h_end = nullptr;           // ... and here (device address of mapped host memory), for the host
    unsigned int *d_err = nullptr;                 // sticky error word of the whole chain (not reset per call)
};

// Kernel 2 (bit packing) over ctx->zz / ctx->cmask.  P.L.ncubes must be set.
static int run_pack_noreset(dct3d_ctx *ctx, EncParams &P, void *d_stream, size_t cap, uint64_t start_bit, uint64_t *end_bit, cudaStream_t st,
                            const Chain *chain = nullptr)
{
    int rc;
    const long long ptiles = (P.L.ncubes + kPackWorkers - 1) / kPackWorkers;
    Ctrl *dc = (Ctrl *)ctx->ctrl.p;
    P.out_words = (uint32_t *)d_stream;
    P.cap_bits = (unsigned long long)(cap / 4) * 32;
    P.start_bit = start_bit;
    P.tile_status = status_words(ctx);
    P.ticket = &dc->ticket; P.err = &dc->err; P.end_bit = &dc->end_bit;
    if (chain) { P.start_bit_dev = chain->d_start; P.end_bit = chain->d_end; P.end_bit_host = chain->h_end; P.err = chain->d_err; }
    // the grid must be resident as a whole: a tile's look-back spins on its predecessors
    // pack_sort: 0 = cubes in stream order, 1 = sorted deal, 2 = sorted deal with balanced warps (tiles of 2 x kPackThreads
    // cubes: the status words and ptiles above, counted in tiles of kPackWorkers, are more than it needs)
    void (*kern)(const EncParams) = ctx->pack_sort == 2 ? (ctx->C == 8 ? eg_pack_balanced_kernel<8> : eg_pack_balanced_kernel<4>)
                                    : ctx->pack_sort ? (ctx->C == 8 ? eg_pack_sorted_kernel<8> : eg_pack_sorted_kernel<4>)
                                                     : (ctx->C == 8 ? eg_pack_kernel<8> : eg_pack_kernel<4>);
    int &occ = ctx->occ_cache[ctx->pack_sort == 2 ? 10 : ctx->pack_sort ? 9 : 4];
    if (occ == 0) CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPackThreads, 0));
    const long long grid = std::min<long long>(ptiles, (long long)ctx->num_sms * std::max(occ, 1));
    kern<<<(unsigned)grid, kPackThreads, 0, st>>>(P);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    if (end_bit) {
        if ((rc = fetch_ctrl(ctx, st))) return rc;
        if (ctx->h_ctrl->err & 16u) return fail(ctx, DCT3D_E_CUDA, "TMA tile load timed out");
        if (ctx->h_ctrl->err & 8u) return fail(ctx, DCT3D_E_CUDA, "tile look-back timed out");
        if (ctx->h_ctrl->err & 1u) return fail(ctx, DCT3D_E_OVERFLOW, "stream buffer of %zu bytes is too small", cap);
        *end_bit = ctx->h_ctrl->end_bit;
        if (d_stream == ctx->clean_ptr) ctx->clean_dirty = (size_t)(ctx->h_ctrl->end_bit / 8) + 1;
    }
    return DCT3D_OK;
}

static int run_pack(dct3d_ctx *ctx, EncParams &P, void *d_stream, size_t cap, uint64_t start_bit, uint64_t *end_bit, cudaStream_t st)
{
    const long long ptiles = (P.L.ncubes + kPackWorkers - 1) / kPackWorkers;
    int rc = reset_ctrl(ctx, ptiles, st);
    return rc ? rc : run_pack_noreset(ctx, P, d_stream, cap, start_bit, end_bit, st);
}

static int encode_common(dct3d_ctx *ctx, const void *d_frames, int nframes, void *d_stream, size_t cap,
                         uint64_t start_bit, uint64_t *end_bit, void *cuda_stream, void *d_qcubes, const Chain *chain = nullptr)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const int C = ctx->C;
    const int nslabs = nframes / C;
    cudaStream_t st = pick(ctx, cuda_stream);
    const bool emit_q = d_qcubes != nullptr;
    if (!emit_q) {
        if (!d_stream || ((uintptr_t)d_stream & 3)) return fail(ctx, DCT3D_E_INVALID, "stream buffer must be non-null and 4-byte aligned");
        if (cap < start_bit / 8 + 16) return fail(ctx, DCT3D_E_OVERFLOW, "stream capacity %zu too small", cap);
        if (nslabs == 0 && (rc = zero_stream(ctx, d_stream, cap, start_bit, st))) return rc;
    }
    if (nslabs == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!d_frames) return fail(ctx, DCT3D_E_INVALID, "null frame pointer");
    if ((uintptr_t)d_frames % (size_t)C) return fail(ctx, DCT3D_E_INVALID, "frame buffer must be aligned to the cube edge (%d bytes)", C);
    if (emit_q && ((uintptr_t)d_qcubes & 15)) return fail(ctx, DCT3D_E_INVALID, "cube buffer must be 16-byte aligned");
    if (chain && ctx->precision == 64) return fail(ctx, DCT3D_E_INVALID, "chained ranges are an fp32-path feature");
    if (ctx->precision == 64) {
        // fp64 mode: u8 -> double, the f64 transform seam, the reference's quantiser in double, then the
        // same zig-zag gather and bit packer as the stage entry point
        if (emit_q) return quantize_f64(ctx, d_frames, nslabs, d_qcubes, st);
        const size_t ncubes = (size_t)nslabs * (ctx->H / C) * (ctx->W / C);
        CU_CHECK(ctx, ctx->q.reserve(ncubes * C * C * C * sizeof(int16_t)));
        if ((rc = quantize_f64(ctx, d_frames, nslabs, ctx->q.p, st))) return rc;
        return dct3d_eg_encode_i16_dev(ctx, ctx->q.p, ncubes, start_bit, d_stream, cap, end_bit, st);
    }
    EncParams P;
    memset(&P, 0, sizeof P);
    P.L = make_layout(ctx->W, ctx->H, C, nslabs);
    P.frames = (const uint8_t *)d_frames;
    P.qcubes = (int16_t *)d_qcubes;
    P.use_tma = ctx->use_tma && !((uintptr_t)d_frames & 15);
    P.debug = ctx->debug;
    CU_CHECK(ctx, ctx->ctrl.reserve(kCtrlBytes + (size_t)((P.L.ncubes + kPackWorkers - 1) / kPackWorkers) * 8));   // final size: the pointer below stays valid
    P.err = chain ? chain->d_err : &((Ctrl *)ctx->ctrl.p)->err;
    if (!emit_q) {
        CU_CHECK(ctx, ctx->zz.reserve((size_t)P.L.ncubes * C * C * C * sizeof(int16_t)));
        CU_CHECK(ctx, ctx->cmask.reserve((size_t)P.L.ncubes * 4));
        P.zzg = (int16_t *)ctx->zz.p;
        P.cmask = (uint32_t *)ctx->cmask.p;
    } else {
        CU_CHECK(ctx, cudaMemsetAsync(ctx->ctrl.p, 0, sizeof(Ctrl), st));
    }
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (P.use_tma && !make_tmap(&tm, d_frames, ctx->W, ctx->H, nslabs * C, C))
        return fail(ctx, DCT3D_E_CUDA, "cuTensorMapEncodeTiled failed (set option tma=0 to use plain loads)");
    if (emit_q) {
        rc = C == 8 ? launch_encode<8, MODE_NAT>(ctx, P, tm, st) : launch_encode<4, MODE_NAT>(ctx, P, tm, st);
        if (rc) return rc;
        if ((rc = fetch_ctrl(ctx, st))) return rc;
        if (ctx->h_ctrl->err & 16u) return fail(ctx, DCT3D_E_CUDA, "TMA tile load timed out");
        return DCT3D_OK;
    }
    // the control block is zeroed here (before kernel 1, which may flag a TMA time-out in it) and
    // again only partially by run_pack: keep one reset, done by run_pack, and run kernel 1 after it
    const long long ptiles = (P.L.ncubes + kPackWorkers - 1) / kPackWorkers;
    if ((rc = reset_ctrl(ctx, ptiles, st))) return rc;
    if (chain) {                                                   // the caller wiped the whole stream buffer once
        rc = launch_encode_zz(ctx, P, tm, st);
        return rc ? rc : run_pack_noreset(ctx, P, d_stream, cap, 0, nullptr, st, chain);
    }
    // The wipe of the stream buffer only has to precede kernel 2: it is forked onto the side stream
    // (after whatever the caller's stream did to the buffer before) and joined in front of the packer,
    // so it runs beside kernel 1, which leaves DRAM 80% idle.
    CU_CHECK(ctx, cudaEventRecord(ctx->ev_fork, st));
    CU_CHECK(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
    rc = zero_stream(ctx, d_stream, cap, start_bit, ctx->aux);
    CU_CHECK(ctx, cudaEventRecord(ctx->ev_join, ctx->aux));
    if (rc == DCT3D_OK) rc = launch_encode_zz(ctx, P, tm, st);
    CU_CHECK(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));     // join the wipe (also on the error path)
    if (rc) return rc;
    return run_pack_noreset(ctx, P, d_stream, cap, start_bit, end_bit, st);
}

int dct3d_encode_u8_dev(dct3d_ctx *ctx, const void *d_frames, int nframes, void *d_stream, size_t cap,
                        uint64_t start_bit, uint64_t *end_bit, void *cuda_stream)
{
    return encode_common(ctx, d_frames, nframes, d_stream, cap, start_bit, end_bit, cuda_stream, nullptr);
}

int dct3d_quantize_u8_dev(dct3d_ctx *ctx, const void *d_frames, int nframes, void *d_qcubes, void *cuda_stream)
{
    if (!d_qcubes) return fail(ctx, DCT3D_E_INVALID, "null output pointer");
    return encode_common(ctx, d_frames, nframes, nullptr, 0, 0, nullptr, cuda_stream, d_qcubes);
}

int dct3d_eg_encode_i16_dev(dct3d_ctx *ctx, const void *d_qcubes, size_t ncubes, uint64_t start_bit,
                            void *d_stream, size_t cap, uint64_t *end_bit, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    cudaStream_t st = pick(ctx, cuda_stream);
    if (!d_stream || ((uintptr_t)d_stream & 3)) return fail(ctx, DCT3D_E_INVALID, "stream buffer must be non-null and 4-byte aligned");
    if (cap < start_bit / 8 + 16) return fail(ctx, DCT3D_E_OVERFLOW, "stream capacity %zu too small", cap);
    if ((rc = zero_stream(ctx, d_stream, cap, start_bit, st))) return rc;
    if (ncubes == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!d_qcubes) return fail(ctx, DCT3D_E_INVALID, "null cube pointer");
    if ((uintptr_t)d_qcubes & 15) return fail(ctx, DCT3D_E_INVALID, "cube buffer must be 16-byte aligned");
    const int C = ctx->C;
    EncParams P;
    memset(&P, 0, sizeof P);
    P.L = make_layout(ctx->W, ctx->H, C, 0);
    P.L.ncubes = (long long)ncubes;
    CU_CHECK(ctx, ctx->zz.reserve(ncubes * C * C * C * sizeof(int16_t)));
    CU_CHECK(ctx, ctx->cmask.reserve(ncubes * 4));
    P.zzg = (int16_t *)ctx->zz.p;
    P.cmask = (uint32_t *)ctx->cmask.p;
    P.qcubes_in = (const int16_t *)d_qcubes;
    const long long grid = std::min<long long>(((long long)ncubes + kWarps - 1) / kWarps, (long long)ctx->num_sms * 16);
    if (C == 8) zz_gather_kernel<8><<<(unsigned)grid, kThreads, 0, st>>>(P);
    else zz_gather_kernel<4><<<(unsigned)grid, kThreads, 0, st>>>(P);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return run_pack(ctx, P, d_stream, cap, start_bit, end_bit, st);
}

// Index discovery + emit: stream -> CSR lists of the cubes' non-zero coefficients (ctx->coo, ctx->coocnt).
// `tail` (optional) enqueues what consumes the lists (the inverse kernel): it is launched BEFORE the host looks at the
// control block, so the GPU does not idle for a host round trip between index discovery and the inverse transform.  In
// the rare case that the fix-up rounds had not converged, or the stream is damaged, the tail has run on inconsistent
// (but bounds-safe) lists: it is run again after convergence, or the call fails and the output is unspecified.
// `part` (optional): count-only pass over a PART of a stream, bits [start_bit, part->count_end_bit), for the distributed
// index discovery of dct3d_multi_locate: no lists are emitted; the segment arrays stay in the context for part_locate().
struct PartOpts {
    uint64_t count_end_bit = 0;   // codes that start before this bit are counted
    int first_entry = 0;          // DecParams::first_entry
    // emit = true: a PIECE of a stream that is parsed piece by piece while it is still being uploaded (pipe_decode): the lists
    // and row pointers of the whole clip are filled in as the pieces arrive, the counts continue from the bases
    bool emit = false;
    uint64_t code_base = 0, nz_base = 0;
    uint64_t ncodes = 0;          // out: codes that start in the part (emit: + code_base)
    uint64_t nz_total = 0;        // out (emit): non-zero codes so far
    uint32_t over_out = 0;        // out: how far the last of them runs past count_end_bit
    uint32_t entry_used = 0;      // out: entry point of the part's first segment
    bool complete = false;        // out (emit): the clip's last code lies in this piece ...
    uint64_t end_bit = 0;         // ... and ends here
};

static int parse_common(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit, size_t ncubes,
                        uint64_t *end_bit, cudaStream_t st, bool locate_only = false, const std::function<int()> *tail = nullptr,
                        PartOpts *part = nullptr)
{
    int rc;
    const int C = ctx->C, CS = C * C * C;
    if (!d_stream || ((uintptr_t)d_stream & 3)) return fail(ctx, DCT3D_E_INVALID, "stream buffer must be non-null and 4-byte aligned");
    if ((uint64_t)nbytes * 8 <= start_bit) return fail(ctx, DCT3D_E_NEED_MORE, "stream holds no data past the start bit");
    DecParams P;
    memset(&P, 0, sizeof P);
    P.L = make_layout(ctx->W, ctx->H, C, 0);
    P.L.ncubes = (long long)ncubes;
    P.words = (const uint32_t *)d_stream;
    P.nwords = ((unsigned long long)nbytes + 3) / 4;
    P.nbits_total = (unsigned long long)nbytes * 8;
    P.start_bit = start_bit;
    P.seg_bits = kSegWords * 32;
    P.count_end_bit = part ? std::min<unsigned long long>(part->count_end_bit, P.nbits_total) : P.nbits_total;
    P.first_entry = part ? part->first_entry : 0;
    P.code_base = part ? part->code_base : 0;
    P.nz_base = part ? part->nz_base : 0;
    if (part && (P.count_end_bit <= start_bit || (part->first_entry < 0 && start_bit < 128)))
        return fail(ctx, DCT3D_E_INVALID, "bad stream part");
    P.nseg = (P.count_end_bit - start_bit + P.seg_bits - 1) / P.seg_bits;
    // seg arrays: count[nseg] over[nseg+1] used[nseg] work[nseg] (u32) first[nseg+1] nzfirst[nseg+1] (u64)
    const size_t n = (size_t)P.nseg;
    const size_t off_first = ((4 * n + 1) * 4 + 7) & ~(size_t)7;
    CU_CHECK(ctx, ctx->seg.reserve(off_first + 2 * (n + 1) * 8));
    CU_CHECK(ctx, ctx->seglist.reserve(((n + 31) / 32) * 32 * kSegListVec * sizeof(uint4)));   // dense-addressed, only the heads are touched
    const long long stiles = (long long)((P.nseg + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems));
    CU_CHECK(ctx, ctx->ctrl.reserve(kCtrlBytes + (size_t)stiles * 16));
    if (!locate_only && (!part || part->emit)) {
        CU_CHECK(ctx, ctx->coo.reserve(ncubes * CS * sizeof(uint32_t) + 2048));   // worst case: every coefficient non-zero, + the tail of the last segment's list
        CU_CHECK(ctx, ctx->coocnt.reserve((ncubes + 1) * 8));
    }
    P.locate_only = locate_only ? 1 : 0;
    P.seg_count = (unsigned int *)ctx->seg.p;
    P.seg_over = P.seg_count + n;
    P.seg_used = P.seg_over + n + 1;
    P.seg_list = (uint4 *)ctx->seglist.p;
    P.seg_first = (unsigned long long *)((uint8_t *)ctx->seg.p + off_first);
    P.seg_nzfirst = P.seg_first + n + 1;
    Ctrl *dc = (Ctrl *)ctx->ctrl.p;
    P.changed = &dc->changed; P.err = &dc->err; P.end_bit = &dc->end_bit;
    P.coo = (uint32_t *)ctx->coo.p;
    P.coo_start = (unsigned long long *)ctx->coocnt.p;
    CU_CHECK(ctx, cudaMemsetAsync(ctx->ctrl.p, 0, kCtrlBytes + (size_t)stiles * 16, st));   // control block + both status arrays of the prefix
    const unsigned sb = kSegThreads, sg = (unsigned)((P.nseg + sb - 1) / sb);
    seg_scan_kernel<<<sg, sb, 0, st>>>(P);
    ctx->launches++;
    unsigned int round = 0;
    auto fix_round = [&]() -> int {
        seg_fix_kernel<<<(unsigned)((P.nseg + 127) / 128), 128, 0, st>>>(P, ++round);
        ctx->launches++;
        return DCT3D_OK;
    };
    bool redo = false;
    auto prefix_and_parse = [&]() -> int {
        if (redo) {                                                          // the first time everything is still zero
            CU_CHECK(ctx, cudaMemsetAsync(status_words(ctx), 0, (size_t)stiles * 16, st));
            CU_CHECK(ctx, cudaMemsetAsync(&dc->ticket, 0, 4, st));
        }
        const long long grid = std::min<long long>(stiles, (long long)ctx->num_sms * 8);
        seg_prefix_kernel<<<(unsigned)grid, kScanThreads, 0, st>>>(P, status_words(ctx), status_words(ctx) + stiles, &dc->ticket);
        if (part && !part->emit) { ctx->launches++; CU_CHECK(ctx, cudaGetLastError()); return DCT3D_OK; }
        const unsigned pg = (unsigned)std::min<unsigned long long>((P.nseg + kEmitThreads - 1) / kEmitThreads, (unsigned long long)ctx->num_sms * 12);
        if (C == 8) seg_emit_kernel<8><<<pg, kEmitThreads, 0, st>>>(P); else seg_emit_kernel<4><<<pg, kEmitThreads, 0, st>>>(P);
        ctx->launches += 2;
        CU_CHECK(ctx, cudaGetLastError());
        return DCT3D_OK;
    };
    // Fix-up rounds re-scan only the segments whose entry point was guessed wrong.  Two rounds are enqueued
    // without looking (with the lead-in walk the first one already finds next to nothing), the rest of
    // the pipeline follows, and the host checks ONCE at the end whether the second round still moved an
    // overhang; only then (the constructed worst case) it iterates to convergence and redoes prefix + emit.
    if ((rc = fix_round()) || (rc = fix_round()) || (rc = prefix_and_parse())) return rc;
    auto fetch_part = [&]() -> int {
        if (!part) return DCT3D_OK;
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_u64 + 1, P.seg_over + n, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_u64 + 2, P.seg_used, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_u64 + 3, P.seg_nzfirst + n, 8, cudaMemcpyDeviceToHost, st));
        return DCT3D_OK;
    };
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_u64, P.seg_first + n, 8, cudaMemcpyDeviceToHost, st));
    if ((rc = fetch_part())) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_ctrl, ctx->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
    CU_CHECK(ctx, cudaEventRecord(ctx->ev_ctrl, st));
    if (tail && (rc = (*tail)())) return rc;
    CU_CHECK(ctx, cudaEventSynchronize(ctx->ev_ctrl));
    if (ctx->h_ctrl->changed == round) {
        for (unsigned long long it = 0; it <= P.nseg && ctx->h_ctrl->changed == round; it++) {
            if ((rc = fix_round()) || (rc = fetch_ctrl(ctx, st))) return rc;
        }
        redo = true;
        CU_CHECK(ctx, cudaMemsetAsync(&dc->err, 0, 4, st));
        if ((rc = prefix_and_parse())) return rc;
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_u64, P.seg_first + n, 8, cudaMemcpyDeviceToHost, st));
        if ((rc = fetch_part())) return rc;
        if ((rc = fetch_ctrl(ctx, st))) return rc;
        if (tail && (rc = (*tail)())) return rc;                 // again, on the converged lists
    }
    if (part) {
        if (ctx->h_ctrl->err & 2u) return fail(ctx, DCT3D_E_STREAM, "malformed Exp-Golomb code in stream");
        part->ncodes = ctx->h_u64[0];
        part->over_out = (uint32_t)ctx->h_u64[1];
        part->entry_used = (uint32_t)ctx->h_u64[2];
        part->nz_total = ctx->h_u64[3];
        if (part->emit) {
            part->complete = part->ncodes >= (unsigned long long)ncubes * CS;
            part->end_bit = ctx->h_ctrl->end_bit;
            if (part->complete && part->end_bit > P.nbits_total)
                return fail(ctx, DCT3D_E_NEED_MORE, "the last code runs past the end of the buffered stream");
            ctx->part_valid = false;
            return DCT3D_OK;
        }
        ctx->part_P = P;
        ctx->part_valid = true;
        return DCT3D_OK;
    }
    ctx->part_valid = false;
    if (ctx->h_u64[0] < (unsigned long long)ncubes * CS)
        return fail(ctx, DCT3D_E_NEED_MORE, "stream holds %llu codes, %llu needed", ctx->h_u64[0], (unsigned long long)ncubes * CS);
    if (ctx->h_ctrl->err & 2u) return fail(ctx, DCT3D_E_STREAM, "malformed Exp-Golomb code in stream");
    if (ctx->h_ctrl->end_bit > P.nbits_total)        // cannot happen with the scan's end-of-stream rule; kept as a guard
        return fail(ctx, DCT3D_E_NEED_MORE, "the last code runs past the end of the buffered stream");
    if (end_bit) *end_bit = ctx->h_ctrl->end_bit;
    return DCT3D_OK;
}

int dct3d_eg_decode_i16_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit,
                            size_t ncubes, void *d_qcubes, uint64_t *end_bit, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (ncubes == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!d_qcubes) return fail(ctx, DCT3D_E_INVALID, "null cube pointer");
    if ((uintptr_t)d_qcubes & 15) return fail(ctx, DCT3D_E_INVALID, "cube buffer must be 16-byte aligned");
    cudaStream_t st = pick(ctx, cuda_stream);
    if ((rc = parse_common(ctx, d_stream, nbytes, start_bit, ncubes, end_bit, st))) return rc;
    DecParams P;
    memset(&P, 0, sizeof P);
    P.L = make_layout(ctx->W, ctx->H, ctx->C, 0);
    P.L.ncubes = (long long)ncubes;
    P.coo = (uint32_t *)ctx->coo.p;
    P.coo_start = (unsigned long long *)ctx->coocnt.p;
    P.qcubes = (int16_t *)d_qcubes;
    CU_CHECK(ctx, cudaMemsetAsync(d_qcubes, 0, ncubes * (size_t)ctx->C * ctx->C * ctx->C * sizeof(int16_t), st));
    const long long grid = std::min<long long>(((long long)ncubes + kWarps - 1) / kWarps, (long long)ctx->num_sms * 16);
    if (ctx->C == 8) coo_scatter_kernel<8><<<(unsigned)grid, kThreads, 0, st>>>(P);
    else coo_scatter_kernel<4><<<(unsigned)grid, kThreads, 0, st>>>(P);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

int dct3d_eg_locate_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit, size_t ncubes,
                        uint64_t *end_bit, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (ncubes == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    return parse_common(ctx, d_stream, nbytes, start_bit, ncubes, end_bit, pick(ctx, cuda_stream), true);
}

int dct3d_reconstruct_i16_dev(dct3d_ctx *ctx, const void *d_qcubes, int nframes, void *d_frames, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const int nslabs = nframes / ctx->C;
    if (nslabs == 0) return DCT3D_OK;
    if (!d_qcubes || !d_frames) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    if (((uintptr_t)d_qcubes & 15) || (uintptr_t)d_frames % (size_t)ctx->C)
        return fail(ctx, DCT3D_E_INVALID, "cube buffer must be 16-byte aligned and the frame buffer aligned to the cube edge");
    const Layout L = make_layout(ctx->W, ctx->H, ctx->C, nslabs);
    cudaStream_t st = pick(ctx, cuda_stream);
    if (ctx->precision == 64) return reconstruct_f64(ctx, d_qcubes, nslabs, d_frames, st);
    return ctx->C == 8 ? launch_reconstruct<8>(ctx, L, d_qcubes, d_frames, st) : launch_reconstruct<4>(ctx, L, d_qcubes, d_frames, st);
}

int dct3d_decode_u8_dev(dct3d_ctx *ctx, const void *d_stream, size_t nbytes, uint64_t start_bit,
                        int nframes, void *d_frames, uint64_t *end_bit, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const int C = ctx->C, nslabs = nframes / C;
    if (nslabs == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!d_frames) return fail(ctx, DCT3D_E_INVALID, "null frame pointer");
    if ((uintptr_t)d_frames % (size_t)C) return fail(ctx, DCT3D_E_INVALID, "frame buffer must be aligned to the cube edge (%d bytes)", C);
    const Layout L = make_layout(ctx->W, ctx->H, C, nslabs);
    cudaStream_t st = pick(ctx, cuda_stream);
    uint64_t end = 0;
    if (ctx->precision == 64) {
        CU_CHECK(ctx, ctx->q.reserve((size_t)L.ncubes * C * C * C * sizeof(int16_t)));
        if ((rc = dct3d_eg_decode_i16_dev(ctx, d_stream, nbytes, start_bit, (size_t)L.ncubes, ctx->q.p, &end, st))) return rc;
        if (end_bit) *end_bit = end;
        return reconstruct_f64(ctx, ctx->q.p, nslabs, d_frames, st);
    }
    const std::function<int()> tail = [&]() -> int {
        return C == 8 ? launch_reconstruct_coo<8>(ctx, L, d_frames, st) : launch_reconstruct_coo<4>(ctx, L, d_frames, st);
    };
    if ((rc = parse_common(ctx, d_stream, nbytes, start_bit, (size_t)L.ncubes, &end, st, false, &tail))) return rc;
    if (end_bit) *end_bit = end;
    return DCT3D_OK;
}

static int rgb_dev(dct3d_ctx *ctx, bool split, void *d_rgb, void *d_r, void *d_g, void *d_b, size_t nbytes, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (nbytes == 0) return DCT3D_OK;
    if (!d_rgb || !d_r || !d_g || !d_b) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    cudaStream_t st = pick(ctx, cuda_stream);
    const int vec_ok = !(((uintptr_t)d_rgb | (uintptr_t)d_r | (uintptr_t)d_g | (uintptr_t)d_b) & 15);
    const unsigned long long work = vec_ok ? nbytes / 48 + 47 : nbytes;
    const unsigned grid = (unsigned)std::min<unsigned long long>((work + 255) / 256, (unsigned long long)ctx->num_sms * 32);
    if (split) rgb_planes_kernel<true><<<grid, 256, 0, st>>>((uint8_t *)d_rgb, (uint8_t *)d_r, (uint8_t *)d_g, (uint8_t *)d_b, nbytes, vec_ok);
    else rgb_planes_kernel<false><<<grid, 256, 0, st>>>((uint8_t *)d_rgb, (uint8_t *)d_r, (uint8_t *)d_g, (uint8_t *)d_b, nbytes, vec_ok);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

int dct3d_rgb_split_dev(dct3d_ctx *ctx, const void *d_rgb, size_t nbytes, void *d_r, void *d_g, void *d_b, void *cuda_stream)
{
    return rgb_dev(ctx, true, const_cast<void *>(d_rgb), d_r, d_g, d_b, nbytes, cuda_stream);
}

int dct3d_rgb_mix_dev(dct3d_ctx *ctx, const void *d_r, const void *d_g, const void *d_b, size_t npixels, void *d_rgb, void *cuda_stream)
{
    return rgb_dev(ctx, false, d_rgb, const_cast<void *>(d_r), const_cast<void *>(d_g), const_cast<void *>(d_b), npixels * 3, cuda_stream);
}

// host buffers: plane p of an n-byte RGB file holds (n + 2 - p) / 3 bytes
int dct3d_rgb_split(dct3d_ctx *ctx, const uint8_t *rgb, size_t nbytes, uint8_t *r, uint8_t *g, uint8_t *b)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (nbytes == 0) return DCT3D_OK;
    if (!rgb || !r || !g || !b) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    const size_t np = (nbytes + 2) / 3, pstride = (np + 255) & ~(size_t)255;
    CU_CHECK(ctx, ctx->fa.reserve(nbytes));
    CU_CHECK(ctx, ctx->fb.reserve(3 * pstride));
    uint8_t *d = (uint8_t *)ctx->fb.p;
    H2D(ctx, ctx->fa.p, rgb, nbytes);
    if ((rc = rgb_dev(ctx, true, ctx->fa.p, d, d + pstride, d + 2 * pstride, nbytes, nullptr))) return rc;
    uint8_t *out[3] = {r, g, b};
    for (int p = 0; p < 3; p++) D2H(ctx, out[p], d + p * pstride, (nbytes + 2 - p) / 3);
    SYNC(ctx);
    return DCT3D_OK;
}

int dct3d_rgb_mix(dct3d_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, size_t npixels, uint8_t *rgb)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (npixels == 0) return DCT3D_OK;
    if (!rgb || !r || !g || !b) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    const size_t pstride = (npixels + 255) & ~(size_t)255;
    CU_CHECK(ctx, ctx->fa.reserve(3 * npixels));
    CU_CHECK(ctx, ctx->fb.reserve(3 * pstride));
    uint8_t *d = (uint8_t *)ctx->fb.p;
    const uint8_t *in[3] = {r, g, b};
    for (int p = 0; p < 3; p++) H2D(ctx, d + p * pstride, in[p], npixels);
    if ((rc = rgb_dev(ctx, false, ctx->fa.p, d, d + pstride, d + 2 * pstride, 3 * npixels, nullptr))) return rc;
    D2H(ctx, rgb, ctx->fa.p, 3 * npixels);
    SYNC(ctx);
    return DCT3D_OK;
}

int dct3d_forward_f32_dev(dct3d_ctx *c, const void *i, void *o, int n, void *s) { return transform_dev<float, true, false>(c, i, o, n, s); }
int dct3d_inverse_f32_dev(dct3d_ctx *c, const void *i, void *o, int n, void *s) { return transform_dev<float, true, true>(c, i, o, n, s); }
int dct3d_forward_f64_dev(dct3d_ctx *c, const void *i, void *o, int nframes, void *s)
{ return c ? transform_dev<double, false, false>(c, i, o, nframes / c->C, s) : fail(nullptr, DCT3D_E_INVALID, "null context"); }
int dct3d_inverse_f64_dev(dct3d_ctx *c, const void *i, void *o, int nframes, void *s)
{ return c ? transform_dev<double, false, true>(c, i, o, nframes / c->C, s) : fail(nullptr, DCT3D_E_INVALID, "null context"); }

// ---- host-buffer entry points ---------------------------------------------------------------

// ---- pipelined host-buffer paths ---------------------------------------------------------------
// dct3d_encode_u8 / dct3d_decode_u8 move a clip as a pipeline of slab-range chunks (about 32 MB of pixels each): the
// H2D copy of chunk i+1 runs beside the kernels of chunk i and beside the D2H copy of what chunk i-1 produced, on three
// streams.  The chunks of one call still form ONE stream: the bit position is chained on the device (Chain above).

static cudaEvent_t pipe_event(dct3d_ctx *ctx, size_t i)
{
    while (ctx->pev.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ctx->pev.push_back(e);
    }
    return ctx->pev[i];
}

static int chunk_slabs(const dct3d_ctx *ctx, int nslabs)
{
    const size_t slab = (size_t)ctx->W * ctx->H * ctx->C;
    const size_t k = ctx->chunk_frames ? (size_t)ctx->chunk_frames / ctx->C : (((size_t)32 << 20) + slab - 1) / slab;
    return (int)std::min<size_t>(std::max<size_t>(k, 1), (size_t)std::max(nslabs, 1));
}

static void pipe_quiesce(dct3d_ctx *ctx)
{
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_d2h);
}

constexpr int kRing = 3;

// Codes host frames into ctx->bits from bit 0.  With `out` the finished bytes flow back to the host beside the following
// chunks and out[0 .. *nbits/8] is the complete stream; without it the stream stays on the device (dct3d_encode_u8_range).
// start_bit (0..7) / first_byte: the stream continues a partial byte carried over from an earlier call (dct3d_stream_encode).
static int pipe_encode(dct3d_ctx *ctx, const uint8_t *frames, int nframes, size_t dcap, uint8_t *out, size_t cap, uint64_t *nbits,
                       unsigned start_bit = 0, uint8_t first_byte = 0)
{
    int rc;
    const int C = ctx->C, nslabs = nframes / C;
    const size_t slab_bytes = (size_t)ctx->W * ctx->H * C;
    cudaStream_t st = ctx->stream;
    ctx->range_valid = false;
    CU_CHECK(ctx, ctx->bits.reserve(dcap));
    const int K = chunk_slabs(ctx, nslabs), nchunks = nslabs ? (nslabs + K - 1) / K : 0;
    ctx->chunks_last = nchunks;
    // the carried partial byte becomes byte 0 of the device buffer
    CU_CHECK(ctx, cudaMemsetAsync(ctx->bits.p, 0, 4, st));
    if (start_bit) {
        ctx->h_byte[0] = first_byte;
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->bits.p, ctx->h_byte, 1, cudaMemcpyHostToDevice, st));
    }
    if (nchunks <= 1 || ctx->precision == 64) {
        // one shot: a single chunk, or the fp64 mode (whose packer is not chained)
        const size_t n = slab_bytes * nslabs;
        CU_CHECK(ctx, ctx->ring[0].reserve(n + 16));
        if (n) H2D(ctx, ctx->ring[0].p, frames, n);
        uint64_t end = 0;
        if ((rc = dct3d_encode_u8_dev(ctx, ctx->ring[0].p, nframes, ctx->bits.p, dcap, start_bit, &end, nullptr))) return rc;
        if (out) {
            const size_t nb = (size_t)(end / 8) + 1;
            if (nb > cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", nb, cap);
            D2H(ctx, out, ctx->bits.p, nb);
        }
        SYNC(ctx);
        *nbits = end;
        return DCT3D_OK;
    }
    // ---- set-up: everything the chunks share is sized for the largest chunk before the first launch, so that no later
    // reserve() can free a buffer under a running kernel
    const size_t cubes_chunk = (size_t)K * (ctx->H / C) * (ctx->W / C);
    for (int b = 0; b < kRing; b++) CU_CHECK(ctx, ctx->ring[b].reserve((size_t)K * slab_bytes + 16));
    CU_CHECK(ctx, ctx->zz.reserve(cubes_chunk * C * C * C * sizeof(int16_t)));
    CU_CHECK(ctx, ctx->cmask.reserve(cubes_chunk * 4));
    CU_CHECK(ctx, ctx->ctrl.reserve(kCtrlBytes + ((cubes_chunk + kPackWorkers - 1) / kPackWorkers) * 8));
    const size_t nslots = (size_t)nchunks + 2;                   // bit position before every chunk and after the last; error word
    CU_CHECK(ctx, ctx->chain.reserve(nslots * 8));
    if (ctx->h_chain_n < nslots) {
        if (ctx->h_chain) cudaFreeHost(ctx->h_chain);
        ctx->h_chain = nullptr; ctx->h_chain_n = 0;
        CU_CHECK(ctx, cudaHostAlloc((void **)&ctx->h_chain, (nslots + 64) * 8, cudaHostAllocMapped));
        CU_CHECK(ctx, cudaHostGetDevicePointer((void **)&ctx->h_chain_dev, ctx->h_chain, 0));
        ctx->h_chain_n = nslots + 64;
    }
    unsigned long long *d_chain = (unsigned long long *)ctx->chain.p;
    unsigned int *d_err = (unsigned int *)(d_chain + nchunks + 1);
    // The end bit of every chunk reaches the host through mapped memory, written by the packer: a D2H copy of 8 bytes would
    // queue behind whatever else the copy engine is doing (another context's frame copies) and stall this thread.
    auto ev = [&](int i, int what) { return pipe_event(ctx, (size_t)2 * i + what); };   // 0: chunk on the device, 1: chunk coded
    if (!ev(nchunks - 1, 1)) return fail(ctx, DCT3D_E_CUDA, "cudaEventCreate failed");
    CU_CHECK(ctx, cudaMemsetAsync(ctx->chain.p, 0, nslots * 8, st));
    if (start_bit) {
        ctx->h_chain[0] = start_bit;
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->chain.p, ctx->h_chain, 8, cudaMemcpyHostToDevice, st));
    }
    if ((rc = zero_stream(ctx, ctx->bits.p, dcap, start_bit, st))) return rc;

    size_t sent = 0;
    auto drain = [&](int j, bool last) -> int {                  // bytes that chunk j completed go home
        CU_CHECK(ctx, cudaEventSynchronize(ev(j, 1)));
        const size_t full = (size_t)(ctx->h_chain[j + 1] / 8), give = last ? full + 1 : full;
        if (give > cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", give, cap);
        if (give > sent) CU_CHECK(ctx, cudaMemcpyAsync(out + sent, (const uint8_t *)ctx->bits.p + sent, give - sent, cudaMemcpyDeviceToHost, ctx->s_d2h));
        sent = std::max(sent, give);
        return DCT3D_OK;
    };
    auto run = [&]() -> int {
        for (int i = 0; i < nchunks; i++) {
            const int s0 = i * K, ns = std::min(K, nslabs - s0), b = i % kRing;
            if (i >= kRing) CU_CHECK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev(i - kRing, 1), 0));
            CU_CHECK(ctx, cudaMemcpyAsync(ctx->ring[b].p, frames + (size_t)s0 * slab_bytes, (size_t)ns * slab_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
            CU_CHECK(ctx, cudaEventRecord(ev(i, 0), ctx->s_h2d));
            CU_CHECK(ctx, cudaStreamWaitEvent(st, ev(i, 0), 0));
            const Chain ch{d_chain + i, d_chain + i + 1, ctx->h_chain_dev + i + 1, d_err};
            if ((rc = encode_common(ctx, ctx->ring[b].p, ns * C, ctx->bits.p, dcap, 0, nullptr, st, nullptr, &ch))) return rc;
            if (i == nchunks - 1)                                // the chain's error word follows the last chunk home
                CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_chain + nchunks + 1, d_err, 8, cudaMemcpyDeviceToHost, st));
            CU_CHECK(ctx, cudaEventRecord(ev(i, 1), st));
            if (out && i >= 1) {
                CU_CHECK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev(i - 1, 1), 0));
                if ((rc = drain(i - 1, false))) return rc;
            }
        }
        CU_CHECK(ctx, cudaEventSynchronize(ev(nchunks - 1, 1)));
        const unsigned int err = (unsigned int)ctx->h_chain[nchunks + 1];
        if (err & 16u) return fail(ctx, DCT3D_E_CUDA, "TMA tile load timed out");
        if (err & 8u) return fail(ctx, DCT3D_E_CUDA, "tile look-back timed out");
        if (err & 1u) return fail(ctx, DCT3D_E_OVERFLOW, "stream buffer of %zu bytes is too small", dcap);
        if (out) {
            CU_CHECK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev(nchunks - 1, 1), 0));
            if ((rc = drain(nchunks - 1, true))) return rc;
        }
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->s_d2h));
        return DCT3D_OK;
    };
    rc = run();
    if (rc) { pipe_quiesce(ctx); ctx->clean_ptr = nullptr; return rc; }
    *nbits = ctx->h_chain[nchunks];
    if (ctx->bits.p == ctx->clean_ptr) ctx->clean_dirty = (size_t)(*nbits / 8) + 1;
    return DCT3D_OK;
}

// Decodes `nframes` frames from a host stream whose bit `start_bit` is the first bit of the first cube.  The stream goes up
// in pieces (option "piece_bytes", default 32 MiB) and is PARSED PIECE BY PIECE while the later pieces are still on their way:
// index discovery is local to a piece once the counts and the overhang of the pieces before it are known, so the lists and
// row pointers of the clip fill in as the pieces arrive, and every slab whose cubes are complete is reconstructed and
// copied home at once.  (First version: upload all, parse all, then reconstruct chunk by chunk: the frame copies, which
// bound the call, started 1.7 ms late on a 76 MB stream.)
static int pipe_decode(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit, int nframes, uint8_t *frames, uint64_t *end_bit)
{
    int rc;
    const int C = ctx->C, CS = C * C * C, nslabs = nframes / C;
    const size_t slab_bytes = (size_t)ctx->W * ctx->H * C;
    cudaStream_t st = ctx->stream;
    ctx->range_valid = false;
    ctx->clean_ptr = nullptr;                                    // ctx->bits is about to hold foreign bytes
    const size_t padded = ((nbytes + 3) & ~(size_t)3) + 8;
    CU_CHECK(ctx, ctx->bits.reserve(padded));
    const int K = chunk_slabs(ctx, nslabs), nchunks = (nslabs + K - 1) / K;
    ctx->chunks_last = nchunks;
    // measured on 1080p x 256 (76 MB of stream): whole 11.3-11.4 ms, 2 MiB pieces 12.7, 8 MiB 11.2, 32 MiB 10.9 (every piece costs a
    // parse with a host round trip, and the uploads share the link with the frame copies going the other way)
    size_t piece = ctx->piece_bytes ? (size_t)ctx->piece_bytes : ((size_t)32 << 20);
    piece = std::max<size_t>((piece + 4095) & ~(size_t)4095, 4096);
    const size_t npieces = (nbytes + piece - 1) / piece;
    if (nchunks <= 1 || ctx->precision == 64 || npieces <= 1) {
        CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)ctx->bits.p + (nbytes & ~(size_t)3), 0, padded - (nbytes & ~(size_t)3), st));
        H2D(ctx, ctx->bits.p, stream, nbytes);
        if (nchunks <= 1 || ctx->precision == 64) {
            const size_t n = slab_bytes * nslabs;
            CU_CHECK(ctx, ctx->ring[0].reserve(n + 16));
            if ((rc = dct3d_decode_u8_dev(ctx, ctx->bits.p, nbytes, start_bit, nframes, ctx->ring[0].p, end_bit, nullptr))) return rc;
            D2H(ctx, frames, ctx->ring[0].p, n);
            SYNC(ctx);
            return DCT3D_OK;
        }
    }
    for (int b = 0; b < kRing; b++) CU_CHECK(ctx, ctx->ring[b].reserve((size_t)K * slab_bytes + 16));
    // events: one per piece (uploaded), then 2 per reconstruct launch (0: reconstructed, 1: on the host); pieces rarely end on
    // chunk boundaries, so there can be more launches than ceil(nslabs / K)
    auto ev_up = [&](size_t p) { return pipe_event(ctx, p); };
    auto ev = [&](int i, int what) { return pipe_event(ctx, npieces + 1 + (size_t)2 * i + what); };
    if (!ev(nchunks + (int)npieces + 1, 1)) return fail(ctx, DCT3D_E_CUDA, "cudaEventCreate failed");
    const Layout L = make_layout(ctx->W, ctx->H, C, nslabs);
    const size_t cubes_per_slab = (size_t)L.by * L.bx;
    int slabs_done = 0, chunk = 0;
    auto reconstruct_upto = [&](int ready_slabs) -> int {        // slabs [slabs_done, ready_slabs): inverse transform + copy home
        while (slabs_done < ready_slabs) {
            const int ns = std::min(K, ready_slabs - slabs_done), b = chunk % kRing;
            if (chunk >= kRing) CU_CHECK(ctx, cudaStreamWaitEvent(st, ev(chunk - kRing, 1), 0));
            const Layout Lc = make_layout(ctx->W, ctx->H, C, ns);
            const long long base = (long long)slabs_done * L.by * L.bx;
            if ((rc = C == 8 ? launch_reconstruct_coo<8>(ctx, Lc, ctx->ring[b].p, st, base) : launch_reconstruct_coo<4>(ctx, Lc, ctx->ring[b].p, st, base))) return rc;
            CU_CHECK(ctx, cudaEventRecord(ev(chunk, 0), st));
            CU_CHECK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev(chunk, 0), 0));
            CU_CHECK(ctx, cudaMemcpyAsync(frames + (size_t)slabs_done * slab_bytes, ctx->ring[b].p, (size_t)ns * slab_bytes, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CU_CHECK(ctx, cudaEventRecord(ev(chunk, 1), ctx->s_d2h));
            slabs_done += ns;
            chunk++;
        }
        return DCT3D_OK;
    };
    auto run = [&]() -> int {
        if (npieces <= 1) {                                      // a small stream: parse it whole
            if ((rc = parse_common(ctx, ctx->bits.p, nbytes, start_bit, (size_t)L.ncubes, end_bit, st))) return rc;
            if ((rc = reconstruct_upto(nslabs))) return rc;
            CU_CHECK(ctx, cudaStreamSynchronize(ctx->s_d2h));
            return DCT3D_OK;
        }
        // all uploads are queued at once on the copy stream (zero padding first: it overlaps the last word of the stream)
        CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)ctx->bits.p + (nbytes & ~(size_t)3), 0, padded - (nbytes & ~(size_t)3), ctx->s_h2d));
        for (size_t p = 0; p < npieces; p++) {
            const size_t off = p * piece, len = std::min(piece, nbytes - off);
            CU_CHECK(ctx, cudaMemcpyAsync((uint8_t *)ctx->bits.p + off, stream + off, len, cudaMemcpyHostToDevice, ctx->s_h2d));
            CU_CHECK(ctx, cudaEventRecord(ev_up(p), ctx->s_h2d));
        }
        PartOpts po;
        po.emit = true;
        uint64_t bit0 = start_bit;
        bool complete = false;
        for (size_t p = 0; p < npieces && !complete; p++) {
            const bool final = p == npieces - 1;
            const size_t uploaded = std::min(nbytes, (p + 1) * piece);
            // codes that start 16 bytes or more before the end of what is on the device are there in full, look-ahead included
            const uint64_t count_end = final ? (uint64_t)nbytes * 8 : (uint64_t)(uploaded - 16) * 8;
            if (count_end <= bit0) continue;
            CU_CHECK(ctx, cudaStreamWaitEvent(st, ev_up(p), 0));
            po.count_end_bit = count_end;
            if ((rc = parse_common(ctx, ctx->bits.p, final ? nbytes : uploaded, bit0, (size_t)L.ncubes, nullptr, st, false, nullptr, &po))) return rc;
            po.code_base = po.ncodes;                            // the next piece continues where this one stopped
            po.nz_base = po.nz_total;
            po.first_entry = (int)po.over_out + 1;
            bit0 = count_end;
            complete = po.complete;
            // a cube can be reconstructed once its successor's first code has been seen (that is when its row pointer is final)
            const size_t ready_cubes = complete ? (size_t)L.ncubes : (po.ncodes ? (size_t)((po.ncodes - 1) / CS) : 0);
            if ((rc = reconstruct_upto((int)(ready_cubes / cubes_per_slab)))) return rc;
        }
        if (!complete)
            return fail(ctx, DCT3D_E_NEED_MORE, "stream holds %llu codes, %llu needed", (unsigned long long)po.ncodes, (unsigned long long)L.ncubes * CS);
        if (end_bit) *end_bit = po.end_bit;
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->s_d2h));
        return DCT3D_OK;
    };
    rc = run();
    if (rc) pipe_quiesce(ctx);
    return rc;
}

static size_t frame_bytes(const dct3d_ctx *ctx, int nframes) { return (size_t)ctx->W * ctx->H * (size_t)(nframes - nframes % ctx->C); }

int dct3d_encode_u8(dct3d_ctx *ctx, const uint8_t *frames, int nframes, uint8_t *stream, size_t cap,
                    uint64_t *nbits, size_t *nbytes)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    if (!stream || cap == 0) return fail(ctx, DCT3D_E_INVALID, "null stream buffer");
    if (frame_bytes(ctx, nframes) && !frames) return fail(ctx, DCT3D_E_INVALID, "null frame pointer");
    // device stream capacity: the caller's cap, rounded up to words, plus slack
    const size_t dcap = ((cap + 3) & ~(size_t)3) + 64;
    uint64_t end = 0;
    if ((rc = pipe_encode(ctx, frames, nframes, dcap, stream, cap, &end))) return rc;
    if (nbits) *nbits = end;
    if (nbytes) *nbytes = (size_t)(end / 8) + 1;
    return DCT3D_OK;
}

int dct3d_decode_u8(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, int nframes, uint8_t *frames)
{
    return dct3d_decode_u8_range(ctx, stream, nbytes, 0, 0, nframes, frames, nullptr);
}

int dct3d_decode_u8_range(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit, uint64_t end_bit_hint,
                          int nframes, uint8_t *frames, uint64_t *end_bit)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    if (frame_bytes(ctx, nframes) == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!stream || !frames) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    const size_t byte0 = (size_t)(start_bit / 8);
    size_t upto = nbytes;
    if (end_bit_hint) {
        if (end_bit_hint < start_bit) return fail(ctx, DCT3D_E_INVALID, "end bit hint before the start bit");
        upto = std::min<size_t>(nbytes, (size_t)(end_bit_hint / 8) + 1);
    }
    if (byte0 >= upto) return fail(ctx, DCT3D_E_STREAM, "Exp-Golomb stream truncated: no data at the start bit");
    uint64_t end = 0;
    rc = pipe_decode(ctx, stream + byte0, upto - byte0, start_bit % 8, nframes, frames, &end);
    if (rc == DCT3D_E_NEED_MORE) return fail(ctx, DCT3D_E_STREAM, "Exp-Golomb stream truncated: %s", ctx->err.c_str());
    if (rc) return rc;
    if (end_bit) *end_bit = (uint64_t)byte0 * 8 + end;
    return DCT3D_OK;
}

// ---- sharded encode: a slab range per GPU, placed into the clip's one stream -------------------------

int dct3d_encode_u8_range(dct3d_ctx *ctx, const uint8_t *frames, int nframes, uint64_t *nbits)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const size_t n = frame_bytes(ctx, nframes);
    if (n && !frames) return fail(ctx, DCT3D_E_INVALID, "null frame pointer");
    // worst case of the codec's own output on u8 input is about 2.6 bytes per sample; start at 1/2 and grow on overflow
    size_t dcap = std::max<size_t>(ctx->bits.cap > 64 ? ctx->bits.cap - 64 : 0, n / 2 + 4096) & ~(size_t)3;
    uint64_t end = 0;
    for (;;) {
        rc = pipe_encode(ctx, frames, nframes, dcap, nullptr, 0, &end);
        if (rc != DCT3D_E_OVERFLOW || dcap >= 4 * n + 4096) break;
        dcap = std::min(dcap * 4, 4 * n + 4096) & ~(size_t)3;
    }
    if (rc) return rc;
    ctx->range_bits = end;
    ctx->range_valid = true;
    if (nbits) *nbits = end;
    return DCT3D_OK;
}

int dct3d_stream_shift_dev(dct3d_ctx *ctx, const void *d_src, uint64_t nbits, unsigned phase, void *d_dst, size_t cap, void *cuda_stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!d_src || !d_dst) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    if (((uintptr_t)d_src | (uintptr_t)d_dst) & 3) return fail(ctx, DCT3D_E_INVALID, "stream buffers must be 4-byte aligned");
    if (phase > 7) return fail(ctx, DCT3D_E_INVALID, "phase must be 0..7");
    const unsigned long long src_words = (nbits + 31) / 32, dst_words = (nbits + phase + 31) / 32 + 1;   // + one zero word of slack
    if ((size_t)dst_words * 4 > cap) return fail(ctx, DCT3D_E_OVERFLOW, "shifted stream needs %llu bytes, buffer has %zu", dst_words * 4, cap);
    stream_shift_kernel<<<ew_grid(ctx, dst_words), 256, 0, pick(ctx, cuda_stream)>>>((const uint32_t *)d_src, src_words, (uint32_t *)d_dst, dst_words, (int)phase);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return DCT3D_OK;
}

int dct3d_encode_u8_place(dct3d_ctx *ctx, uint64_t start_bit, int last, uint8_t *stream, size_t cap, uint8_t *first_byte)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ctx->range_valid) return fail(ctx, DCT3D_E_INVALID, "no coded range to place (call dct3d_encode_u8_range first)");
    if (!stream) return fail(ctx, DCT3D_E_INVALID, "null stream buffer");
    if (first_byte) *first_byte = 0;
    const unsigned phase = (unsigned)(start_bit % 8);
    const size_t byte0 = (size_t)(start_bit / 8);
    const uint64_t n = ctx->range_bits, endp = phase + n;        // bits of the placed range, counted from byte0
    if (n == 0) {                                                // an empty range owns no byte, except the stream's closing one
        if (last && phase == 0) { if (byte0 >= cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream buffer too small"); stream[byte0] = 0; }
        return DCT3D_OK;
    }
    const uint8_t *src = (const uint8_t *)ctx->bits.p;
    if (phase) {
        const size_t need = (size_t)((endp + 31) / 32 + 1) * 4;
        CU_CHECK(ctx, ctx->bits2.reserve(need));
        if ((rc = dct3d_stream_shift_dev(ctx, ctx->bits.p, n, phase, ctx->bits2.p, ctx->bits2.cap, nullptr))) return rc;
        src = (const uint8_t *)ctx->bits2.p;
    }
    // bytes first..lastb of the placed range go straight to their place; the first one is shared with the predecessor
    // when phase != 0 and is handed to the caller instead; the byte after the last bit is written only if the range has
    // bits in it or closes the stream (otherwise it belongs to the successor, which starts there at phase 0)
    const size_t first = phase ? 1 : 0;
    const size_t lastb = (endp % 8 != 0 || last) ? (size_t)(endp / 8) : (size_t)(endp / 8) - 1;
    if (byte0 + lastb + 1 > cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", byte0 + lastb + 1, cap);
    if (lastb >= first) D2H(ctx, stream + byte0 + first, src + first, lastb - first + 1);
    if (phase) D2H(ctx, ctx->h_byte, src, 1);
    SYNC(ctx);
    if (phase) {
        if (!first_byte) return fail(ctx, DCT3D_E_INVALID, "the range starts inside a byte: first_byte must not be null");
        *first_byte = ctx->h_byte[0];
    }
    return DCT3D_OK;
}

int dct3d_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return fail(nullptr, DCT3D_E_INVALID, "null buffer");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DCT3D_E_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    return DCT3D_OK;
}

int dct3d_host_unregister(void *p)
{
    if (p && cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DCT3D_E_CUDA, "cudaHostUnregister failed"); }
    return DCT3D_OK;
}

// ---- several GPUs in one process: contiguous slab ranges, one host thread per GPU ----------------------------------
// Every slab is an independent key-frame group (reference README.md:10, slab loop C/encoder.c:203-278); the only coupling
// is the stream's bit position.  GPU g codes slabs [g n/G, (g+1) n/G) from bit 0 of its own buffer, the G bit counts are
// prefix-summed on the host, and every GPU moves its bits to its phase and copies them to their place in the caller's
// one stream; the host ORs the G-1 shared boundary bytes.  No collective.

struct dct3d_multi {
    std::vector<dct3d_ctx *> ctx;
    int W = 0, H = 0, C = 8;
    std::string err;
    uint8_t carry_byte = 0;          // streaming state (dct3d_multi_stream_*), as in dct3d_ctx
    int carry_bits = 0;
    std::vector<double> weight;      // share of the slabs every GPU gets (empty = equal shares): dct3d_multi_set_weights
};

namespace {

int mfail(dct3d_multi *m, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (m) m->err = buf;
    return code;
}

class HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0;
    unsigned long gen = 0;
public:
    explicit HostBarrier(int count) : n(count) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

void slab_range(int nslabs, int g, int G, int &lo, int &hi)
{
    lo = (int)((long long)g * nslabs / G);
    hi = (int)((long long)(g + 1) * nslabs / G);
}

// the same with the context's weights: GPU g gets slabs [round(n W_g), round(n W_{g+1})), W = normalised cumulative weights
void slab_range_w(const dct3d_multi *m, int nslabs, int g, int &lo, int &hi)
{
    const int G = (int)m->ctx.size();
    if (m->weight.empty()) { slab_range(nslabs, g, G, lo, hi); return; }
    double total = 0, before = 0;
    for (int j = 0; j < G; j++) { total += m->weight[j]; if (j < g) before += m->weight[j]; }
    lo = g == 0 ? 0 : (int)(nslabs * (before / total) + 0.5);
    hi = g == G - 1 ? nslabs : (int)(nslabs * ((before + m->weight[g]) / total) + 0.5);
}

// first failure of the per-GPU calls, with the failing context's message
int first_failure(dct3d_multi *m, const std::vector<int> &rcs)
{
    for (size_t g = 0; g < rcs.size(); g++)
        if (rcs[g]) return mfail(m, rcs[g], "GPU %d: %s", m->ctx[g]->device, m->ctx[g]->err.c_str());
    return DCT3D_OK;
}

}  // namespace

int dct3d_multi_create(dct3d_multi **out, const int *devices, int ndevices, int width, int height, int cube)
{
    if (!out) return mfail(nullptr, DCT3D_E_INVALID, "null output pointer");
    *out = nullptr;
    if (ndevices <= 0 || ndevices > 64) return mfail(nullptr, DCT3D_E_INVALID, "device count %d out of range", ndevices);
    dct3d_multi *m = new dct3d_multi();
    m->W = width; m->H = height; m->C = cube;
    for (int i = 0; i < ndevices; i++) {
        dct3d_ctx *c = nullptr;
        const int rc = dct3d_create(&c, devices ? devices[i] : i, width, height, cube);
        if (rc) { const std::string keep = g_last_error; dct3d_multi_destroy(m); g_last_error = keep; return rc; }
        m->ctx.push_back(c);
    }
    *out = m;
    return DCT3D_OK;
}

void dct3d_multi_destroy(dct3d_multi *m)
{
    if (!m) return;
    for (dct3d_ctx *c : m->ctx) dct3d_destroy(c);
    delete m;
}

const char *dct3d_multi_last_error(const dct3d_multi *m) { return m ? m->err.c_str() : g_last_error.c_str(); }

int dct3d_multi_set_option(dct3d_multi *m, const char *key, long value)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    for (dct3d_ctx *c : m->ctx) {
        const int rc = dct3d_set_option(c, key, value);
        if (rc) return mfail(m, rc, "%s", c->err.c_str());
    }
    return DCT3D_OK;
}

int dct3d_multi_set_weights(dct3d_multi *m, const double *weights)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    m->weight.clear();
    if (!weights) return DCT3D_OK;                               // back to equal shares
    double total = 0;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        if (!(weights[g] >= 0)) return mfail(m, DCT3D_E_INVALID, "weights must not be negative");
        total += weights[g];
    }
    if (!(total > 0)) return mfail(m, DCT3D_E_INVALID, "weights must not all be zero");
    m->weight.assign(weights, weights + m->ctx.size());
    return DCT3D_OK;
}

int dct3d_multi_probe_links(dct3d_multi *m, double *h2d_gbs, double *d2h_gbs, double *weights)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    const int G = (int)m->ctx.size();
    const size_t n = (size_t)32 << 20;
    const int reps = 6;
    std::vector<int> rcs(G, 0);
    std::vector<double> up(G, 0), down(G, 0);
    HostBarrier bar(G);
    auto work = [&](int g) {
        dct3d_ctx *ctx = m->ctx[g];
        void *h = nullptr;
        if ((rcs[g] = bind(ctx))) { bar.wait(); bar.wait(); bar.wait(); return; }
        if (ctx->ring[0].reserve(n) != cudaSuccess || !(h = dct3d_host_alloc(n))) rcs[g] = fail(ctx, DCT3D_E_CUDA, "link probe: allocation failed");
        for (int dir = 0; dir < 2; dir++) {
            bar.wait();                                          // every GPU copies the same way at the same time
            if (rcs[g]) continue;
            const auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; r++)
                cudaMemcpyAsync(dir ? h : ctx->ring[0].p, dir ? ctx->ring[0].p : h, n, dir ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rcs[g] = fail(ctx, DCT3D_E_CUDA, "link probe: copy failed"); continue; }
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            (dir ? down : up)[g] = (double)n * reps / s / 1e9;
        }
        bar.wait();
        if (h) dct3d_host_free(h);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    const int rc = first_failure(m, rcs);
    if (rc) return rc;
    for (int g = 0; g < G; g++) {
        if (h2d_gbs) h2d_gbs[g] = up[g];
        if (d2h_gbs) d2h_gbs[g] = down[g];
        // a frame crosses the link once each way in an encode + decode: time per frame ~ 1/up + 1/down
        if (weights) weights[g] = 1.0 / (1.0 / up[g] + 1.0 / down[g]);
    }
    return DCT3D_OK;
}

dct3d_ctx *dct3d_multi_context(dct3d_multi *m, int index)
{
    return (m && index >= 0 && index < (int)m->ctx.size()) ? m->ctx[index] : nullptr;
}

int dct3d_multi_encode_u8(dct3d_multi *m, const uint8_t *frames, int nframes, uint8_t *stream, size_t cap,
                          uint64_t *nbits, size_t *nbytes, uint64_t *range_start_bits)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    if (nframes < 0) return mfail(m, DCT3D_E_INVALID, "negative frame count");
    if (!stream || cap == 0) return mfail(m, DCT3D_E_INVALID, "null stream buffer");
    const int G = (int)m->ctx.size(), C = m->C, nslabs = nframes / C;
    const size_t slab_bytes = (size_t)m->W * m->H * C;
    if (nslabs && !frames) return mfail(m, DCT3D_E_INVALID, "null frame pointer");
    if (G == 1) {                                                // one GPU: its finished bytes stream home beside the coding
        uint64_t nb = 0;
        const int rc = dct3d_encode_u8(m->ctx[0], frames, nframes, stream, cap, &nb, nbytes);
        if (rc) return mfail(m, rc, "%s", m->ctx[0]->err.c_str());
        if (nbits) *nbits = nb;
        if (range_start_bits) { range_start_bits[0] = 0; range_start_bits[1] = nb; }
        return DCT3D_OK;
    }
    std::vector<uint64_t> bits(G, 0), start(G + 1, 0);
    std::vector<int> rcs(G, 0);
    std::vector<uint8_t> fb(G, 0);
    HostBarrier bar(G);
    auto work = [&](int g) {
        int lo, hi;
        slab_range_w(m, nslabs, g, lo, hi);
        rcs[g] = dct3d_encode_u8_range(m->ctx[g], frames + (size_t)lo * slab_bytes, (hi - lo) * C, &bits[g]);
        bar.wait();                                              // every GPU's bit count is known
        for (int j = 0; j < G; j++) if (rcs[j]) return;
        uint64_t b = 0;
        for (int j = 0; j < g; j++) b += bits[j];
        rcs[g] = dct3d_encode_u8_place(m->ctx[g], b, g == G - 1, stream, cap, &fb[g]);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    int rc = first_failure(m, rcs);
    if (rc) return rc;
    for (int g = 0; g < G; g++) start[g + 1] = start[g] + bits[g];
    for (int g = 1; g < G; g++)                                  // the byte a range shares with its predecessor
        if (start[g] % 8) stream[start[g] / 8] |= fb[g];
    if (nbits) *nbits = start[G];
    if (nbytes) *nbytes = (size_t)(start[G] / 8) + 1;
    if (range_start_bits) for (int g = 0; g <= G; g++) range_start_bits[g] = start[g];
    return DCT3D_OK;
}

// ---- distributed index discovery -----------------------------------------------------------------------------------
// The format stores no index (J/ExpGolombReader.java:19-63 is the only way the reference knows where a code starts), so a
// GPU that is to decode slab range g of a stream it did not code has to find the range's first bit.  Instead of every GPU
// scanning the stream in front of its range (G/2 times the stream in total), the BYTES of the stream are cut into G equal
// parts, GPU g counts the codes that start in part g (scan from a guessed entry point, fix-up, prefix: the decoder's own
// index discovery without the emit step), the G counts and overhangs meet on the host, where every part's entry guess is
// checked against its predecessor's overhang (a wrong guess -- rare -- is recounted from the verified entry), and the
// first code of every slab range is then located inside the one part that holds it.  One pass over the stream in total,
// G + G scalars through the host, no collective.

namespace {

// Count the codes of stream bits [part_lo, part_hi) on ctx's GPU.  entry < 0: guess the entry point (part in mid-stream).
int part_count(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t part_lo, uint64_t part_hi, int entry,
               PartOpts &po, uint64_t &window_bit0)
{
    int rc = bind(ctx);
    if (rc) return rc;
    // window: 16 bytes in front of the part (the lead-in walk of the entry guess) and 16 behind (codes that overhang), word aligned
    const size_t lead = entry == 0 ? 0 : 16;
    size_t wb0 = (size_t)(part_lo / 8);
    wb0 = wb0 >= lead ? wb0 - lead : 0;
    wb0 &= ~(size_t)3;
    const size_t wb1 = std::min<size_t>(nbytes, (size_t)((part_hi + 7) / 8) + 16);
    const size_t wn = wb1 - wb0, padded = ((wn + 3) & ~(size_t)3) + 8;
    ctx->range_valid = false;
    ctx->clean_ptr = nullptr;
    CU_CHECK(ctx, ctx->bits.reserve(padded));
    CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)ctx->bits.p + (wn & ~(size_t)3), 0, padded - (wn & ~(size_t)3), ctx->stream));
    H2D(ctx, ctx->bits.p, stream + wb0, wn);
    window_bit0 = (uint64_t)wb0 * 8;
    po.count_end_bit = part_hi - window_bit0;
    po.first_entry = entry;
    return parse_common(ctx, ctx->bits.p, wn, part_lo - window_bit0, 0, nullptr, ctx->stream, false, nullptr, &po);
}

// Bit (relative to the window of the last part_count) at which code number `target` of the part starts.
int part_locate(dct3d_ctx *ctx, uint64_t target, uint64_t *bit)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ctx->part_valid) return fail(ctx, DCT3D_E_INVALID, "no counted stream part");
    Ctrl *dc = (Ctrl *)ctx->ctrl.p;
    seg_locate_kernel<<<1, 32, 0, ctx->stream>>>(ctx->part_P, target, &dc->end_bit);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    if ((rc = fetch_ctrl(ctx, ctx->stream))) return rc;
    if (ctx->h_ctrl->err & 2u) return fail(ctx, DCT3D_E_STREAM, "malformed Exp-Golomb code in stream");
    *bit = ctx->h_ctrl->end_bit;
    return DCT3D_OK;
}

// Start bit of every slab range of the `nframes` frames whose first code is at bit `start_bit` of `stream`.
// sb[0] = start_bit, sb[G] = 0 (unknown: the last range ends where its decoder says).  DCT3D_E_NEED_MORE when the
// buffered bytes do not hold the first code of every range yet.
int multi_discover(dct3d_multi *m, const uint8_t *stream, size_t nbytes, uint64_t start_bit, int nframes, uint64_t *sb)
{
    const int G = (int)m->ctx.size(), C = m->C, nslabs = nframes / C;
    const uint64_t codes_per_slab = (uint64_t)(m->W / C) * (m->H / C) * C * C * C;
    for (int g = 0; g <= G; g++) sb[g] = 0;
    sb[0] = start_bit;
    if (G == 1 || nslabs == 0) return DCT3D_OK;
    const uint64_t total_bits = (uint64_t)nbytes * 8;
    if (total_bits <= start_bit) return mfail(m, DCT3D_E_NEED_MORE, "stream holds no data past the start bit");
    std::vector<int> lo(G + 1, nslabs);
    for (int g = 0; g < G; g++) { int hi; slab_range_w(m, nslabs, g, lo[g], hi); }
    std::vector<int> rcs(G, 0);
    const uint64_t span = total_bits - start_bit;
    if (span < (uint64_t)G * (1u << 19)) {
        // a small stream: every GPU g > 0 simply scans from the start bit to its own range (dct3d_eg_locate)
        const size_t byte0 = (size_t)(start_bit / 8);
        auto work = [&](int g) {
            if (lo[g] == 0) { sb[g] = start_bit; return; }
            uint64_t e = 0;
            rcs[g] = dct3d_eg_locate(m->ctx[g], stream + byte0, nbytes - byte0, start_bit % 8, (size_t)(lo[g] * (codes_per_slab / (C * C * C))), &e);
            if (rcs[g] == DCT3D_E_STREAM && strstr(m->ctx[g]->err.c_str(), "truncated")) rcs[g] = DCT3D_E_NEED_MORE;
            sb[g] = (uint64_t)byte0 * 8 + e;
        };
        std::vector<std::thread> th;
        for (int g = 2; g < G; g++) th.emplace_back(work, g);
        work(1);
        for (auto &t : th) t.join();
        return first_failure(m, rcs);
    }
    // ---- phase 1: every GPU counts its part of the bytes ----------------------------------------------------------
    std::vector<uint64_t> plo(G + 1), win0(G, 0);
    for (int g = 0; g <= G; g++) plo[g] = g == 0 ? start_bit : g == G ? total_bits : ((start_bit + span * g / G) & ~(uint64_t)31);
    std::vector<PartOpts> po(G);
    auto count = [&](int g, int entry) { rcs[g] = part_count(m->ctx[g], stream, nbytes, plo[g], plo[g + 1], entry, po[g], win0[g]); };
    {
        std::vector<std::thread> th;
        for (int g = 1; g < G; g++) th.emplace_back(count, g, -1);
        count(0, 0);
        for (auto &t : th) t.join();
    }
    int rc = first_failure(m, rcs);
    if (rc) return rc;
    // ---- phase 2: the guesses against the predecessors' overhangs (a recount moves the part's own overhang: in order) -----
    for (int g = 1; g < G; g++) {
        if (po[g].entry_used == po[g - 1].over_out) continue;
        count(g, (int)po[g - 1].over_out + 1);
        if ((rc = first_failure(m, rcs))) return rc;
    }
    std::vector<uint64_t> first(G + 1, 0);               // number of the first code of every part
    for (int g = 0; g < G; g++) first[g + 1] = first[g] + po[g].ncodes;
    // ---- phase 3: the first code of every slab range, inside the part that holds it --------------------------------
    for (int g = 1; g < G; g++) {
        const uint64_t target = (uint64_t)lo[g] * codes_per_slab;
        if (lo[g] == 0) { sb[g] = start_bit; continue; }
        if (target >= first[G]) return mfail(m, DCT3D_E_NEED_MORE, "stream holds %llu codes, range %d starts at code %llu",
                                             (unsigned long long)first[G], g, (unsigned long long)target);
        int owner = 0;
        while (owner + 1 < G && first[owner + 1] <= target) owner++;
        uint64_t bit = 0;
        if ((rc = part_locate(m->ctx[owner], target - first[owner], &bit))) return mfail(m, rc, "GPU %d: %s", m->ctx[owner]->device, m->ctx[owner]->err.c_str());
        sb[g] = win0[owner] + bit;
    }
    return DCT3D_OK;
}

}  // namespace

int dct3d_multi_locate(dct3d_multi *m, const uint8_t *stream, size_t nbytes, int nframes, uint64_t *range_start_bits)
{
    if (!m || !range_start_bits || !stream) return mfail(m, DCT3D_E_INVALID, "null argument");
    const int rc = multi_discover(m, stream, nbytes, 0, nframes, range_start_bits);
    if (rc == DCT3D_E_NEED_MORE) return mfail(m, DCT3D_E_STREAM, "Exp-Golomb stream truncated: %s", m->err.c_str());
    return rc;
}

int dct3d_multi_decode_u8(dct3d_multi *m, const uint8_t *stream, size_t nbytes, int nframes, uint8_t *frames,
                          const uint64_t *range_start_bits)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    if (nframes < 0) return mfail(m, DCT3D_E_INVALID, "negative frame count");
    const int G = (int)m->ctx.size(), C = m->C, nslabs = nframes / C;
    if (nslabs == 0) return DCT3D_OK;
    if (!stream || !frames) return mfail(m, DCT3D_E_INVALID, "null pointer");
    const size_t slab_bytes = (size_t)m->W * m->H * C;
    std::vector<uint64_t> sb(G + 1, 0);
    if (range_start_bits) {
        for (int g = 0; g <= G; g++) sb[g] = range_start_bits[g];
    } else if (G > 1) {
        const int rc = dct3d_multi_locate(m, stream, nbytes, nframes, sb.data());
        if (rc) return rc;
    }
    std::vector<int> rcs(G, 0);
    auto work = [&](int g) {
        int lo, hi;
        slab_range_w(m, nslabs, g, lo, hi);
        if (hi == lo) return;
        // the range's last bit when known (the next non-empty range's first): bounds the H2D copy of the stream
        uint64_t hint = 0;
        for (int j = g + 1; j <= G && !hint; j++) hint = sb[j];
        rcs[g] = dct3d_decode_u8_range(m->ctx[g], stream, nbytes, sb[g], hint > sb[g] ? hint : 0, (hi - lo) * C,
                                       frames + (size_t)lo * slab_bytes, nullptr);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    return first_failure(m, rcs);
}

// ---- streaming over several GPUs: the C codec's batch loop with a carried bit position ------------------------------

int dct3d_multi_stream_begin(dct3d_multi *m)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    m->carry_byte = 0;
    m->carry_bits = 0;
    return dct3d_stream_begin(m->ctx[0]);
}

int dct3d_multi_stream_encode(dct3d_multi *m, const uint8_t *frames, int nframes, int last, uint8_t *out, size_t cap, size_t *nbytes)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    const int G = (int)m->ctx.size(), C = m->C;
    if (G == 1) {
        const int rc = dct3d_stream_encode(m->ctx[0], frames, nframes, last, out, cap, nbytes);
        return rc ? mfail(m, rc, "%s", m->ctx[0]->err.c_str()) : DCT3D_OK;
    }
    if (nframes < 0 || nframes % C) return mfail(m, DCT3D_E_INVALID, "streaming encode needs a multiple of %d frames", C);
    if (!out || cap == 0) return mfail(m, DCT3D_E_INVALID, "null output buffer");
    const int nslabs = nframes / C;
    if (nslabs && !frames) return mfail(m, DCT3D_E_INVALID, "null frame pointer");
    const size_t slab_bytes = (size_t)m->W * m->H * C;
    std::vector<uint64_t> bits(G, 0), start(G + 1, 0);
    std::vector<int> rcs(G, 0);
    std::vector<uint8_t> fb(G, 0);
    HostBarrier bar(G);
    const uint64_t carry = (uint64_t)m->carry_bits;
    auto work = [&](int g) {
        int lo, hi;
        slab_range_w(m, nslabs, g, lo, hi);
        rcs[g] = dct3d_encode_u8_range(m->ctx[g], frames + (size_t)lo * slab_bytes, (hi - lo) * C, &bits[g]);
        bar.wait();
        for (int j = 0; j < G; j++) if (rcs[j]) return;
        uint64_t b = carry;
        for (int j = 0; j < g; j++) b += bits[j];
        rcs[g] = dct3d_encode_u8_place(m->ctx[g], b, last && g == G - 1, out, cap, &fb[g]);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    int rc = first_failure(m, rcs);
    if (rc) return rc;
    start[0] = carry;
    for (int g = 0; g < G; g++) start[g + 1] = start[g] + bits[g];
    const uint64_t end = start[G];
    const size_t full = (size_t)(end / 8), give = last ? full + 1 : full;
    if (give > cap || (end % 8 && full >= cap)) return mfail(m, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", full + 1, cap);
    if (end == carry) {                                          // no slab at all: the carried byte is all there is
        if (carry || last) out[0] = m->carry_byte;
    } else {
        // the first range starts inside the carried byte, the others inside their predecessor's last byte
        for (int g = 0; g < G; g++) {
            if (bits[g] == 0 || start[g] % 8 == 0) continue;
            if (start[g] / 8 == 0 && start[g] == carry) out[0] = (uint8_t)(m->carry_byte | fb[g]);
            else out[start[g] / 8] |= fb[g];
        }
    }
    m->carry_byte = last ? 0 : (end % 8 ? out[full] : 0);
    m->carry_bits = last ? 0 : (int)(end % 8);
    if (nbytes) *nbytes = give;
    return DCT3D_OK;
}

int dct3d_multi_stream_decode(dct3d_multi *m, const uint8_t *in, size_t nbytes, uint64_t *bitpos, int nframes, uint8_t *frames)
{
    if (!m) return mfail(nullptr, DCT3D_E_INVALID, "null context");
    const int G = (int)m->ctx.size(), C = m->C;
    if (G == 1) {
        const int rc = dct3d_stream_decode(m->ctx[0], in, nbytes, bitpos, nframes, frames);
        return rc ? mfail(m, rc, "%s", m->ctx[0]->err.c_str()) : DCT3D_OK;
    }
    if (!bitpos) return mfail(m, DCT3D_E_INVALID, "null bit position");
    const int nslabs = nframes / C;
    if (nslabs == 0) return DCT3D_OK;
    if (!in || !frames) return mfail(m, DCT3D_E_INVALID, "null pointer");
    const size_t slab_bytes = (size_t)m->W * m->H * C;
    std::vector<uint64_t> sb(G + 1, 0), ends(G, 0);
    int rc = multi_discover(m, in, nbytes, *bitpos, nframes, sb.data());
    if (rc) return rc;                                           // DCT3D_E_NEED_MORE: nothing has changed
    std::vector<int> rcs(G, 0);
    auto work = [&](int g) {
        int lo, hi;
        slab_range_w(m, nslabs, g, lo, hi);
        if (hi == lo) return;
        dct3d_ctx *ctx = m->ctx[g];
        if ((rcs[g] = bind(ctx))) return;
        uint64_t hint = 0;
        for (int j = g + 1; j < G && !hint; j++) hint = sb[j] > sb[g] ? sb[j] : 0;
        const size_t byte0 = (size_t)(sb[g] / 8);
        const size_t upto = hint ? std::min<size_t>(nbytes, (size_t)(hint / 8) + 1) : nbytes;
        uint64_t e = 0;
        rcs[g] = byte0 < upto ? pipe_decode(ctx, in + byte0, upto - byte0, sb[g] % 8, (hi - lo) * C, frames + (size_t)lo * slab_bytes, &e)
                              : fail(ctx, DCT3D_E_NEED_MORE, "stream holds no data at the range's start bit");
        ends[g] = (uint64_t)byte0 * 8 + e;
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back(work, g);
    work(0);
    for (auto &t : th) t.join();
    for (int g = 0; g < G; g++) if (rcs[g] == DCT3D_E_NEED_MORE) return mfail(m, DCT3D_E_NEED_MORE, "GPU %d: %s", m->ctx[g]->device, m->ctx[g]->err.c_str());
    if ((rc = first_failure(m, rcs))) return rc;
    for (int g = G - 1; g >= 0; g--) if (ends[g]) { *bitpos = ends[g]; break; }
    return DCT3D_OK;
}

int dct3d_stream_begin(dct3d_ctx *ctx)
{
    if (!ctx) return fail(nullptr, DCT3D_E_INVALID, "null context");
    ctx->carry_byte = 0;
    ctx->carry_bits = 0;
    return DCT3D_OK;
}

int dct3d_stream_encode(dct3d_ctx *ctx, const uint8_t *frames, int nframes, int last, uint8_t *out, size_t cap, size_t *nbytes)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    if (nframes % ctx->C) return fail(ctx, DCT3D_E_INVALID, "streaming encode needs a multiple of %d frames", ctx->C);
    if (!out) return fail(ctx, DCT3D_E_INVALID, "null output buffer");
    if (frame_bytes(ctx, nframes) && !frames) return fail(ctx, DCT3D_E_INVALID, "null frame pointer");
    const size_t dcap = ((cap + 3) & ~(size_t)3) + 64;
    // any number of slabs per call: they flow through the chunk pipeline (pipe_encode) and continue the carried byte
    uint64_t end = 0;
    if ((rc = pipe_encode(ctx, frames, nframes, dcap, nullptr, 0, &end, (unsigned)ctx->carry_bits, ctx->carry_byte))) return rc;
    const size_t full = (size_t)(end / 8);
    const size_t give = last ? full + 1 : full;
    if (give > cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", give, cap);
    if (give) D2H(ctx, out, ctx->bits.p, give);
    D2H(ctx, ctx->h_byte + 8, (uint8_t *)ctx->bits.p + full, 1);
    SYNC(ctx);
    ctx->carry_byte = last ? 0 : ctx->h_byte[8];
    ctx->carry_bits = last ? 0 : (int)(end % 8);
    if (nbytes) *nbytes = give;
    return DCT3D_OK;
}

int dct3d_stream_decode(dct3d_ctx *ctx, const uint8_t *in, size_t nbytes, uint64_t *bitpos, int nframes, uint8_t *frames)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    if (!bitpos) return fail(ctx, DCT3D_E_INVALID, "null bit position");
    if (frame_bytes(ctx, nframes) == 0) return DCT3D_OK;
    if (!in || !frames) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    if ((uint64_t)nbytes * 8 <= *bitpos) return fail(ctx, DCT3D_E_NEED_MORE, "stream holds no data past the start bit");
    const size_t byte0 = (size_t)(*bitpos / 8);
    uint64_t end = 0;
    if ((rc = pipe_decode(ctx, in + byte0, nbytes - byte0, *bitpos % 8, nframes, frames, &end))) return rc;   // DCT3D_E_NEED_MORE changes nothing
    *bitpos = (uint64_t)byte0 * 8 + end;
    return DCT3D_OK;
}

int dct3d_forward_f32(dct3d_ctx *ctx, const float *in, float *out, int nslabs)
{
    if (!ctx || nslabs < 0) return fail(ctx, DCT3D_E_INVALID, "bad argument");
    return transform_host<float>(ctx, in, out, (size_t)ctx->W * ctx->H * ctx->C * nslabs, nslabs, dct3d_forward_f32_dev);
}
int dct3d_inverse_f32(dct3d_ctx *ctx, const float *in, float *out, int nslabs)
{
    if (!ctx || nslabs < 0) return fail(ctx, DCT3D_E_INVALID, "bad argument");
    return transform_host<float>(ctx, in, out, (size_t)ctx->W * ctx->H * ctx->C * nslabs, nslabs, dct3d_inverse_f32_dev);
}
int dct3d_forward_f64(dct3d_ctx *ctx, const double *in, double *out, int nframes)
{
    if (!ctx || nframes < 0) return fail(ctx, DCT3D_E_INVALID, "bad argument");
    return transform_host<double>(ctx, in, out, frame_bytes(ctx, nframes), nframes, dct3d_forward_f64_dev);
}
int dct3d_inverse_f64(dct3d_ctx *ctx, const double *in, double *out, int nframes)
{
    if (!ctx || nframes < 0) return fail(ctx, DCT3D_E_INVALID, "bad argument");
    return transform_host<double>(ctx, in, out, frame_bytes(ctx, nframes), nframes, dct3d_inverse_f64_dev);
}

int dct3d_quantize_u8(dct3d_ctx *ctx, const uint8_t *frames, int nframes, int16_t *qcubes)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const size_t n = frame_bytes(ctx, nframes);
    if (n == 0) return DCT3D_OK;
    if (!frames || !qcubes) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    CU_CHECK(ctx, ctx->frames.reserve(n + 16));
    CU_CHECK(ctx, ctx->q.reserve(n * 2));
    H2D(ctx, ctx->frames.p, frames, n);
    if ((rc = dct3d_quantize_u8_dev(ctx, ctx->frames.p, nframes, ctx->q.p, nullptr))) return rc;
    D2H(ctx, qcubes, ctx->q.p, n * 2);
    SYNC(ctx);
    return DCT3D_OK;
}

int dct3d_reconstruct_i16(dct3d_ctx *ctx, const int16_t *qcubes, int nframes, uint8_t *frames)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if ((rc = check_frames(ctx, nframes))) return rc;
    const size_t n = frame_bytes(ctx, nframes);
    if (n == 0) return DCT3D_OK;
    if (!frames || !qcubes) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    CU_CHECK(ctx, ctx->frames.reserve(n + 16));
    CU_CHECK(ctx, ctx->q.reserve(n * 2));
    H2D(ctx, ctx->q.p, qcubes, n * 2);
    if ((rc = dct3d_reconstruct_i16_dev(ctx, ctx->q.p, nframes, ctx->frames.p, nullptr))) return rc;
    D2H(ctx, frames, ctx->frames.p, n);
    SYNC(ctx);
    return DCT3D_OK;
}

int dct3d_eg_encode_i16(dct3d_ctx *ctx, const int16_t *qcubes, size_t ncubes, uint64_t start_bit,
                        uint8_t *stream, size_t cap, uint64_t *end_bit)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!stream || cap == 0) return fail(ctx, DCT3D_E_INVALID, "null stream buffer");
    const size_t cs = (size_t)ctx->C * ctx->C * ctx->C;
    const size_t first = (size_t)(start_bit / 8);
    if (first >= cap) return fail(ctx, DCT3D_E_OVERFLOW, "start bit beyond the buffer");
    const size_t dcap = ((cap + 3) & ~(size_t)3) + 64;
    CU_CHECK(ctx, ctx->q.reserve(ncubes * cs * 2 + 16));
    CU_CHECK(ctx, ctx->bits.reserve(dcap));
    if (ncubes) H2D(ctx, ctx->q.p, qcubes, ncubes * cs * 2);
    // bytes up to and including the partial byte come from the caller
    CU_CHECK(ctx, cudaMemsetAsync(ctx->bits.p, 0, ((first + 4) & ~(size_t)3), ctx->stream));
    H2D(ctx, ctx->bits.p, stream, first + ((start_bit % 8) ? 1 : 0));
    uint64_t end = 0;
    if ((rc = dct3d_eg_encode_i16_dev(ctx, ctx->q.p, ncubes, start_bit, ctx->bits.p, dcap, &end, nullptr))) return rc;
    const size_t nb = (size_t)(end / 8) + 1;
    if (nb > cap) return fail(ctx, DCT3D_E_OVERFLOW, "stream needs %zu bytes, buffer has %zu", nb, cap);
    D2H(ctx, stream + first, (uint8_t *)ctx->bits.p + first, nb - first);
    SYNC(ctx);
    if (end_bit) *end_bit = end;
    return DCT3D_OK;
}

int dct3d_eg_locate(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit, size_t ncubes, uint64_t *end_bit)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (ncubes == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!stream) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    const size_t padded = ((nbytes + 3) & ~(size_t)3) + 8;
    CU_CHECK(ctx, ctx->bits.reserve(padded));
    CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)ctx->bits.p + (nbytes & ~(size_t)3), 0, padded - (nbytes & ~(size_t)3), ctx->stream));
    H2D(ctx, ctx->bits.p, stream, nbytes);
    rc = dct3d_eg_locate_dev(ctx, ctx->bits.p, nbytes, start_bit, ncubes, end_bit, nullptr);
    if (rc == DCT3D_E_NEED_MORE) return fail(ctx, DCT3D_E_STREAM, "Exp-Golomb stream truncated: %s", ctx->err.c_str());
    return rc;
}

int dct3d_eg_decode_i16(dct3d_ctx *ctx, const uint8_t *stream, size_t nbytes, uint64_t start_bit,
                        size_t ncubes, int16_t *qcubes, uint64_t *end_bit)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (ncubes == 0) { if (end_bit) *end_bit = start_bit; return DCT3D_OK; }
    if (!stream || !qcubes) return fail(ctx, DCT3D_E_INVALID, "null pointer");
    const size_t cs = (size_t)ctx->C * ctx->C * ctx->C;
    const size_t padded = ((nbytes + 3) & ~(size_t)3) + 8;
    CU_CHECK(ctx, ctx->bits.reserve(padded));
    CU_CHECK(ctx, ctx->q.reserve(ncubes * cs * 2 + 16));
    CU_CHECK(ctx, cudaMemsetAsync((uint8_t *)ctx->bits.p + (nbytes & ~(size_t)3), 0, padded - (nbytes & ~(size_t)3), ctx->stream));
    H2D(ctx, ctx->bits.p, stream, nbytes);
    uint64_t end = 0;
    rc = dct3d_eg_decode_i16_dev(ctx, ctx->bits.p, nbytes, start_bit, ncubes, ctx->q.p, &end, nullptr);
    if (rc == DCT3D_E_NEED_MORE) return fail(ctx, DCT3D_E_STREAM, "Exp-Golomb stream truncated: %s", ctx->err.c_str());
    if (rc) return rc;
    D2H(ctx, qcubes, ctx->q.p, ncubes * cs * 2);
    SYNC(ctx);
    if (end_bit) *end_bit = end;
    return DCT3D_OK;
}

}  // extern "C"
