// C ABI of the B200-native ResEnc U-Net hot path (see include/resenc_b200.h for the contract
// and the reference call sites each entry point replaces).  Host side only: argument checks,
// launch geometry, TMA descriptor encoding.  Single translation unit: all kernels are included.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "../../include/resenc_b200.h"
#include "common.cuh"
#include "conv_tc5.cuh"
#include "conv_tc5t.cuh"
#include "conv_slab.cuh"
#include "loss.cuh"
#include "conv_generic.cuh"
#include "wgrad_tc5.cuh"
#include "wgrad2_tc5.cuh"
#include "elementwise.cuh"
#include "split.cuh"
#include "blend.cuh"
#include "optim.cuh"   // after elementwise.cuh: the fused update + pack kernel shares the pack store phase

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return RB_OK;
}

#define RB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail(RB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int num_sms() {
    static int n = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    });
    return n;
}

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda
// dependency, so the library also loads on a machine without a driver).
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ---- CUtensorMap cache ("the library allocates nothing persistent except cached CUtensorMaps keyed by
// (pointer, shape)").  A training step encodes ~2000 descriptors; the caching allocator hands the same buffers to the
// same layers step after step, so after the first step every descriptor is a table hit (~0.1 us) instead of a driver
// call (~1 us).  Only matters for eager launches - a CUDA-graph replay carries its descriptors as kernel parameters.
struct TmapKey {
    uint64_t ptr, dims[5], strides[4];
    uint32_t box[5], estr[5], rank, dtype, swizzle, l2;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");

EncodeTiledFn g_real_encode = nullptr;
std::mutex g_tmap_mutex;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>* g_tmap_cache = nullptr;
std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};

CUresult encode_cached(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                       const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                       CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    static const bool off = getenv("RESENC_NO_TMAP_CACHE") != nullptr;
    if (off || rank > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || oob != CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        return g_real_encode(out, dt, rank, ptr, dims, strides, box, estr, il, sw, l2, oob);
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.ptr = (uint64_t)(uintptr_t)ptr; k.rank = rank; k.dtype = (uint32_t)dt; k.swizzle = (uint32_t)sw; k.l2 = (uint32_t)l2;
    for (cuuint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; }
    for (cuuint32_t i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
    {
        std::lock_guard<std::mutex> lock(g_tmap_mutex);
        if (!g_tmap_cache) g_tmap_cache = new std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>();
        auto it = g_tmap_cache->find(k);
        if (it != g_tmap_cache->end()) {
            memcpy(out, &it->second, sizeof(CUtensorMap));
            g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
            return CUDA_SUCCESS;
        }
    }
    CUresult r = g_real_encode(out, dt, rank, ptr, dims, strides, box, estr, il, sw, l2, oob);
    if (r == CUDA_SUCCESS) {
        std::lock_guard<std::mutex> lock(g_tmap_mutex);
        if (g_tmap_cache->size() > 32768) g_tmap_cache->clear();      // bounded: 8 MB
        g_tmap_cache->emplace(k, *out);
        g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
    }
    return r;
}

EncodeTiledFn encode_tiled_fn() {
    // cuTensorMapEncodeTiled is a driver call: the calling thread (e.g. an autograd worker that has not touched the
    // runtime yet) must have the primary context bound, which any runtime call does
    thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        g_real_encode = fn;
    });
    return fn ? encode_cached : nullptr;
}

CUtensorMapSwizzle swizzle_for(int kw) {
    return kw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

struct Tc5Plan {
    bool ok = false;
    int KW = 0, Ntile = 0, nTilesN = 0, stages = 0;
    int tw = 0, th = 0, td = 0, tn = 0, tilesW = 0, tilesH = 0, tilesD = 0, tilesNB = 0;
    size_t smem = 0;
    long long tiles = 0;
    int ksteps = 0;
    bool statSmem = false;
    int tps = 1;
};

Tc5Plan plan_tc5(const RbConvDesc& d) {
    Tc5Plan pl;
    if (d.nsrc < 1 || d.nsrc > 2) return pl;
    if (d.srcC0 % 16 != 0 || (d.nsrc == 2 && d.srcC1 % 16 != 0)) return pl;
    if (d.Nout % 32 != 0) return pl;
    if (d.mode == 1 && (d.psC % 32 != 0)) return pl;
    if (d.outC1 > 0 && d.outC0 % 32 != 0) return pl;
    if (d.outC0 % 8 != 0 || d.outC1 % 8 != 0) return pl;
    int kw = 64;
    while (kw > 16 && (d.srcC0 % kw != 0 || (d.nsrc == 2 && d.srcC1 % kw != 0))) kw >>= 1;
    pl.KW = kw;
    if (d.Nout <= 256) pl.Ntile = d.Nout;
    else if (d.Nout % 256 == 0) pl.Ntile = 256;
    else if (d.Nout % 128 == 0) pl.Ntile = 128;
    else if (d.Nout % 64 == 0) pl.Ntile = 64;
    else pl.Ntile = 32;
    pl.nTilesN = d.Nout / pl.Ntile;
    // tile box: power-of-two factorisation of 128 output voxels minimising the tile count
    long long best = -1;
    for (int tw = 1; tw <= 128; tw <<= 1)
        for (int th = 1; tw * th <= 128; th <<= 1)
            for (int td = 1; tw * th * td <= 128; td <<= 1) {
                const int tn = 128 / (tw * th * td);
                if ((tw - 1) * d.istrW + 1 > 256 || (th - 1) * d.istrH + 1 > 256 || (td - 1) * d.istrD + 1 > 256 || tn > 256) continue;
                const long long t = (long long)((d.OW + tw - 1) / tw) * ((d.OH + th - 1) / th) * ((d.OD + td - 1) / td) *
                                    ((d.NB + tn - 1) / tn);
                // ties: keep samples apart (tn == 1) first, then prefer the widest tw (longer contiguous runs
                // for TMA and the epilogue stores)
                if (best < 0 || t < best || (t == best && (tn < pl.tn || (tn == pl.tn && tw > pl.tw)))) {
                    best = t;
                    pl.tw = tw; pl.th = th; pl.td = td; pl.tn = tn;
                }
            }
    if (best < 0) return pl;
    pl.tilesW = (d.OW + pl.tw - 1) / pl.tw;
    pl.tilesH = (d.OH + pl.th - 1) / pl.th;
    pl.tilesD = (d.OD + pl.td - 1) / pl.td;
    pl.tilesNB = (d.NB + pl.tn - 1) / pl.tn;
    pl.tiles = best * pl.nTilesN;
    // narrow-K layers (K chunk <= 32 channels): put all tapW taps of a row into one pipeline stage
    pl.tps = (pl.KW <= 32 && d.tapW == 3) ? 3 : 1;
    const size_t stageBytes = (size_t)(128 + pl.Ntile) * pl.KW * 2 * pl.tps;
    // per-epilogue-warp statistics accumulators 4 x [2][NB][Nout] fp32 live behind the stages when they fit in 16 KB
    const size_t statBytes = (size_t)8 * d.NB * d.Nout * sizeof(float);
    pl.statSmem = statBytes <= 16 * 1024;
    const size_t reserve = pl.statSmem ? statBytes : 0;
    int st = (int)((200 * 1024 - reserve) / stageBytes);
    if (st > 8) st = 8;
    if (st < 2) return pl;
    pl.stages = st;
    pl.smem = 1024 /*align slack*/ + 1024 /*barriers*/ + (size_t)st * stageBytes + reserve;
    const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
    pl.ksteps = d.tapD * d.tapH * d.tapW * (ctot / pl.KW);
    pl.ok = true;
    return pl;
}

// ---- "weights on M, 256 voxels on N" orientation (conv_tc5t.cuh): mode 0, Nout <= 128 -----------------------
struct Tc5tPlan {
    bool ok = false;
    int KW = 0, lw = 0, lh = 0, ld = 0, tn = 0, tilesW = 0, tilesH = 0, tilesD = 0, tilesNB = 0, tilesM = 0, stages = 0;
    long long tiles = 0;
    size_t smem = 0;
    bool statSmem = false;
    int splitK = 1, tapsPer = 27;
    bool hm = false;   // h-major 32 x 8 tile: one X load per (kd, kh) serves the three kw taps (conv_tc5t.cuh)
};

Tc5tPlan plan_tc5t(const RbConvDesc& d) {
    Tc5tPlan pl;
    if (d.mode != 0 || d.nsrc < 1 || d.nsrc > 2) return pl;
    if (d.srcC0 % 16 != 0 || (d.nsrc == 2 && d.srcC1 % 16 != 0)) return pl;
    if (d.Nout % 8 != 0) return pl;
    int kw = 64;
    while (kw > 16 && (d.srcC0 % kw != 0 || (d.nsrc == 2 && d.srcC1 % kw != 0))) kw >>= 1;
    pl.KW = kw;
    pl.tilesM = (d.Nout + 127) / 128;
    long long best = -1;
    int btn = 0, btw = 0;
    for (int lw = 0; lw <= 8; ++lw)
        for (int lh = 0; lw + lh <= 8; ++lh)
            for (int ld = 0; lw + lh + ld <= 8; ++ld) {
                const int tw = 1 << lw, th = 1 << lh, td = 1 << ld, tn = 256 >> (lw + lh + ld);
                if ((tw - 1) * d.istrW + 1 > 256 || (th - 1) * d.istrH + 1 > 256 || (td - 1) * d.istrD + 1 > 256) continue;
                const long long t = (long long)((d.OW + tw - 1) / tw) * ((d.OH + th - 1) / th) * ((d.OD + td - 1) / td) *
                                    ((d.NB + tn - 1) / tn);
                if (best < 0 || t < best || (t == best && (tn < btn || (tn == btn && tw > btw)))) {
                    best = t; btn = tn; btw = tw;
                    pl.lw = lw; pl.lh = lh; pl.ld = ld; pl.tn = tn;
                }
            }
    if (best < 0) return pl;
    const int tw = 1 << pl.lw, th = 1 << pl.lh, td = 1 << pl.ld;
    pl.tilesW = (d.OW + tw - 1) / tw; pl.tilesH = (d.OH + th - 1) / th; pl.tilesD = (d.OD + td - 1) / td;
    pl.tilesNB = (d.NB + pl.tn - 1) / pl.tn;
    pl.tiles = best * pl.tilesM;
    const size_t stageBytes = (size_t)(256 + 128) * pl.KW * 2;
    const size_t statBytes = (size_t)8 * d.NB * d.Nout * sizeof(float);
    pl.statSmem = statBytes <= 16 * 1024;
    const size_t reserve = pl.statSmem ? statBytes : 0;
    int st = (int)((200 * 1024 - reserve) / stageBytes);
    if (st > 8) st = 8;
    if (st < 2) return pl;
    pl.stages = st;
    pl.smem = 1024 + 1024 + (size_t)st * stageBytes + reserve;
    // deep layers: a handful of voxel tiles with a very long K loop -> split the taps over the chip
    const int ntaps = d.tapD * d.tapH * d.tapW;
    const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
    pl.tapsPer = ntaps;
    pl.splitK = 1;
    static const int split_max_tiles = getenv("RESENC_SPLIT_MAX_TILES") ? atoi(getenv("RESENC_SPLIT_MAX_TILES")) : 74;
    // Up to half a wave of tiles (4^3 / 8^3: 4-16 tiles, the 256-channel layers at 16^3: 64 tiles on 148 SMs): split the
    // taps over the chip.  Every (tile, tap slice) item stores its partial tile into its own workspace slice and
    // split_finish_kernel sums them (no atomics), so the cost model is simply the longest per-SM queue: items per
    // persistent CTA x (taps per item + ~2 taps worth of pipeline fill and epilogue), + 1 for the finish pass.
    const bool plain = d.ostrD == 1 && d.ostrH == 1 && d.ostrW == 1 && d.ooffD == 0 && d.ooffH == 0 && d.ooffW == 0 &&
                       d.FD == d.OD && d.FH == d.OH && d.FW == d.OW && d.Nout % 32 == 0 && (d.outC1 == 0 || d.outC0 % 32 == 0);
    if (pl.tiles <= split_max_tiles && ntaps >= 8 && ctot >= 256 && plain) {
        long long bestCost = (long long)ntaps + 2;     // unsplit: one item of ntaps taps per CTA
        for (int per = 1; per < ntaps; ++per) {
            const int sk = (ntaps + per - 1) / per;
            const long long items = pl.tiles * sk;
            const long long cost = ((items + num_sms() - 1) / num_sms()) * (per + 2) + 1;
            if (cost < bestCost) { bestCost = cost; pl.tapsPer = per; pl.splitK = sk; }
        }
    }
    // h-major tile for plain 3-tap-wide stride-1 convolutions on rows that are whole multiples of 32 voxels: the per-tap
    // gather is TMA-row-rate bound at 64/128-byte channel rows (cycle counters: 18.6 k cycles of loads against 13.5 k of
    // MMAs per 64->64 tile), this layout loads 9 boxes of 272 rows per tile instead of 27 boxes of 256
    static const bool no_hm = getenv("RESENC_NO_TC5T_HM") != nullptr;
    if (!no_hm && pl.splitK == 1 && d.mode == 0 && d.tapW == 3 && d.offW == -1 && d.istrD == 1 && d.istrH == 1 && d.istrW == 1 &&
        d.ostrD == 1 && d.ostrH == 1 && d.ostrW == 1 && d.ooffD == 0 && d.ooffH == 0 && d.ooffW == 0 && d.FD == d.OD &&
        d.FH == d.OH && d.FW == d.OW && d.OW % 32 == 0 && d.OW == d.IW && d.Nout <= 128 && pl.KW >= 32) {
        const int wRows = d.Nout;
        const size_t sb = ((size_t)34 * 8 + 2 * wRows + 128) * pl.KW * 2;
        int sth = (int)((200 * 1024 - reserve) / sb);
        if (sth > 6) sth = 6;
        if (sth >= 2 && sb % 1024 == 0) {
            pl.hm = true;
            pl.lw = 5; pl.lh = 3; pl.ld = 0; pl.tn = 1;
            pl.tilesW = d.OW / 32; pl.tilesH = (d.OH + 7) / 8; pl.tilesD = d.OD; pl.tilesNB = d.NB;
            pl.tiles = (long long)pl.tilesW * pl.tilesH * pl.tilesD * pl.tilesNB * pl.tilesM;
            pl.stages = sth;
            pl.smem = 1024 + 1024 + (size_t)sth * sb + reserve;
        }
    }
    pl.ok = true;
    return pl;
}

bool auto_prefers_tc5t(const RbConvDesc& d, const Tc5tPlan& pl) {
    static const bool off = getenv("RESENC_NO_TC5T") != nullptr;
    if (!pl.ok || off) return false;
    if (pl.splitK > 1) return true;           // deep layers: tap-split tcgen05 instead of split-K mma.sync
    if (pl.tiles < 64 || d.Nout > 128) return false;
    // Measured (profiles/r1_convbench.json): the N = 256 orientation wins when a tile carries enough MMA work to
    // hide its epilogue (one warp per 32 output channels) and the K chunks are >= 128-byte TMA rows; 32-channel
    // inputs (64-byte rows, TMA row-rate bound) and few-tap data-gradient classes stay on the voxels-on-M kernel.
    const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
    const int ntaps = d.tapD * d.tapH * d.tapW;
    static const bool force = getenv("RESENC_FORCE_TC5T") != nullptr;   // experiments
    if (force) return true;
    if (ntaps * ctot >= 864 && ctot >= 64) return true;
    // 32-channel full-resolution layers that the slab kernel does not take (W not in {32, 64, 96, 128}: 192^3 patches, the
    // reference's [64, 192, 192]): the h-major tile loads 9 boxes of 272 rows instead of 27 boxes of 256 per 256 voxels,
    // and the per-tap kernels are TMA-row-rate bound at 64-byte rows (1.26 ms against 0.275 for 32->32 @128^3)
    static const bool no_hm32 = getenv("RESENC_NO_TC5T_HM32") != nullptr;
    return !no_hm32 && pl.hm && ntaps == 27 && ctot >= 32 && d.Nout >= 32;
}

int launch_tc5t(const RbConvDesc& d, const Tc5tPlan& pl, const void* src0, const void* src1, const void* w, void* out0,
                void* out1, float* stat_sum, float* stat_sq, float* ws, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    rb::Tc5tConvParams p;
    memset(&p, 0, sizeof(p));
    const int tw = 1 << pl.lw, th = 1 << pl.lh, td = 1 << pl.ld;
    const void* srcs[2] = {src0, src1};
    const int srcC[2] = {d.srcC0, d.srcC1};
    for (int s = 0; s < d.nsrc; ++s) {
        const cuuint64_t C = (cuuint64_t)srcC[s];
        cuuint64_t dims[5] = {C, (cuuint64_t)d.IW, (cuuint64_t)d.IH, (cuuint64_t)d.ID, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {C * 2, C * 2 * d.IW, C * 2 * d.IW * d.IH, C * 2 * d.IW * d.IH * d.ID};
        cuuint32_t box[5] = {(cuuint32_t)pl.KW, (cuuint32_t)((tw - 1) * d.istrW + 1), (cuuint32_t)((th - 1) * d.istrH + 1),
                             (cuuint32_t)((td - 1) * d.istrD + 1), (cuuint32_t)pl.tn};
        cuuint32_t estr[5] = {1, (cuuint32_t)d.istrW, (cuuint32_t)d.istrH, (cuuint32_t)d.istrD, 1};
        if (pl.hm) {   // dimension order (c, h, w, d, n): the box lands in shared memory as [w][h][c]
            dims[1] = (cuuint64_t)d.IH; dims[2] = (cuuint64_t)d.IW;
            strides[0] = C * 2 * d.IW; strides[1] = C * 2;
            box[1] = 8; box[2] = 34; box[3] = 1; box[4] = 1;
            estr[1] = estr[2] = estr[3] = 1;
        }
        CUresult r = enc(&p.mapX[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(srcs[s]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.KW), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(X%d) failed: %d", s, (int)r);
    }
    {
        const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
        const int ntaps = d.tapD * d.tapH * d.tapW;
        cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)d.Nout, (cuuint64_t)ntaps};
        cuuint64_t strides[2] = {(cuuint64_t)ctot * 2, (cuuint64_t)ctot * 2 * d.Nout};
        cuuint32_t box[3] = {(cuuint32_t)pl.KW, (cuuint32_t)(d.Nout < 128 ? d.Nout : 128), (cuuint32_t)(pl.hm ? 3 : 1)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.KW), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
    p.nsrc = d.nsrc; p.srcC[0] = d.srcC0; p.srcC[1] = d.nsrc == 2 ? d.srcC1 : 0;
    p.KW = pl.KW;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW; p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.istrD = d.istrD; p.istrH = d.istrH; p.istrW = d.istrW;
    p.lw = pl.lw; p.lh = pl.lh; p.ld = pl.ld;
    p.tilesW = pl.tilesW; p.tilesH = pl.tilesH; p.tilesD = pl.tilesD; p.tilesNB = pl.tilesNB; p.tilesM = pl.tilesM;
    p.OW = d.OW; p.OH = d.OH; p.OD = d.OD; p.NB = d.NB; p.Nout = d.Nout;
    p.ostrD = d.ostrD; p.ostrH = d.ostrH; p.ostrW = d.ostrW; p.ooffD = d.ooffD; p.ooffH = d.ooffH; p.ooffW = d.ooffW;
    p.FD = d.FD; p.FH = d.FH; p.FW = d.FW;
    p.out0 = out0; p.out1 = out1; p.outC0 = d.outC0; p.outC1 = d.outC1; p.outF32 = d.outF32;
    p.stages = pl.stages; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.statSmem = pl.statSmem ? 1 : 0;
    p.wRows = d.Nout < 128 ? d.Nout : 128;
    p.hm = pl.hm ? 1 : 0;
    p.fdTilesM = rb::make_fastdiv(pl.tilesM); p.fdTilesW = rb::make_fastdiv(pl.tilesW);
    p.fdTilesH = rb::make_fastdiv(pl.tilesH); p.fdTilesD = rb::make_fastdiv(pl.tilesD);
    p.splitK = pl.splitK; p.tapsPer = pl.tapsPer; p.fdSplitK = rb::make_fastdiv(pl.splitK); p.ws = ws;
    p.wsSlice = (pl.tiles / pl.tilesM) * 256LL * d.Nout;
    {
        static const int dbg = getenv("RESENC_TC5T_DEBUG") ? atoi(getenv("RESENC_TC5T_DEBUG")) : 0;
        p.debug = dbg & 15;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(rb::tc5t_gather_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(tc5t): %s", cudaGetErrorString(attr_err));
    const long long items = pl.tiles * pl.splitK;
    long long grid = items < num_sms() ? items : num_sms();
    rb::tc5t_gather_conv_kernel<<<(int)grid, rb::TC5T_THREADS, pl.smem, st>>>(p);
    return check_launch("tc5t_gather_conv_kernel");
}

// ---- slab kernel (conv_slab.cuh): 3x3x3 stride-1 pad-1 convolutions between 32-channel tensors at W in {32,64,128} ----
struct SlabPlan {
    bool ok = false;
    int R = 0, lw = 0, DC = 0, hTiles = 0, dChunks = 0;
    long long items = 0;
    size_t smem = 0;
};

SlabPlan plan_slab(const RbConvDesc& d) {
    SlabPlan pl;
    static const bool off = getenv("RESENC_NO_SLAB") != nullptr;
    if (off || d.mode != 0) return pl;
    if (d.tapD != 3 || d.tapH != 3 || d.tapW != 3 || d.offD != -1 || d.offH != -1 || d.offW != -1) return pl;
    if (d.istrD != 1 || d.istrH != 1 || d.istrW != 1 || d.ostrD != 1 || d.ostrH != 1 || d.ostrW != 1) return pl;
    if (d.ooffD != 0 || d.ooffH != 0 || d.ooffW != 0) return pl;
    if (d.OD != d.ID || d.OH != d.IH || d.OW != d.IW || d.FD != d.OD || d.FH != d.OH || d.FW != d.OW) return pl;
    if (d.srcC0 != 32 || (d.nsrc == 2 && d.srcC1 != 32)) return pl;
    if (!((d.Nout == 32 && d.outC0 == 32 && d.outC1 == 0) || (d.Nout == 64 && d.outC0 == 32 && d.outC1 == 32))) return pl;
    if (d.nsrc == 2 && !d.outF32) return pl;   // the second source accumulates into an fp32 / fp16 destination
    // a tile = R full rows of W voxels on the N side of the MMA: N = R * W must be a multiple of 64 (two warp groups x
    // 64-column hand-offs) and <= 256, rows must start at multiples of 32 columns; W = 192 (R = 1) needs four 36 KB
    // planes + 72 KB of weights + staging > 227 KB of shared memory and stays on the h-major gather kernel
    if (d.IW != 32 && d.IW != 64 && d.IW != 96 && d.IW != 128) return pl;
    pl.R = 256 / d.IW;
    if (d.IH % pl.R != 0) return pl;
    pl.lw = d.IW == 32 ? 5 : d.IW == 64 ? 6 : 7;   // (informational; the kernel indexes with R and W)
    pl.hTiles = d.IH / pl.R;
    // output planes per work item: whole waves of CTAs first, then fewer halo re-loads
    double bestScore = -1.0;
    for (int dc = 4; dc <= 32; dc <<= 1) {
        const int chunks = (d.OD + dc - 1) / dc;
        const long long items = (long long)d.NB * pl.hTiles * chunks;
        const long long waves = (items + num_sms() - 1) / num_sms();
        const double eff = (double)items / (double)(waves * num_sms()) * ((double)dc / (dc + 0.5));
        if (eff > bestScore) { bestScore = eff; pl.DC = dc; pl.dChunks = chunks; pl.items = items; }
    }
    if ((long long)d.NB * pl.hTiles * d.OD < 2LL * num_sms()) return pl;   // too little work for a z-marching CTA per SM
    const size_t slot = (size_t)(pl.R + 2) * d.IW * 64;
    pl.smem = 1024 + 1024 + rb::SLAB_WBYTES + 4 * slot + rb::SLAB_OBYTES;
    if (pl.smem > 227 * 1024) return pl;
    pl.ok = true;
    return pl;
}

int launch_slab_one(const RbConvDesc& d, const SlabPlan& pl, const void* src, int c0, int m0, const void* w, void* out, int accumulate,
                    float* stat_sum, float* stat_sq, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    rb::SlabConvParams p;
    memset(&p, 0, sizeof(p));
    {
        cuuint64_t dims[5] = {32, (cuuint64_t)d.IW, (cuuint64_t)d.IH, (cuuint64_t)d.ID, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {64, (cuuint64_t)64 * d.IW, (cuuint64_t)64 * d.IW * d.IH, (cuuint64_t)64 * d.IW * d.IH * d.ID};
        cuuint32_t box[5] = {32, (cuuint32_t)d.IW, (cuuint32_t)(pl.R + 2), 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&p.mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(src), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(slab X) failed: %d", (int)r);
    }
    {
        const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
        cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)d.Nout, 27};
        cuuint64_t strides[2] = {(cuuint64_t)ctot * 2, (cuuint64_t)ctot * 2 * d.Nout};
        cuuint32_t box[3] = {32, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(slab W) failed: %d", (int)r);
    }
    p.c0 = c0; p.m0 = m0;
    p.W = d.IW; p.H = d.IH; p.D = d.ID; p.NB = d.NB; p.lw = pl.lw; p.R = pl.R; p.DC = pl.DC;
    p.hTiles = pl.hTiles; p.dChunks = pl.dChunks;
    p.fdH = rb::make_fastdiv(pl.hTiles); p.fdDC = rb::make_fastdiv(pl.dChunks);
    p.out = out;
    p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.statPitch = d.Nout; p.statC0 = m0;
    {
        static const int dbg = getenv("RESENC_SLAB_DEBUG") ? atoi(getenv("RESENC_SLAB_DEBUG")) : 0;
        p.debug = dbg;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        const void* fns[10] = {(const void*)rb::slab_conv_kernel<0, 256>, (const void*)rb::slab_conv_kernel<1, 256>,
                               (const void*)rb::slab_conv_kernel<2, 256>, (const void*)rb::slab_conv_kernel<3, 256>,
                               (const void*)rb::slab_conv_kernel<4, 256>, (const void*)rb::slab_conv_kernel<0, 192>,
                               (const void*)rb::slab_conv_kernel<1, 192>, (const void*)rb::slab_conv_kernel<2, 192>,
                               (const void*)rb::slab_conv_kernel<3, 192>, (const void*)rb::slab_conv_kernel<4, 192>};
        for (int i = 0; i < 10 && attr_err == cudaSuccess; ++i)
            attr_err = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(slab): %s", cudaGetErrorString(attr_err));
    const long long grid = pl.items < num_sms() ? pl.items : num_sms();
    // destination element type (RbConvDesc.outF32): 0 bf16, 1 fp32, 2 fp16; the second source of a virtual concat accumulates
    const int mode = d.outF32 == 0 ? 0 : d.outF32 == 1 ? (accumulate ? 2 : 1) : (accumulate ? 4 : 3);
    const bool n192 = pl.R * d.IW == 192;
#define RB_SLAB_LAUNCH(M, N) rb::slab_conv_kernel<M, N><<<(int)grid, rb::SLAB_THREADS, pl.smem, st>>>(p)
#define RB_SLAB_MODES(N) \
    switch (mode) { case 0: RB_SLAB_LAUNCH(0, N); break; case 1: RB_SLAB_LAUNCH(1, N); break; case 2: RB_SLAB_LAUNCH(2, N); break; \
                    case 3: RB_SLAB_LAUNCH(3, N); break; default: RB_SLAB_LAUNCH(4, N); break; }
    if (!n192) { RB_SLAB_MODES(256) } else { RB_SLAB_MODES(192) }
#undef RB_SLAB_MODES
#undef RB_SLAB_LAUNCH
    return check_launch("slab_conv_kernel");
}

// one launch per (destination half, source): the second source of a virtual concat adds into the fp32 destination
int launch_slab(const RbConvDesc& d, const SlabPlan& pl, const void* src0, const void* src1, const void* w, void* out0, void* out1,
                float* stat_sum, float* stat_sq, cudaStream_t st) {
    const void* srcs[2] = {src0, src1};
    void* outs[2] = {out0, out1};
    for (int o = 0; o < d.Nout / 32; ++o)
        for (int s = 0; s < d.nsrc; ++s) {
            const bool last = s == d.nsrc - 1;
            int rc = launch_slab_one(d, pl, srcs[s], s * 32, o * 32, w, outs[o], s > 0 ? 1 : 0, last ? stat_sum : nullptr,
                                     last ? stat_sq : nullptr, st);
            if (rc) return rc;
        }
    return RB_OK;
}

int validate_conv(const RbConvDesc& d) {
    if (d.nsrc < 1 || d.nsrc > 2) return fail(RB_ERR_INVALID, "conv: nsrc must be 1 or 2");
    if (d.srcC0 <= 0 || d.srcC0 % 8 != 0 || (d.nsrc == 2 && (d.srcC1 <= 0 || d.srcC1 % 8 != 0)))
        return fail(RB_ERR_INVALID, "conv: source channels must be positive multiples of 8");
    if (d.Nout <= 0 || d.Nout % 8 != 0) return fail(RB_ERR_INVALID, "conv: Nout must be a positive multiple of 8");
    if (d.NB <= 0 || d.ID <= 0 || d.IH <= 0 || d.IW <= 0 || d.OD <= 0 || d.OH <= 0 || d.OW <= 0)
        return fail(RB_ERR_INVALID, "conv: empty grid");
    if (d.tapD < 1 || d.tapH < 1 || d.tapW < 1 || d.tapD > 3 || d.tapH > 3 || d.tapW > 3)
        return fail(RB_ERR_INVALID, "conv: taps per axis must be 1..3");
    if (d.istrD < 1 || d.istrH < 1 || d.istrW < 1 || d.ostrD < 1 || d.ostrH < 1 || d.ostrW < 1)
        return fail(RB_ERR_INVALID, "conv: strides must be >= 1");
    if (d.outC0 <= 0 || d.outC0 % 8 != 0 || d.outC1 < 0 || d.outC1 % 8 != 0)
        return fail(RB_ERR_INVALID, "conv: destination channels must be multiples of 8");
    if (d.outF32 < 0 || d.outF32 > 2) return fail(RB_ERR_INVALID, "conv: outF32 must be 0 (bf16), 1 (fp32) or 2 (fp16)");
    if (d.mode == 0) {
        if (d.Nout != d.outC0 + d.outC1) return fail(RB_ERR_INVALID, "conv: Nout != outC0 + outC1");
        if ((long long)(d.OD - 1) * d.ostrD + d.ooffD >= d.FD || (long long)(d.OH - 1) * d.ostrH + d.ooffH >= d.FH ||
            (long long)(d.OW - 1) * d.ostrW + d.ooffW >= d.FW || d.ooffD < 0 || d.ooffH < 0 || d.ooffW < 0)
            return fail(RB_ERR_INVALID, "conv: output class grid exceeds the destination");
    } else if (d.mode == 1) {
        if (d.psC <= 0 || d.psC % 8 != 0 || d.psD < 1 || d.psH < 1 || d.psW < 1 || d.psD > 2 || d.psH > 2 || d.psW > 2)
            return fail(RB_ERR_INVALID, "conv: bad pixel-shuffle factors");
        if (d.Nout != d.psC * d.psD * d.psH * d.psW) return fail(RB_ERR_INVALID, "conv: Nout != psC * parities");
        if (d.psC != d.outC0 + d.outC1) return fail(RB_ERR_INVALID, "conv: psC != outC0 + outC1");
        if (d.ostrD != d.psD || d.ostrH != d.psH || d.ostrW != d.psW) return fail(RB_ERR_INVALID, "conv: pixel shuffle needs ostr == ps");
        if (d.OD * d.psD > d.FD || d.OH * d.psH > d.FH || d.OW * d.psW > d.FW) return fail(RB_ERR_INVALID, "conv: pixel shuffle exceeds the destination");
    } else {
        return fail(RB_ERR_INVALID, "conv: mode must be 0 or 1");
    }
    return RB_OK;
}

int generic_splitk(const RbConvDesc& d) {
    if (d.splitK == 1) return 1;
    const long long M = (long long)d.NB * d.OD * d.OH * d.OW;
    const long long ctas = ((M + rb::GC_BM - 1) / rb::GC_BM) * ((d.Nout + rb::GC_BN - 1) / rb::GC_BN);
    const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
    const int nIt = d.tapD * d.tapH * d.tapW * ((ctot + rb::GC_BK - 1) / rb::GC_BK);
    if (d.splitK > 1) return d.splitK < nIt ? d.splitK : nIt;
    if (ctas >= num_sms() || nIt < 16) return 1;
    long long s = (2LL * num_sms() + ctas - 1) / ctas;
    if (s > nIt / 4) s = nIt / 4;
    if (s > 64) s = 64;
    return s < 2 ? 1 : (int)s;
}

bool auto_prefers_tc5(const RbConvDesc& d, const Tc5Plan& pl) {
    if (!pl.ok) return false;
    // a handful of 128-row tiles with a very long K loop is latency bound on one CTA each:
    // the split-K mma.sync path spreads it over the whole chip instead
    if (pl.tiles < 48 && pl.ksteps > 64) return false;
    return true;
}

int launch_tc5(const RbConvDesc& d, const Tc5Plan& pl, const void* src0, const void* src1, const void* w, void* out0,
               void* out1, float* stat_sum, float* stat_sq, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    rb::Tc5ConvParams p;
    memset(&p, 0, sizeof(p));
    const void* srcs[2] = {src0, src1};
    const int srcC[2] = {d.srcC0, d.srcC1};
    for (int s = 0; s < d.nsrc; ++s) {
        const cuuint64_t C = (cuuint64_t)srcC[s];
        cuuint64_t dims[5] = {C, (cuuint64_t)d.IW, (cuuint64_t)d.IH, (cuuint64_t)d.ID, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {C * 2, C * 2 * d.IW, C * 2 * d.IW * d.IH, C * 2 * d.IW * d.IH * d.ID};
        cuuint32_t box[5] = {(cuuint32_t)pl.KW, (cuuint32_t)((pl.tw - 1) * d.istrW + 1), (cuuint32_t)((pl.th - 1) * d.istrH + 1),
                             (cuuint32_t)((pl.td - 1) * d.istrD + 1), (cuuint32_t)pl.tn};
        cuuint32_t estr[5] = {1, (cuuint32_t)d.istrW, (cuuint32_t)d.istrH, (cuuint32_t)d.istrD, 1};
        CUresult r = enc(&p.mapA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(srcs[s]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.KW), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
    }
    {
        const int ctot = d.srcC0 + (d.nsrc == 2 ? d.srcC1 : 0);
        const int ntaps = d.tapD * d.tapH * d.tapW;
        cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)d.Nout, (cuuint64_t)ntaps};
        cuuint64_t strides[2] = {(cuuint64_t)ctot * 2, (cuuint64_t)ctot * 2 * d.Nout};
        cuuint32_t box[3] = {(cuuint32_t)pl.KW, (cuuint32_t)pl.Ntile, (cuuint32_t)pl.tps};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(pl.KW), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    p.nsrc = d.nsrc; p.srcC[0] = d.srcC0; p.srcC[1] = d.nsrc == 2 ? d.srcC1 : 0;
    p.KW = pl.KW;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW;
    p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.istrD = d.istrD; p.istrH = d.istrH; p.istrW = d.istrW;
    p.tw = pl.tw; p.th = pl.th; p.td = pl.td; p.tn = pl.tn;
    p.tilesW = pl.tilesW; p.tilesH = pl.tilesH; p.tilesD = pl.tilesD; p.tilesNB = pl.tilesNB;
    p.OW = d.OW; p.OH = d.OH; p.OD = d.OD; p.NB = d.NB;
    p.Nout = d.Nout; p.Ntile = pl.Ntile; p.nTilesN = pl.nTilesN;
    p.mode = d.mode;
    p.ostrD = d.ostrD; p.ostrH = d.ostrH; p.ostrW = d.ostrW; p.ooffD = d.ooffD; p.ooffH = d.ooffH; p.ooffW = d.ooffW;
    p.FD = d.FD; p.FH = d.FH; p.FW = d.FW;
    p.out0 = out0; p.out1 = out1; p.outC0 = d.outC0; p.outC1 = d.outC1;
    p.psC = d.psC; p.psD = d.psD; p.psH = d.psH; p.psW = d.psW;
    p.stages = pl.stages;
    p.stat_sum = stat_sum; p.stat_sq = stat_sq;
    p.outF32 = d.outF32;
    p.statSmem = pl.statSmem ? 1 : 0;
    p.tps = pl.tps;
    p.fdTilesN = rb::make_fastdiv(pl.nTilesN); p.fdTilesW = rb::make_fastdiv(pl.tilesW);
    p.fdTilesH = rb::make_fastdiv(pl.tilesH); p.fdTilesD = rb::make_fastdiv(pl.tilesD);
    p.fdTw = rb::make_fastdiv(pl.tw); p.fdTwTh = rb::make_fastdiv(pl.tw * pl.th); p.fdTwThTd = rb::make_fastdiv(pl.tw * pl.th * pl.td);
    {
        static const int dbg = getenv("RESENC_TC5_DEBUG") ? atoi(getenv("RESENC_TC5_DEBUG")) : 0;
        p.debug = dbg;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(rb::tc5_gather_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(tc5): %s", cudaGetErrorString(attr_err));
    long long grid = pl.tiles < num_sms() ? pl.tiles : num_sms();
    rb::tc5_gather_conv_kernel<<<(int)grid, rb::TC5_THREADS, pl.smem, st>>>(p);
    return check_launch("tc5_gather_conv_kernel");
}

struct Tw5Plan {
    bool ok = false;
    int aw = 0, bw = 0, aAtoms = 0, bn = 0, aTiles = 0, bTiles = 0;
    int cw = 0, ch = 0, cd = 0, cn = 0, chunksW = 0, chunksH = 0, chunksD = 0, chunksN = 0;
    int splits = 1, chunksPerSplit = 0, stages = 0, accBufs = 1;
    size_t smem = 0;
    bool swap = false;
    int tpi = 1, tapGroups = 1;
    int kbox = 64;
};

int atom_width(int c) { return c % 64 == 0 ? 64 : c % 32 == 0 ? 32 : 16; }

Tw5Plan plan_tw5(const RbWgradDesc& d) {
    Tw5Plan pl;
    const int qct = d.QC0 + (d.nq == 2 ? d.QC1 : 0);
    if (d.PC % 16 != 0 || d.QC0 % 16 != 0 || (d.nq == 2 && d.QC1 % 16 != 0)) return pl;
    if (qct % 32 != 0) return pl;
    int qw = atom_width(d.QC0);
    if (d.nq == 2) { const int w1 = atom_width(d.QC1); if (w1 < qw) qw = w1; }
    if (d.nq == 2 && d.QC0 % qw != 0) return pl;
    const int taps_all = d.tapD * d.tapH * d.tapW;
    pl.swap = (qct == 32 || qct == 64) && d.PC % 32 == 0 && d.PC <= 256 && taps_all > 1;
    if (pl.swap) {
        // tap stacking: M = tpi taps x qct Q channels, N = PC
        pl.tpi = 128 / qct;
        pl.tapGroups = (taps_all + pl.tpi - 1) / pl.tpi;
        pl.aw = qw;                 // A atoms come from Q
        pl.aAtoms = 128 / qw;
        pl.bw = atom_width(d.PC);   // B atoms come from P
        pl.bn = d.PC;
        pl.aTiles = 1;
        pl.bTiles = 1;
    } else {
        pl.aw = atom_width(d.PC);
        pl.bw = qw;
        pl.aAtoms = 128 / pl.aw;
        pl.aTiles = (d.PC + 127) / 128;
        pl.bn = qct < 256 ? qct : 256;
        if (qct > 256 && qct % 256 != 0) pl.bn = qct % 128 == 0 ? 128 : qct % 64 == 0 ? 64 : 32;
        if (pl.bn / pl.bw > 8) pl.bn = 8 * pl.bw;   // the producer keeps an 8-entry atom table
        if (pl.bn % pl.bw != 0 || pl.bn % 32 != 0) return pl;
        pl.bTiles = (qct + pl.bn - 1) / pl.bn;
    }
    // narrow tiles are issue-latency bound: give every mbarrier round trip 128 voxels (8 MMAs) instead of 64
    const long long gvox = (long long)d.NB * d.GD * d.GH * d.GW;
    pl.kbox = (pl.swap && gvox >= (1 << 16)) ? 128 : 64;
    long long best = -1;
    for (int cw = 1; cw <= pl.kbox; cw <<= 1)
        for (int ch = 1; cw * ch <= pl.kbox; ch <<= 1)
            for (int cd = 1; cw * ch * cd <= pl.kbox; cd <<= 1) {
                const int cn = pl.kbox / (cw * ch * cd);
                if ((cw - 1) * d.istrW + 1 > 256 || (ch - 1) * d.istrH + 1 > 256 || (cd - 1) * d.istrD + 1 > 256) continue;
                const long long t = (long long)((d.GW + cw - 1) / cw) * ((d.GH + ch - 1) / ch) * ((d.GD + cd - 1) / cd) *
                                    ((d.NB + cn - 1) / cn);
                if (best < 0 || t < best || (t == best && cw > pl.cw)) {
                    best = t;
                    pl.cw = cw; pl.ch = ch; pl.cd = cd; pl.cn = cn;
                }
            }
    if (best < 0 || best > 2000000000LL) return pl;
    pl.chunksW = (d.GW + pl.cw - 1) / pl.cw;
    pl.chunksH = (d.GH + pl.ch - 1) / pl.ch;
    pl.chunksD = (d.GD + pl.cd - 1) / pl.cd;
    pl.chunksN = (d.NB + pl.cn - 1) / pl.cn;
    const int taps = pl.swap ? pl.tapGroups : taps_all;
    const long long base_items = (long long)pl.aTiles * pl.bTiles * taps;
    long long splits = d.splits > 0 ? d.splits : (3LL * num_sms() + base_items - 1) / base_items;
    // a wave or more of (tile, tap) items already (the 512-channel layers: 216 / 432 items): no voxel split, so that every
    // element has one writer - plain stores, no zero-fill, no red.global (enc_s4 wgrad: 60 us at 18 % of peak with 3 splits)
    static const bool no_single = getenv("RESENC_WGRAD_ALWAYS_SPLIT") != nullptr;
    if (d.splits <= 0 && !no_single && !pl.swap && base_items >= num_sms()) splits = 1;
    long long maxs = (best + 3) / 4;      // at least 4 chunks (256 voxels) per item
    if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    pl.chunksPerSplit = (int)((best + splits - 1) / splits);
    pl.splits = (int)((best + pl.chunksPerSplit - 1) / pl.chunksPerSplit);
    const size_t stageBytes = (size_t)pl.kbox * 2 * (128 + pl.bn);
    int st = (int)((200 * 1024) / stageBytes);
    if (st > 8) st = 8;
    if (st < 2) return pl;
    pl.stages = st;
    pl.accBufs = 2 * pl.bn <= 512 ? 2 : 1;
    pl.smem = 1024 + 1024 + (size_t)st * stageBytes;
    pl.ok = true;
    return pl;
}

// ---- two-sided tap stacking (wgrad2_tc5.cuh): stride-1 convs, channel counts multiples of 32 ----------------------
struct Tw52Plan {
    bool ok = false;
    int pw = 0, qw = 0, mAtomsTotal = 0, nAtomsTotal = 0, mPerGroup = 0, nPerGroup = 0, mGroups = 0, nGroups = 0;
    int kbox = 64, cw = 0, ch = 0, cd = 0, cn = 0, chunksW = 0, chunksH = 0, chunksD = 0, chunksN = 0;
    int splits = 1, chunksPerSplit = 0, stages = 0;
    size_t smem = 0;
    bool merged = false;
};

Tw52Plan plan_tw52(const RbWgradDesc& d) {
    Tw52Plan pl;
    static const bool off = getenv("RESENC_NO_WGRAD2") != nullptr;
    if (off) return pl;
    if (d.istrD != 1 || d.istrH != 1 || d.istrW != 1) return pl;
    if (d.GD != d.QD || d.GH != d.QH || d.GW != d.QW) return pl;
    const int taps = d.tapD * d.tapH * d.tapW;
    if (taps < 9 || d.tapW > 3) return pl;
    const int qct = d.QC0 + (d.nq == 2 ? d.QC1 : 0);
    if (d.PC % 32 != 0 || d.QC0 % 32 != 0 || (d.nq == 2 && d.QC1 % 32 != 0)) return pl;
    pl.pw = d.PC % 64 == 0 ? 64 : 32;
    pl.qw = d.QC0 % 64 == 0 ? 64 : 32;
    if (d.nq == 2 && d.QC1 % 64 != 0) pl.qw = 32;
    pl.mPerGroup = 128 / pl.qw;
    pl.nPerGroup = 256 / pl.pw;
    pl.mAtomsTotal = d.tapW * (qct / pl.qw);
    pl.nAtomsTotal = d.tapD * d.tapH * (d.PC / pl.pw);
    pl.mGroups = (pl.mAtomsTotal + pl.mPerGroup - 1) / pl.mPerGroup;
    pl.nGroups = (pl.nAtomsTotal + pl.nPerGroup - 1) / pl.nPerGroup;
    // 32-channel P: all nine (kd,kh) atoms (288 columns) fit one TMEM accumulator, so one item issues both the 256- and
    // the 32-column MMA on the same Q boxes instead of two items that each load them (the kernel is TMA-row-rate bound:
    // 960 box rows per 64 voxels against 768 merged)
    static const bool no_merge = getenv("RESENC_NO_WGRAD2_MERGE") != nullptr;
    if (!no_merge && pl.nGroups > 1 && pl.nAtomsTotal * pl.pw <= 512 && pl.nAtomsTotal <= 16) {
        pl.merged = true;
        pl.nPerGroup = pl.nAtomsTotal;
        pl.nGroups = 1;
    }
    long long best = -1;
    for (int cw = 1; cw <= pl.kbox; cw <<= 1)
        for (int ch = 1; cw * ch <= pl.kbox; ch <<= 1)
            for (int cd = 1; cw * ch * cd <= pl.kbox; cd <<= 1) {
                const int cn = pl.kbox / (cw * ch * cd);
                const long long t = (long long)((d.GW + cw - 1) / cw) * ((d.GH + ch - 1) / ch) * ((d.GD + cd - 1) / cd) *
                                    ((d.NB + cn - 1) / cn);
                if (best < 0 || t < best || (t == best && cw > pl.cw)) {
                    best = t;
                    pl.cw = cw; pl.ch = ch; pl.cd = cd; pl.cn = cn;
                }
            }
    if (best < 0 || best > 2000000000LL) return pl;
    pl.chunksW = (d.GW + pl.cw - 1) / pl.cw; pl.chunksH = (d.GH + pl.ch - 1) / pl.ch;
    pl.chunksD = (d.GD + pl.cd - 1) / pl.cd; pl.chunksN = (d.NB + pl.cn - 1) / pl.cn;
    const long long base_items = (long long)pl.mGroups * pl.nGroups;
    long long splits = d.splits > 0 ? d.splits : (3LL * num_sms() + base_items - 1) / base_items;
    long long maxs = (best + 7) / 8;      // at least 8 chunks (512 voxels) per item
    if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    pl.chunksPerSplit = (int)((best + splits - 1) / splits);
    pl.splits = (int)((best + pl.chunksPerSplit - 1) / pl.chunksPerSplit);
    const size_t stageBytes = (size_t)pl.kbox * 2 * (128 + (pl.merged ? pl.nAtomsTotal * pl.pw : 256));
    int st = (int)((200 * 1024) / stageBytes);
    if (st > 8) st = 8;
    if (st < 2) return pl;
    pl.stages = st;
    pl.smem = 1024 + 1024 + (size_t)st * stageBytes;
    pl.ok = true;
    return pl;
}

bool prefers_tw52(const RbWgradDesc& d, const Tw52Plan& pl) {
    if (!pl.ok) return false;
    // enough voxels to amortise the 288-column epilogue of every item
    return (long long)d.NB * d.GD * d.GH * d.GW >= 4096;
}

CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
    return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int launch_tw5(const RbWgradDesc& d, const Tw5Plan& pl, const void* P, const void* Q0, const void* Q1, float* dw, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    rb::Tc5WgradParams p;
    memset(&p, 0, sizeof(p));
    {
        const cuuint64_t C = (cuuint64_t)d.PC;
        cuuint64_t dims[5] = {C, (cuuint64_t)d.GW, (cuuint64_t)d.GH, (cuuint64_t)d.GD, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {C * 2, C * 2 * d.GW, C * 2 * d.GW * d.GH, C * 2 * d.GW * d.GH * d.GD};
        const int pw = pl.swap ? pl.bw : pl.aw;
        cuuint32_t box[5] = {(cuuint32_t)pw, (cuuint32_t)pl.cw, (cuuint32_t)pl.ch, (cuuint32_t)pl.cd, (cuuint32_t)pl.cn};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&p.mapP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(P), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(pw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(P) failed: %d", (int)r);
    }
    const void* qs[2] = {Q0, Q1};
    const int qc[2] = {d.QC0, d.QC1};
    for (int s = 0; s < d.nq; ++s) {
        const cuuint64_t C = (cuuint64_t)qc[s];
        cuuint64_t dims[5] = {C, (cuuint64_t)d.QW, (cuuint64_t)d.QH, (cuuint64_t)d.QD, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {C * 2, C * 2 * d.QW, C * 2 * d.QW * d.QH, C * 2 * d.QW * d.QH * d.QD};
        const int qw = pl.swap ? pl.aw : pl.bw;
        cuuint32_t box[5] = {(cuuint32_t)qw, (cuuint32_t)((pl.cw - 1) * d.istrW + 1), (cuuint32_t)((pl.ch - 1) * d.istrH + 1),
                             (cuuint32_t)((pl.cd - 1) * d.istrD + 1), (cuuint32_t)pl.cn};
        cuuint32_t estr[5] = {1, (cuuint32_t)d.istrW, (cuuint32_t)d.istrH, (cuuint32_t)d.istrD, 1};
        CUresult r = enc(&p.mapQ[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(qs[s]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(qw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(Q%d) failed: %d", s, (int)r);
    }
    p.PC = d.PC; p.QC[0] = d.QC0; p.QC[1] = d.nq == 2 ? d.QC1 : 0; p.nq = d.nq;
    p.aw = pl.aw; p.bw = pl.bw; p.aAtoms = pl.aAtoms; p.bn = pl.bn; p.aTiles = pl.aTiles; p.bTiles = pl.bTiles;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW; p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.istrD = d.istrD; p.istrH = d.istrH; p.istrW = d.istrW;
    p.cw = pl.cw; p.ch = pl.ch; p.cd = pl.cd; p.cn = pl.cn;
    p.chunksW = pl.chunksW; p.chunksH = pl.chunksH; p.chunksD = pl.chunksD; p.chunksN = pl.chunksN;
    p.splits = pl.splits; p.chunksPerSplit = pl.chunksPerSplit; p.stages = pl.stages; p.accBufs = pl.accBufs;
    p.dw = dw;
    p.swap = pl.swap ? 1 : 0; p.tpi = pl.tpi; p.tapGroups = pl.tapGroups;
    p.kbox = pl.kbox;
    p.fdChunksW = rb::make_fastdiv(pl.chunksW); p.fdChunksH = rb::make_fastdiv(pl.chunksH); p.fdChunksD = rb::make_fastdiv(pl.chunksD);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(rb::tc5_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(tw5): %s", cudaGetErrorString(attr_err));
    const long long items = (long long)pl.aTiles * pl.bTiles * (pl.swap ? pl.tapGroups : d.tapD * d.tapH * d.tapW) * pl.splits;
    const long long grid = items < num_sms() ? items : num_sms();
    rb::tc5_wgrad_kernel<<<(int)grid, rb::TW5_THREADS, pl.smem, st>>>(p);
    return check_launch("tc5_wgrad_kernel");
}

int launch_tw52(const RbWgradDesc& d, const Tw52Plan& pl, const void* P, const void* Q0, const void* Q1, float* dw, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    rb::Tc5Wgrad2Params p;
    memset(&p, 0, sizeof(p));
    auto encode = [&](CUtensorMap* m, const void* ptr, int C, int W, int H, int D, int width) -> CUresult {
        const cuuint64_t c = (cuuint64_t)C;
        cuuint64_t dims[5] = {c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)d.NB};
        cuuint64_t strides[4] = {c * 2, c * 2 * W, c * 2 * W * H, c * 2 * W * H * D};
        cuuint32_t box[5] = {(cuuint32_t)width, (cuuint32_t)pl.cw, (cuuint32_t)pl.ch, (cuuint32_t)pl.cd, (cuuint32_t)pl.cn};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(width * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = encode(&p.mapP, P, d.PC, d.GW, d.GH, d.GD, pl.pw);
    if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(P) failed: %d", (int)r);
    r = encode(&p.mapQ[0], Q0, d.QC0, d.QW, d.QH, d.QD, pl.qw);
    if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(Q0) failed: %d", (int)r);
    if (d.nq == 2) {
        r = encode(&p.mapQ[1], Q1, d.QC1, d.QW, d.QH, d.QD, pl.qw);
        if (r != CUDA_SUCCESS) return fail(RB_ERR_CUDA, "cuTensorMapEncodeTiled(Q1) failed: %d", (int)r);
    }
    p.PC = d.PC; p.QC[0] = d.QC0; p.QC[1] = d.nq == 2 ? d.QC1 : 0; p.nq = d.nq;
    p.pw = pl.pw; p.qw = pl.qw; p.mAtomsTotal = pl.mAtomsTotal; p.nAtomsTotal = pl.nAtomsTotal;
    p.mPerGroup = pl.mPerGroup; p.nPerGroup = pl.nPerGroup; p.mGroups = pl.mGroups; p.nGroups = pl.nGroups;
    p.merged = pl.merged ? 1 : 0;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW; p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.kbox = pl.kbox; p.cw = pl.cw; p.ch = pl.ch; p.cd = pl.cd; p.cn = pl.cn;
    p.chunksW = pl.chunksW; p.chunksH = pl.chunksH; p.chunksD = pl.chunksD; p.chunksN = pl.chunksN;
    p.fdChunksW = rb::make_fastdiv(pl.chunksW); p.fdChunksH = rb::make_fastdiv(pl.chunksH); p.fdChunksD = rb::make_fastdiv(pl.chunksD);
    p.fdMGroups = rb::make_fastdiv(pl.mGroups); p.fdNGroups = rb::make_fastdiv(pl.nGroups);
    p.splits = pl.splits; p.chunksPerSplit = pl.chunksPerSplit; p.stages = pl.stages;
    p.dw = dw;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(rb::tc5_wgrad2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(rb::tc5_wgrad2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return fail(RB_ERR_CUDA, "cudaFuncSetAttribute(tw52): %s", cudaGetErrorString(attr_err));
    const long long items = (long long)pl.mGroups * pl.nGroups * pl.splits;
    const long long grid = items < num_sms() ? items : num_sms();
    if (pl.merged) rb::tc5_wgrad2_kernel<true><<<(int)grid, rb::TW52_THREADS, pl.smem, st>>>(p);
    else rb::tc5_wgrad2_kernel<false><<<(int)grid, rb::TW52_THREADS, pl.smem, st>>>(p);
    return check_launch("tc5_wgrad2_kernel");
}

}  // namespace

extern "C" {

const char* rb_last_error(void) { return g_err; }
int rb_version(void) { return 100; }
long long rb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int rb_device_error(void* stream) {
    RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    int h = 0;
    RB_CUDA(cudaMemcpyFromSymbol(&h, rb::g_dev_error, sizeof(int)));
    if (h != 0) {
        int z = 0;
        RB_CUDA(cudaMemcpyToSymbol(rb::g_dev_error, &z, sizeof(int)));
        return fail(RB_ERR_DEVICE, "device pipeline timeout, code %d", h);
    }
    return RB_OK;
}

int rb_debug_counters(unsigned long long* out16) {
    if (!out16) return fail(RB_ERR_INVALID, "debug_counters: null pointer");
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpyFromSymbol(out16, rb::g_dbg, 16 * sizeof(unsigned long long)));
    return RB_OK;
}

int rb_conv_gather_tc5_supported(const RbConvDesc* d) {
    if (!d) return 0;
    return plan_tc5(*d).ok ? 1 : 0;
}

enum ConvChoice { CH_UNSUPPORTED = 0, CH_MMA = 1, CH_TC5 = 2, CH_TC5T = 3, CH_TC5T_SPLIT = 4, CH_SLAB = 5 };
struct ConvDecision {
    int choice = CH_MMA;
    Tc5Plan pl;
    Tc5tPlan plt;
    SlabPlan pls;
};

// One place decides which kernel runs a descriptor (used by the plan / workspace queries and by the launch).
ConvDecision decide_conv(const RbConvDesc& d, bool stats_requested) {
    ConvDecision r;
    if (d.impl == RB_IMPL_MMA_SYNC) return r;
    r.pls = plan_slab(d);
    if (r.pls.ok) { r.choice = CH_SLAB; return r; }
    r.pl = plan_tc5(d);
    r.plt = plan_tc5t(d);
    static const bool no_t = getenv("RESENC_NO_TC5T") != nullptr;
    const bool split_ok = r.plt.ok && r.plt.splitK > 1 && !no_t;   // split_finish_kernel takes the statistics
    (void)stats_requested;
    if (split_ok) { r.choice = CH_TC5T_SPLIT; return r; }
    const bool tc5 = d.impl == RB_IMPL_TCGEN05 ? r.pl.ok : auto_prefers_tc5(d, r.pl);
    if (tc5) {
        r.choice = auto_prefers_tc5t(d, r.plt) ? CH_TC5T : CH_TC5;
        return r;
    }
    if (d.impl == RB_IMPL_TCGEN05) r.choice = CH_UNSUPPORTED;
    return r;
}

int rb_conv_gather_plan(const RbConvDesc* d) {
    if (!d) return RB_ERR_INVALID;
    const ConvDecision r = decide_conv(*d, false);
    switch (r.choice) {
        case CH_MMA: return RB_IMPL_MMA_SYNC;
        case CH_TC5: case CH_TC5T: return RB_IMPL_TCGEN05;
        case CH_SLAB: return RB_IMPL_TCGEN05_SLAB;
        case CH_TC5T_SPLIT: return RB_IMPL_TCGEN05_SPLITK;
        default: return RB_ERR_UNSUPPORTED;
    }
}

size_t rb_conv_gather_workspace(const RbConvDesc* d) {
    if (!d) return 0;
    const ConvDecision r = decide_conv(*d, false);
    const size_t full = (size_t)d->NB * d->OD * d->OH * d->OW * d->Nout * sizeof(float);
    if (r.choice == CH_TC5T_SPLIT)      // one slice per tap split, in tile-local order (tiles padded to 256 voxels)
        return (size_t)(r.plt.tiles / r.plt.tilesM) * 256 * d->Nout * sizeof(float) * (size_t)r.plt.splitK;
    if (r.choice == CH_MMA && generic_splitk(*d) > 1) return full;
    return 0;
}

static void fill_generic_params(const RbConvDesc& d, const void* src0, const void* src1, const void* w, void* out0, void* out1,
                                rb::GConvParams& p) {
    memset(&p, 0, sizeof(p));
    p.src[0] = (const rb::bf16*)src0; p.src[1] = (const rb::bf16*)src1;
    p.srcC[0] = d.srcC0; p.srcC[1] = d.nsrc == 2 ? d.srcC1 : 0; p.nsrc = d.nsrc;
    p.ID = d.ID; p.IH = d.IH; p.IW = d.IW; p.NB = d.NB;
    p.w = (const rb::bf16*)w;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW; p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.istrD = d.istrD; p.istrH = d.istrH; p.istrW = d.istrW;
    p.OD = d.OD; p.OH = d.OH; p.OW = d.OW; p.Nout = d.Nout;
    p.mode = d.mode; p.ostrD = d.ostrD; p.ostrH = d.ostrH; p.ostrW = d.ostrW;
    p.ooffD = d.ooffD; p.ooffH = d.ooffH; p.ooffW = d.ooffW; p.FD = d.FD; p.FH = d.FH; p.FW = d.FW;
    p.out0 = out0; p.out1 = out1; p.outC0 = d.outC0; p.outC1 = d.outC1;
    p.psC = d.psC; p.psD = d.psD; p.psH = d.psH; p.psW = d.psW;
    p.outF32 = d.outF32;
}

int rb_conv_gather(const RbConvDesc* dp, const void* src0, const void* src1, const void* w, void* out0, void* out1,
                   float* stat_sum, float* stat_sq, void* workspace, size_t workspace_bytes, void* stream) {
    if (!dp) return fail(RB_ERR_INVALID, "conv: null descriptor");
    const RbConvDesc& d = *dp;
    int rc = validate_conv(d);
    if (rc) return rc;
    if (!src0 || !w || !out0 || (d.nsrc == 2 && !src1) || (d.outC1 > 0 && !out1)) return fail(RB_ERR_INVALID, "conv: null pointer");
    if (!aligned16(src0) || !aligned16(src1) || !aligned16(w) || !aligned16(out0) || !aligned16(out1))
        return fail(RB_ERR_INVALID, "conv: pointers must be 16-byte aligned");
    if ((stat_sum == nullptr) != (stat_sq == nullptr)) return fail(RB_ERR_INVALID, "conv: stat_sum and stat_sq go together");
    if (d.impl != RB_IMPL_AUTO && d.impl != RB_IMPL_MMA_SYNC && d.impl != RB_IMPL_TCGEN05) return fail(RB_ERR_INVALID, "conv: unknown impl %d", d.impl);
    cudaStream_t st = (cudaStream_t)stream;
    const long long M = (long long)d.NB * d.OD * d.OH * d.OW;
    const size_t need = (size_t)M * d.Nout * sizeof(float);

    ConvDecision dec = decide_conv(d, stat_sum != nullptr);
    const size_t need_split = dec.choice == CH_TC5T_SPLIT
        ? (size_t)(dec.plt.tiles / dec.plt.tilesM) * 256 * d.Nout * sizeof(float) * (size_t)dec.plt.splitK : 0;
    if (dec.choice == CH_TC5T_SPLIT && (workspace == nullptr || workspace_bytes < need_split)) {
        // no room for the per-slice workspace: run the same kernel unsplit
        dec.plt.splitK = 1; dec.plt.tapsPer = d.tapD * d.tapH * d.tapW;
        dec.choice = CH_TC5T;
    }
    if (dec.choice == CH_UNSUPPORTED) return fail(RB_ERR_UNSUPPORTED, "conv: shape does not qualify for the tcgen05 kernel");
    if (dec.choice == CH_SLAB) return launch_slab(d, dec.pls, src0, src1, w, out0, out1, stat_sum, stat_sq, st);
    if (dec.choice == CH_TC5) return launch_tc5(d, dec.pl, src0, src1, w, out0, out1, stat_sum, stat_sq, st);
    if (dec.choice == CH_TC5T) {
        Tc5tPlan plt = dec.plt;
        plt.splitK = 1; plt.tapsPer = d.tapD * d.tapH * d.tapW;
        return launch_tc5t(d, plt, src0, src1, w, out0, out1, stat_sum, stat_sq, nullptr, st);
    }
    if (dec.choice == CH_TC5T_SPLIT) {
        rc = launch_tc5t(d, dec.plt, src0, src1, w, out0, out1, nullptr, nullptr, (float*)workspace, st);
        if (rc) return rc;
        rb::SplitFinishParams f;
        memset(&f, 0, sizeof(f));
        f.ws = (const float*)workspace; f.sliceStride = (dec.plt.tiles / dec.plt.tilesM) * 256LL * d.Nout; f.slices = dec.plt.splitK;
        f.S = d.OD * d.OH * d.OW; f.NB = d.NB; f.Nout = d.Nout;
        f.OW = d.OW; f.OH = d.OH; f.OD = d.OD; f.lw = dec.plt.lw; f.lh = dec.plt.lh; f.ld = dec.plt.ld;
        f.tilesW = dec.plt.tilesW; f.tilesH = dec.plt.tilesH; f.tilesD = dec.plt.tilesD;
        f.out0 = out0; f.out1 = out1; f.outC0 = d.outC0; f.outC1 = d.outC1; f.outF32 = d.outF32;
        f.stat_sum = stat_sum; f.stat_sq = stat_sq;
        f.vpw = f.S <= 1024 ? 4 : 16;   // a warp moves 4 voxels x 32 channels per instruction; the pass is latency bound
        const int nw = f.S <= 256 ? 2 : 8;       // warps per block
        f.runsPerSample = (f.S + nw * f.vpw - 1) / (nw * f.vpw);
        const long long blocks = (long long)d.NB * f.runsPerSample * (d.Nout / 32);
        rb::split_finish_kernel<<<(unsigned)blocks, 32 * nw, 0, st>>>(f);
        return check_launch("split_finish_kernel");
    }
    if (stat_sum) return fail(RB_ERR_UNSUPPORTED, "conv: fused statistics need the tcgen05 path");

    rb::GConvParams p;
    fill_generic_params(d, src0, src1, w, out0, out1, p);
    int sk = generic_splitk(d);
    if (sk > 1 && (workspace == nullptr || workspace_bytes < need)) sk = 1;
    p.splitK = sk;
    p.ws = sk > 1 ? (float*)workspace : nullptr;
    const long long gx = (M + rb::GC_BM - 1) / rb::GC_BM;
    if (gx > 2147483647LL) return fail(RB_ERR_INVALID, "conv: grid too large");
    dim3 grid((unsigned)gx, (unsigned)((d.Nout + rb::GC_BN - 1) / rb::GC_BN), (unsigned)sk);
    if (sk > 1) RB_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    rb::gather_conv_mma_kernel<<<grid, rb::GC_THREADS, 0, st>>>(p);
    rc = check_launch("gather_conv_mma_kernel");
    if (rc) return rc;
    if (sk > 1) {
        rb::gather_finish_kernel<<<grid_for(M * (d.Nout / 2), 256), 256, 0, st>>>(p);
        rc = check_launch("gather_finish_kernel");
    }
    return rc;
}

int rb_wgrad_gather(const RbWgradDesc* dp, const void* P, const void* Q0, const void* Q1, float* dw, void* stream) {
    if (!dp) return fail(RB_ERR_INVALID, "wgrad: null descriptor");
    const RbWgradDesc& d = *dp;
    if (d.PC <= 0 || d.PC % 8 != 0 || d.nq < 1 || d.nq > 2 || d.QC0 <= 0 || d.QC0 % 8 != 0 || (d.nq == 2 && (d.QC1 <= 0 || d.QC1 % 8 != 0)))
        return fail(RB_ERR_INVALID, "wgrad: channel counts must be positive multiples of 8");
    if (!P || !Q0 || !dw || (d.nq == 2 && !Q1)) return fail(RB_ERR_INVALID, "wgrad: null pointer");
    if (!aligned16(P) || !aligned16(Q0) || !aligned16(Q1)) return fail(RB_ERR_INVALID, "wgrad: pointers must be 16-byte aligned");
    if (d.NB <= 0 || d.GD <= 0 || d.GH <= 0 || d.GW <= 0 || d.QD <= 0 || d.QH <= 0 || d.QW <= 0) return fail(RB_ERR_INVALID, "wgrad: empty grid");
    const size_t dw_bytes = (size_t)d.tapD * d.tapH * d.tapW * d.PC * (d.QC0 + (d.nq == 2 ? d.QC1 : 0)) * sizeof(float);
    if (d.impl != RB_IMPL_MMA_SYNC) {
        Tw52Plan pl2 = plan_tw52(d);
        if (prefers_tw52(d, pl2)) {
            RB_CUDA(cudaMemsetAsync(dw, 0, dw_bytes, (cudaStream_t)stream));
            return launch_tw52(d, pl2, P, Q0, Q1, dw, (cudaStream_t)stream);
        }
        Tw5Plan pl = plan_tw5(d);
        if (pl.ok) {
            // every element has exactly one writer when nothing is split: no zero-fill, plain stores
            if (pl.swap || pl.splits > 1) RB_CUDA(cudaMemsetAsync(dw, 0, dw_bytes, (cudaStream_t)stream));
            return launch_tw5(d, pl, P, Q0, Q1, dw, (cudaStream_t)stream);
        }
        if (d.impl == RB_IMPL_TCGEN05) return fail(RB_ERR_UNSUPPORTED, "wgrad: shape does not qualify for the tcgen05 kernel");
    }
    rb::GWgradParams p;
    memset(&p, 0, sizeof(p));
    p.P = (const rb::bf16*)P; p.PC = d.PC;
    p.Q[0] = (const rb::bf16*)Q0; p.Q[1] = (const rb::bf16*)Q1; p.QC[0] = d.QC0; p.QC[1] = d.nq == 2 ? d.QC1 : 0; p.nq = d.nq;
    p.NB = d.NB; p.GD = d.GD; p.GH = d.GH; p.GW = d.GW; p.QD = d.QD; p.QH = d.QH; p.QW = d.QW;
    p.tapD = d.tapD; p.tapH = d.tapH; p.tapW = d.tapW; p.offD = d.offD; p.offH = d.offH; p.offW = d.offW;
    p.istrD = d.istrD; p.istrH = d.istrH; p.istrW = d.istrW;
    p.dw = dw;
    const int qct = p.QC[0] + p.QC[1];
    const int tiles = ((d.PC + rb::GW_BA - 1) / rb::GW_BA) * ((qct + rb::GW_BB - 1) / rb::GW_BB);
    const int taps = d.tapD * d.tapH * d.tapW;
    const long long M = (long long)d.NB * d.GD * d.GH * d.GW;
    long long splits = d.splits;
    if (splits <= 0) {
        splits = (4LL * num_sms() + (long long)tiles * taps - 1) / ((long long)tiles * taps);
        const long long maxs = (M + 255) / 256;  // at least 256 voxels per slice
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    if (splits > 65535) splits = 65535;
    long long per = (M + splits - 1) / splits;
    per = (per + rb::GW_BK - 1) / rb::GW_BK * rb::GW_BK;
    splits = (M + per - 1) / per;
    p.mPerSplit = (int)per;
    dim3 grid((unsigned)tiles, (unsigned)taps, (unsigned)splits);
    RB_CUDA(cudaMemsetAsync(dw, 0, dw_bytes, (cudaStream_t)stream));
    rb::gather_wgrad_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("gather_wgrad_mma_kernel");
}

int rb_plane_reduce(int kind, const void* y, int y_f32, const void* dz, const void* z, const float* sign_scale,
                    const float* sign_shift, double* out, int NB, long long S, int C, int W, int perW, float slope, void* stream) {
    if (!y || !out || (kind == 1 && !dz)) return fail(RB_ERR_INVALID, "plane_reduce: null pointer");
    if ((sign_scale == nullptr) != (sign_shift == nullptr) || (sign_scale && (z || perW || kind != 1)))
        return fail(RB_ERR_INVALID, "plane_reduce: sign_scale/sign_shift go together, kind 1 only, without z and per-w");
    if (C <= 0 || C % 8 != 0 || NB <= 0 || S <= 0) return fail(RB_ERR_INVALID, "plane_reduce: bad shape");
    if (perW && (W <= 0 || S % W != 0)) return fail(RB_ERR_INVALID, "plane_reduce: S must be a multiple of W");
    if (kind != 0 && kind != 1) return fail(RB_ERR_INVALID, "plane_reduce: kind must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    rb::ReduceParams p;
    p.y = y; p.dz = (const rb::bf16*)dz; p.z = (const rb::bf16*)z; p.out = out; p.sgnA = sign_scale; p.sgnB = sign_shift;
    p.S = S; p.C = C; p.W = W; p.perW = perW; p.slope = slope; p.kind = kind; p.yF32 = y_f32;
    const int cg = C / 8;
    const int rows = 256 / cg > 0 ? 256 / cg : 1;
    int gx;
    if (perW) {
        gx = W;
    } else {
        RB_CUDA(cudaMemsetAsync(out, 0, (size_t)NB * C * 2 * sizeof(double), st));
        long long want = (4LL * num_sms() + NB - 1) / NB;
        const long long maxb = (S + rows - 1) / rows;
        if (want > maxb) want = maxb;
        gx = (int)(want < 1 ? 1 : want);
    }
    const size_t smem = (size_t)256 * 16 * sizeof(float);
    // two voxels in flight per thread (RESENC_NORM_U2=0 restores the one-voxel kernels)
    static const bool u2 = !(getenv("RESENC_NORM_U2") && atoi(getenv("RESENC_NORM_U2")) == 0);
    if (u2 && !perW && y_f32 != 1 && cg <= 256 && 256 % cg == 0 && S * cg < (1LL << 31)) {
        if (sign_scale) rb::plane_reduce_u2_kernel<true><<<dim3(gx, NB), 256, smem, st>>>(p);
        else rb::plane_reduce_u2_kernel<false><<<dim3(gx, NB), 256, smem, st>>>(p);
        return check_launch("plane_reduce_u2_kernel");
    }
    if (sign_scale) rb::plane_reduce_kernel<true><<<dim3(gx, NB), 256, smem, st>>>(p);
    else rb::plane_reduce_kernel<false><<<dim3(gx, NB), 256, smem, st>>>(p);
    return check_launch("plane_reduce_kernel");
}

int rb_in_finalize_fwd(const double* sums, const float* fsum, const float* fsq, const float* gamma, const float* beta,
                       float* mean, float* rstd, float* scale, float* shift, int NB, int C, double S, double eps, void* stream) {
    if ((!sums && (!fsum || !fsq)) || !mean || !rstd || !scale || !shift) return fail(RB_ERR_INVALID, "in_finalize_fwd: null pointer");
    rb::FinalizeParams p{sums, fsum, fsq, gamma, beta, mean, rstd, scale, shift, NB, C, S, eps};
    rb::in_finalize_fwd_kernel<<<(NB * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("in_finalize_fwd_kernel");
}

int rb_in_finalize_bwd(const double* red, const float* mean, const float* rstd, const float* gamma, float* k1, float* k2,
                       float* k3, float* dgamma, float* dbeta, int NB, int C, double S, void* stream) {
    if (!red || !mean || !rstd || !k1 || !k2 || !k3) return fail(RB_ERR_INVALID, "in_finalize_bwd: null pointer");
    rb::FinalizeBwdParams p{red, mean, rstd, gamma, k1, k2, k3, dgamma, dbeta, NB, C, S};
    rb::in_finalize_bwd_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("in_finalize_bwd_kernel");
}

static int check_apply_shape(const char* who, int NB, long long S, int C, int W, int perW) {
    if (C <= 0 || C % 8 != 0 || NB <= 0 || S <= 0) return fail(RB_ERR_INVALID, "%s: bad shape", who);
    if (S * (C / 8) >= (1LL << 31)) return fail(RB_ERR_INVALID, "%s: one sample exceeds 2^31 vectors", who);
    if (perW && (W <= 0 || S % W != 0)) return fail(RB_ERR_INVALID, "%s: S must be a multiple of W", who);
    return RB_OK;
}

int rb_norm_act_fwd(const void* y, int y_f32, const void* res, void* z, const float* scale, const float* shift, int NB,
                    long long S, int C, int W, int perW, int act, float slope, void* stream) {
    if (!y || !z || !scale || !shift) return fail(RB_ERR_INVALID, "norm_act_fwd: null pointer");
    int rc = check_apply_shape("norm_act_fwd", NB, S, C, W, perW);
    if (rc) return rc;
    rb::ApplyParams p{y, (const rb::bf16*)res, (rb::bf16*)z, scale, shift, S, NB, C, W, perW, act, slope, y_f32};
    const long long per = S * (C / 8);
    int gx = grid_for(per, 256, 8);
    static const int variant = getenv("RESENC_NORM_VARIANT") ? atoi(getenv("RESENC_NORM_VARIANT")) : 0;
    const int cg = C / 8;
    // variant 1 (coefficients hoisted, two groups in flight) measured SLOWER (profiles/r2_norm_bench.txt: 113 / 62 registers
    // cost more occupancy than the saved L1 loads give back): kept behind RESENC_NORM_VARIANT=1 for the record
    if (variant >= 1 && !perW && (cg & (cg - 1)) == 0 && cg <= 64) {
        gx = grid_for((per + 1) / 2, 256, 8);
        rb::norm_act_fwd_v1_kernel<2><<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("norm_act_fwd_v1_kernel");
    }
    // two vectors per thread: 121 vs 127 us on the 128^3 launches but 37 vs 35 / 21 vs 15 us on the 64^3 / 32^3 ones
    // (tools/norm_bench.py) - the pass sits at the mixed read / write DRAM limit (~4.4 TB/s at 1:1), not at a
    // bytes-in-flight limit; opt-in only
    static const bool x2 = getenv("RESENC_NORM_X2") != nullptr;
    if (x2 && y_f32 != 1 && C % 16 == 0) {
        gx = grid_for(per / 2, 256, 8);
        rb::norm_act_fwd_x2_kernel<<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("norm_act_fwd_x2_kernel");
    }
    static const bool un = !(getenv("RESENC_NORM_U2") && atoi(getenv("RESENC_NORM_U2")) == 0);
    if (un && y_f32 != 1) {
        const cudaStream_t st = (cudaStream_t)stream;
        if (res) {
            const dim3 grid(grid_for((per + 1) / 2, 256, 8), NB);
            if (y_f32 == 2) rb::norm_act_fwd_un_kernel<true, true, 2><<<grid, 256, 0, st>>>(p);
            else rb::norm_act_fwd_un_kernel<false, true, 2><<<grid, 256, 0, st>>>(p);
        } else {
            const dim3 grid(grid_for((per + 3) / 4, 256, 8), NB);
            if (y_f32 == 2) rb::norm_act_fwd_un_kernel<true, false, 4><<<grid, 256, 0, st>>>(p);
            else rb::norm_act_fwd_un_kernel<false, false, 4><<<grid, 256, 0, st>>>(p);
        }
        return check_launch("norm_act_fwd_un_kernel");
    }
    rb::norm_act_fwd_kernel<<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("norm_act_fwd_kernel");
}

int rb_norm_act_bwd(const void* dz, const void* z, const float* sign_scale, const float* sign_shift, const void* y, int y_f32,
                    void* dy, void* dres, const float* k1, const float* k2, const float* k3, int NB, long long S, int C, int W,
                    int perW, int act, float slope, void* stream) {
    if (!dz || !y || !dy || !k1 || !k2 || !k3) return fail(RB_ERR_INVALID, "norm_act_bwd: null pointer");
    if (act && !z && !(sign_scale && sign_shift)) return fail(RB_ERR_INVALID, "norm_act_bwd: the activation needs z or sign_scale/sign_shift");
    int rc = check_apply_shape("norm_act_bwd", NB, S, C, W, perW);
    if (rc) return rc;
    rb::ApplyBwdParams p{(const rb::bf16*)dz, (const rb::bf16*)z, y, (rb::bf16*)dy, (rb::bf16*)dres,
                         k1, k2, k3, S, NB, C, W, perW, act, slope, y_f32, sign_scale, sign_shift};
    const long long per = S * (C / 8);
    int gx = grid_for(per, 256, 8);
    static const int variant = getenv("RESENC_NORM_VARIANT") ? atoi(getenv("RESENC_NORM_VARIANT")) : 0;
    const int cg = C / 8;
    if (variant >= 1 && !perW && (cg & (cg - 1)) == 0 && cg <= 64) {
        gx = grid_for((per + 1) / 2, 256, 8);
        rb::norm_act_bwd_v1_kernel<2><<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("norm_act_bwd_v1_kernel");
    }
    static const bool u2 = !(getenv("RESENC_NORM_U2") && atoi(getenv("RESENC_NORM_U2")) == 0);
    if (u2 && y_f32 != 1) {
        gx = grid_for((per + 1) / 2, 256, 8);
        const int zm = !act ? 0 : z ? 1 : 2;
        const dim3 grid(gx, NB);
        const cudaStream_t st = (cudaStream_t)stream;
        if (y_f32 == 2) {
            if (zm == 0) rb::norm_act_bwd_u2_kernel<true, 0><<<grid, 256, 0, st>>>(p);
            else if (zm == 1) rb::norm_act_bwd_u2_kernel<true, 1><<<grid, 256, 0, st>>>(p);
            else rb::norm_act_bwd_u2_kernel<true, 2><<<grid, 256, 0, st>>>(p);
        } else {
            if (zm == 0) rb::norm_act_bwd_u2_kernel<false, 0><<<grid, 256, 0, st>>>(p);
            else if (zm == 1) rb::norm_act_bwd_u2_kernel<false, 1><<<grid, 256, 0, st>>>(p);
            else rb::norm_act_bwd_u2_kernel<false, 2><<<grid, 256, 0, st>>>(p);
        }
        return check_launch("norm_act_bwd_u2_kernel");
    }
    rb::norm_act_bwd_kernel<<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("norm_act_bwd_kernel");
}

static int check_pool(int NB, int D, int H, int W, int C, int sd, int sh, int sw) {
    if (NB <= 0 || C <= 0 || C % 8 != 0) return fail(RB_ERR_INVALID, "avgpool: bad shape");
    if (sd < 1 || sh < 1 || sw < 1 || sd > 2 || sh > 2 || sw > 2) return fail(RB_ERR_INVALID, "avgpool: window must be 1 or 2 per axis");
    if (D % sd || H % sh || W % sw) return fail(RB_ERR_INVALID, "avgpool: dims must be divisible by the window");
    return RB_OK;
}

int rb_avgpool_fwd(const void* in, void* out, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream) {
    if (!in || !out) return fail(RB_ERR_INVALID, "avgpool_fwd: null pointer");
    int rc = check_pool(NB, D, H, W, C, sd, sh, sw);
    if (rc) return rc;
    rb::PoolParams p{(const rb::bf16*)in, (rb::bf16*)out, NB, D, H, W, C, sd, sh, sw};
    const long long total = (long long)NB * (D / sd) * (H / sh) * (W / sw) * (C / 8);
    if (total < (1LL << 31) - (1LL << 22)) rb::avgpool_fwd_kernel<unsigned><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    else rb::avgpool_fwd_kernel<long long><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("avgpool_fwd_kernel");
}

int rb_avgpool_bwd(const void* dout, void* din, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream) {
    if (!dout || !din) return fail(RB_ERR_INVALID, "avgpool_bwd: null pointer");
    int rc = check_pool(NB, D, H, W, C, sd, sh, sw);
    if (rc) return rc;
    rb::PoolParams p{(const rb::bf16*)dout, (rb::bf16*)din, NB, D, H, W, C, sd, sh, sw};
    const long long total = (long long)NB * D * H * W * (C / 8);
    if (total < (1LL << 31) - (1LL << 22)) rb::avgpool_bwd_kernel<unsigned><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    else rb::avgpool_bwd_kernel<long long><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("avgpool_bwd_kernel");
}

static int launch_norm_head(const void* y, int y_mode, const void* res, const float* scale, const float* shift, const float* w,
                            const float* b, float* out, int NB, long long S, int C, int K, int act, float slope, int head_act,
                            void* stream) {
    rb::NormHeadParams p{y, (const rb::bf16*)res, scale, shift, w, b, out, S, NB, C, K, act, head_act, slope, y_mode};
    const size_t smem = (size_t)(K * C + K) * sizeof(float);
    const int vpw = 32 / (C / 8);                                   // voxels per warp instruction
    const long long warpIters = (S + vpw - 1) / vpw;                // per sample
    int gx = (int)std::min<long long>((warpIters + 8 * rb::NH_U - 1) / (8 * rb::NH_U), (long long)num_sms() * 6);   // 8 warps x NH_U voxel sets per block pass
    if (gx < 1) gx = 1;
    const dim3 grid(gx, NB);
    const cudaStream_t st = (cudaStream_t)stream;
#define RB_NH_LAUNCH(KM, NORM, YM, RES) rb::norm_act_head_fwd_kernel<KM, NORM, YM, RES><<<grid, 256, smem, st>>>(p)
#define RB_NH_K(NORM, YM, RES) do { if (K == 1) RB_NH_LAUNCH(1, NORM, YM, RES); else if (K <= 4) RB_NH_LAUNCH(4, NORM, YM, RES); \
                                    else RB_NH_LAUNCH(8, NORM, YM, RES); } while (0)
    if (scale == nullptr) RB_NH_K(0, 0, 0);
    else if (y_mode == 2) { if (res) RB_NH_K(1, 2, 1); else RB_NH_K(1, 2, 0); }
    else if (y_mode == 1) { if (res) RB_NH_K(1, 1, 1); else RB_NH_K(1, 1, 0); }
    else { if (res) RB_NH_K(1, 0, 1); else RB_NH_K(1, 0, 0); }
#undef RB_NH_K
#undef RB_NH_LAUNCH
    return check_launch("norm_act_head_fwd_kernel");
}

int rb_head_fwd(const void* x, const float* w, const float* b, float* out, int NB, long long S, int C, int K, int act, void* stream) {
    if (!x || !w || !out) return fail(RB_ERR_INVALID, "head_fwd: null pointer");
    if (K < 1 || K > rb::HEAD_MAXK || C <= 0 || C % 8 != 0 || C > 1024) return fail(RB_ERR_INVALID, "head_fwd: need 1 <= K <= 8, C %% 8 == 0, C <= 1024");
    const int cg = C / 8;
    static const bool legacy = getenv("RESENC_HEAD_LEGACY") != nullptr;
    if (!legacy && (cg & (cg - 1)) == 0 && cg <= 32 && S < (1LL << 31))   // coalesced (voxel, channel-group) lanes
        return launch_norm_head(x, 0, nullptr, nullptr, nullptr, w, b, out, NB, S, C, K, 0, 0.f, act, stream);
    rb::HeadParams p{(const rb::bf16*)x, w, b, out, S, NB, C, K, act};
    const size_t smem = (size_t)(K * C + K) * sizeof(float);
    rb::head_fwd_kernel<<<grid_for((long long)NB * S, 256), 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("head_fwd_kernel");
}

int rb_norm_act_head_fwd(const void* y, int y_mode, const void* res, const float* scale, const float* shift, const float* w,
                         const float* b, float* out, int NB, long long S, int C, int K, int act, float slope, int head_act,
                         void* stream) {
    if (!y || !scale || !shift || !w || !out) return fail(RB_ERR_INVALID, "norm_act_head_fwd: null pointer");
    if (K < 1 || K > rb::HEAD_MAXK) return fail(RB_ERR_INVALID, "norm_act_head_fwd: need 1 <= K <= 8");
    const int cg = C / 8;
    if (C <= 0 || C % 8 != 0 || (cg & (cg - 1)) != 0 || cg > 32)
        return fail(RB_ERR_UNSUPPORTED, "norm_act_head_fwd: C / 8 must be a power of two <= 32 (C = %d)", C);
    if (NB <= 0 || S <= 0 || S >= (1LL << 31)) return fail(RB_ERR_INVALID, "norm_act_head_fwd: bad shape");
    if (y_mode < 0 || y_mode > 2 || head_act < 0 || head_act > 2) return fail(RB_ERR_INVALID, "norm_act_head_fwd: bad mode");
    return launch_norm_head(y, y_mode, res, scale, shift, w, b, out, NB, S, C, K, act, slope, head_act, stream);
}

int rb_head_bwd(const void* x, const float* w, const float* dl, void* dx, float* dw, float* db, int NB, long long S, int C,
                int K, void* stream) {
    if (!x || !w || !dl || !dx || !dw) return fail(RB_ERR_INVALID, "head_bwd: null pointer");
    if (K < 1 || K > rb::HEAD_MAXK || C <= 0 || C % 8 != 0 || C > 1024) return fail(RB_ERR_INVALID, "head_bwd: need 1 <= K <= 8, C %% 8 == 0, C <= 1024");
    rb::HeadBwdParams p{(const rb::bf16*)x, w, dl, (rb::bf16*)dx, dw, db, S, NB, C, K};
    const size_t smem = (size_t)(2 * K * C + K) * sizeof(float);
    const int grid = grid_for((long long)NB * S * (C / 8), 256, 4);
    if (K == 1) rb::head_bwd_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    else if (K == 2) rb::head_bwd_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    else if (K <= 4) rb::head_bwd_kernel<4><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    else rb::head_bwd_kernel<8><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("head_bwd_kernel");
}

int rb_stem_im2col(const float* x, void* col, int NB, int Cin, int D, int H, int W, int kd, int kh, int kw, int Kp, void* stream) {
    if (!x || !col) return fail(RB_ERR_INVALID, "stem_im2col: null pointer");
    if (Kp % 8 != 0 || Kp < kd * kh * kw * Cin) return fail(RB_ERR_INVALID, "stem_im2col: Kp must be a multiple of 8 covering taps*Cin");
    rb::Im2colParams p{x, (rb::bf16*)col, NB, Cin, D, H, W, kd, kh, kw, Kp};
    if (Cin == 1 && kd == 3 && kh == 3 && kw == 3 && Kp == 32) {
        rb::stem_im2col_fixed_kernel<3, 3, 3, 1, 32><<<grid_for((long long)NB * D * H * W, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("stem_im2col_fixed_kernel");
    }
    const long long total = (long long)NB * D * H * W * (Kp / 8);
    rb::stem_im2col_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("stem_im2col_kernel");
}

// ---- split-precision (bf16x3) inference tier: layout producers (split.cuh) ----
int rb_split_apply(const float* y, const void* res, void* z, const float* scale, const float* shift, int NB, long long S,
                   int C, int act, float slope, void* stream) {
    if (!y || !z) return fail(RB_ERR_INVALID, "split_apply: null pointer");
    if ((scale == nullptr) != (shift == nullptr)) return fail(RB_ERR_INVALID, "split_apply: scale and shift go together");
    int rc = check_apply_shape("split_apply", NB, S, C, 0, 0);
    if (rc) return rc;
    if (!aligned16(y) || !aligned16(res) || !aligned16(z)) return fail(RB_ERR_INVALID, "split_apply: pointers must be 16-byte aligned");
    rb::SplitApplyParams p{y, (const rb::bf16*)res, (rb::bf16*)z, scale, shift, S, NB, C, act, slope};
    const long long per = S * (C / 8);
    rb::split_apply_kernel<<<dim3(grid_for(per, 256, 8), NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("split_apply_kernel");
}

int rb_avgpool_split(const void* in, void* out, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream) {
    if (!in || !out) return fail(RB_ERR_INVALID, "avgpool_split: null pointer");
    int rc = check_pool(NB, D, H, W, C, sd, sh, sw);
    if (rc) return rc;
    rb::PoolParams p{(const rb::bf16*)in, (rb::bf16*)out, NB, D, H, W, C, sd, sh, sw};
    const long long total = (long long)NB * (D / sd) * (H / sh) * (W / sw) * (C / 8);
    rb::avgpool_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("avgpool_split_kernel");
}

int rb_stem_im2col_split(const float* x, void* col, int NB, int Cin, int D, int H, int W, int kd, int kh, int kw, int Kp, void* stream) {
    if (!x || !col) return fail(RB_ERR_INVALID, "stem_im2col_split: null pointer");
    if (Kp % 8 != 0 || Kp < kd * kh * kw * Cin) return fail(RB_ERR_INVALID, "stem_im2col_split: Kp must be a multiple of 8 covering taps*Cin");
    rb::Im2colParams p{x, (rb::bf16*)col, NB, Cin, D, H, W, kd, kh, kw, Kp};
    const long long total = (long long)NB * D * H * W * (Kp / 8);
    rb::stem_im2col_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("stem_im2col_split_kernel");
}

int rb_pack_conv_weights(const float* w, void* out_f, void* out_d, int Cout, int Cin, int T, void* stream) {
    if (!w || (!out_f && !out_d)) return fail(RB_ERR_INVALID, "pack_conv_weights: null pointer");
    if (Cout <= 0 || Cin <= 0 || T <= 0 || T > 27) return fail(RB_ERR_INVALID, "pack_conv_weights: need 1 <= taps <= 27");
    rb::WPackParams p{w, (rb::bf16*)out_f, (rb::bf16*)out_d, Cout, Cin, T};
    const size_t smem = (size_t)32 * (32 * T + 2) * sizeof(rb::bf16);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(rb::pack_conv_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); });
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32);
    static const bool no_vec = getenv("RESENC_NO_VEC_PACK") != nullptr;
    if (!no_vec && Cin % 32 == 0 && aligned16(w) && (reinterpret_cast<uintptr_t>(out_f) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out_d) & 3u) == 0) {
        static std::once_flag once2;
        std::call_once(once2, [] { cudaFuncSetAttribute(rb::pack_conv_weights_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); });
        rb::pack_conv_weights_vec_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
        return check_launch("pack_conv_weights_vec_kernel");
    }
    rb::pack_conv_weights_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("pack_conv_weights_kernel");
}

int rb_pack_conv_dgrad_merged(const float* w, void* out, int Cout, int Cin, int K0, int K1, int K2, const int* ntaps,
                              const int* stride, const signed char* kidx, void* stream) {
    if (!w || !out || !ntaps || !stride || !kidx) return fail(RB_ERR_INVALID, "pack_conv_dgrad_merged: null pointer");
    if (Cout <= 0 || Cin <= 0 || (Cout & 1)) return fail(RB_ERR_UNSUPPORTED, "pack_conv_dgrad_merged: Cout must be even");
    rb::DMergePackParams p;
    p.w = w; p.out = (rb::bf16*)out; p.Cout = Cout; p.Cin = Cin; p.K0 = K0; p.K1 = K1; p.K2 = K2;
    p.nt0 = ntaps[0]; p.nt1 = ntaps[1]; p.nt2 = ntaps[2]; p.s0 = stride[0]; p.s1 = stride[1]; p.s2 = stride[2];
    const int K[3] = {K0, K1, K2};
    for (int a = 0; a < 3; ++a) {
        if (ntaps[a] < 1 || ntaps[a] > 4 || stride[a] < 1 || stride[a] > 2 || K[a] < 1 || K[a] > 3)
            return fail(RB_ERR_INVALID, "pack_conv_dgrad_merged: window taps 1..4, stride 1..2, kernel 1..3 per axis");
        for (int r = 0; r < 2; ++r)
            for (int u = 0; u < 4; ++u) {
                const signed char k = kidx[(a * 2 + r) * 4 + u];
                if (k >= K[a]) return fail(RB_ERR_INVALID, "pack_conv_dgrad_merged: kernel index out of range");
                p.kidx[a][r][u] = k;
            }
    }
    const long long total = (long long)p.nt0 * p.nt1 * p.nt2 * p.s0 * p.s1 * p.s2 * Cin * (Cout / 2);
    rb::pack_dgrad_merged_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("pack_dgrad_merged_kernel");
}

int rb_unpack_wgrad(const float* dwp, float* grad, int A, int B, int T, void* stream) {
    if (!dwp || !grad) return fail(RB_ERR_INVALID, "unpack_wgrad: null pointer");
    if (A <= 0 || B <= 0 || T <= 0 || T > 27) return fail(RB_ERR_INVALID, "unpack_wgrad: need 1 <= taps <= 27");
    rb::WUnpackParams p{dwp, grad, A, B, T};
    const size_t smem = (size_t)8 * (32 * T + 1) * sizeof(float);
    dim3 grid((B + 31) / 32, (A + 7) / 8);
    if (T == 27) rb::unpack_wgrad_kernel<27><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    else if (T == 8) rb::unpack_wgrad_kernel<8><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    else rb::unpack_wgrad_kernel<0><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("unpack_wgrad_kernel");
}

// ---- optimiser step (train.py:79-83 AdamW, train.py:227 clip_grad_norm_) ---------------------------------------
static int opt_fill(rb::OptTensorList& L, const rb_opt_tensor* t, int cnt, bool need_state, long long* max_n) {
    *max_n = 0;
    for (int i = 0; i < cnt; ++i) {
        if (!t[i].g || t[i].n <= 0 || (need_state && (!t[i].p || !t[i].m || !t[i].v))) return fail(RB_ERR_INVALID, "optimiser tensor list: null pointer or empty tensor");
        L.p[i] = (float*)t[i].p; L.g[i] = (const float*)t[i].g; L.m[i] = (float*)t[i].m; L.v[i] = (float*)t[i].v; L.n[i] = t[i].n;
        if (t[i].n > *max_n) *max_n = t[i].n;
    }
    for (int i = cnt; i < rb::OPT_MAX_TENSORS; ++i) { L.p[i] = nullptr; L.g[i] = nullptr; L.m[i] = nullptr; L.v[i] = nullptr; L.n[i] = 0; }
    return RB_OK;
}

int rb_grad_sumsq(const rb_opt_tensor* tensors, int count, double* sumsq, void* stream) {
    if (!tensors || count <= 0 || !sumsq) return fail(RB_ERR_INVALID, "grad_sumsq: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    RB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
    for (int i0 = 0; i0 < count; i0 += rb::OPT_MAX_TENSORS) {
        const int cnt = count - i0 < rb::OPT_MAX_TENSORS ? count - i0 : rb::OPT_MAX_TENSORS;
        rb::OptTensorList L;
        long long max_n;
        int rc = opt_fill(L, tensors + i0, cnt, false, &max_n);
        if (rc) return rc;
        long long gx = (max_n / 4 + 256 * 4 - 1) / (256 * 4);      // four vectors per thread and pass
        const long long cap = std::max<long long>(1, (long long)num_sms() * 8 / cnt);
        gx = std::min<long long>(std::max<long long>(gx, 1), std::max<long long>(cap, 8));
        rb::grad_sumsq_kernel<<<dim3((unsigned)gx, cnt), 256, 0, st>>>(L, sumsq);
        rc = check_launch("grad_sumsq_kernel");
        if (rc) return rc;
    }
    return RB_OK;
}

int rb_adamw_clip_step(const rb_opt_tensor* tensors, int count, const float* lr, const float* step, const double* sumsq,
                       float max_norm, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    if (!tensors || count <= 0 || !lr || !step) return fail(RB_ERR_INVALID, "adamw_clip_step: bad arguments");
    if (sumsq && !(max_norm > 0.f)) return fail(RB_ERR_INVALID, "adamw_clip_step: max_norm must be positive when clipping");
    cudaStream_t st = (cudaStream_t)stream;
    rb::OptHyper h{lr, step, sumsq, beta1, beta2, eps, weight_decay, max_norm};
    for (int i0 = 0; i0 < count; i0 += rb::OPT_MAX_TENSORS) {
        const int cnt = count - i0 < rb::OPT_MAX_TENSORS ? count - i0 : rb::OPT_MAX_TENSORS;
        rb::OptTensorList L;
        long long max_n;
        int rc = opt_fill(L, tensors + i0, cnt, true, &max_n);
        if (rc) return rc;
        long long gx = (max_n / 4 + 256 * 2 - 1) / (256 * 2);      // two vectors per stream, thread and pass
        const long long cap = std::max<long long>(1, (long long)num_sms() * 8 / cnt);
        gx = std::min<long long>(std::max<long long>(gx, 1), std::max<long long>(cap, 8));
        rb::adamw_clip_kernel<<<dim3((unsigned)gx, cnt), 256, 0, st>>>(L, h);
        rc = check_launch("adamw_clip_kernel");
        if (rc) return rc;
    }
    return RB_OK;
}

int rb_adamw_clip_pack_step(float* w, const float* g, float* m, float* v, void* out_f, void* out_d, int Cout, int Cin,
                            int taps, const float* lr, const float* step, const double* sumsq, float max_norm, float beta1,
                            float beta2, float eps, float weight_decay, void* stream) {
    if (!w || !g || !m || !v || !lr || !step || (!out_f && !out_d)) return fail(RB_ERR_INVALID, "adamw_clip_pack_step: null pointer");
    if (Cout <= 0 || Cin <= 0 || Cin % 32 != 0 || taps <= 0 || taps > 27)
        return fail(RB_ERR_UNSUPPORTED, "adamw_clip_pack_step: need Cin %% 32 == 0 and 1 <= taps <= 27");
    if (sumsq && !(max_norm > 0.f)) return fail(RB_ERR_INVALID, "adamw_clip_pack_step: max_norm must be positive when clipping");
    if (!aligned16(w) || !aligned16(g) || !aligned16(m) || !aligned16(v))
        return fail(RB_ERR_UNSUPPORTED, "adamw_clip_pack_step: 16-byte aligned tensors only");
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(rb::adamw_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); });
    rb::AdamPackParams p{w, g, m, v, (rb::bf16*)out_f, (rb::bf16*)out_d, Cout, Cin, taps};
    rb::OptHyper h{lr, step, sumsq, beta1, beta2, eps, weight_decay, max_norm};
    const size_t smem = (size_t)32 * (32 * taps + 2) * sizeof(rb::bf16);
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32);
    rb::adamw_pack_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p, h);
    return check_launch("adamw_pack_kernel");
}

int rb_ncdhw_to_cl(const float* src, void* dst, int NB, int C, long long S, void* stream) {
    if (!src || !dst || C <= 0 || C % 8 != 0) return fail(RB_ERR_INVALID, "ncdhw_to_cl: bad arguments");
    rb::LayoutParams p{src, (rb::bf16*)dst, S, NB, C};
    rb::ncdhw_to_cl_kernel<<<grid_for((long long)NB * (C / 8) * S, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("ncdhw_to_cl_kernel");
}

int rb_cl_to_ncdhw(const void* src, float* dst, int NB, int C, long long S, void* stream) {
    if (!src || !dst || C <= 0 || C % 8 != 0) return fail(RB_ERR_INVALID, "cl_to_ncdhw: bad arguments");
    rb::LayoutBackParams p{(const rb::bf16*)src, dst, S, NB, C};
    rb::cl_to_ncdhw_kernel<<<grid_for((long long)NB * (C / 8) * S, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("cl_to_ncdhw_kernel");
}

int rb_blend_accumulate(const float* pred, const float* weight, float* sum, float* wsum, int C, int PZ, int PY, int PX, int VZ,
                        int VY, int VX, int z0, int y0, int x0, int activation, void* stream) {
    if (!pred || !sum) return fail(RB_ERR_INVALID, "blend_accumulate: null pointer");
    if (C < 1 || PZ < 1 || PY < 1 || PX < 1 || VZ < 1 || VY < 1 || VX < 1) return fail(RB_ERR_INVALID, "blend_accumulate: bad shape");
    if (activation < 0 || activation > 2) return fail(RB_ERR_INVALID, "blend_accumulate: bad activation");
    rb::BlendParams p{pred, weight, sum, wsum, C, PZ, PY, PX, VZ, VY, VX, z0, y0, x0, activation};
    rb::blend_accumulate_kernel<<<grid_for((long long)PZ * PY * PX, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("blend_accumulate_kernel");
}

int rb_blend_finalize_cast(const float* sum, const float* wsum, void* out, float* favg, long long V, int C, int kind, void* stream) {
    if (!sum || !wsum || !out) return fail(RB_ERR_INVALID, "blend_finalize_cast: null pointer");
    if (kind != 0 && kind != 1) return fail(RB_ERR_INVALID, "blend_finalize_cast: kind must be 0 or 1");
    if (C < 1 || V < 1) return fail(RB_ERR_INVALID, "blend_finalize_cast: bad shape");
    rb::FinalizeCastParams p{sum, wsum, out, favg, V, C, kind};
    rb::blend_finalize_cast_kernel<<<grid_for(V, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("blend_finalize_cast_kernel");
}

int rb_blend_accumulate_multi(const RbBlendTarget* targets, int ntargets, const float* weight, float* wsum, int PZ, int PY, int PX,
                              int VZ, int VY, int VX, int z0, int y0, int x0, void* stream) {
    if (!targets || ntargets < 1 || ntargets > rb::BLEND_MAX_TARGETS)
        return fail(RB_ERR_INVALID, "blend_accumulate_multi: need 1..%d targets", rb::BLEND_MAX_TARGETS);
    if (PZ < 1 || PY < 1 || PX < 1 || VZ < 1 || VY < 1 || VX < 1) return fail(RB_ERR_INVALID, "blend_accumulate_multi: bad shape");
    if ((long long)PZ * PY * PX >= (1LL << 31)) return fail(RB_ERR_INVALID, "blend_accumulate_multi: patch exceeds 2^31 voxels");
    rb::BlendMultiParams p;
    memset(&p, 0, sizeof(p));
    bool vec = PX % 4 == 0 && VX % 4 == 0 && x0 % 4 == 0 && x0 >= 0 && x0 + PX <= VX && aligned16(weight) && aligned16(wsum);
    for (int i = 0; i < ntargets; ++i) {
        const RbBlendTarget& t = targets[i];
        if (!t.pred || !t.sum) return fail(RB_ERR_INVALID, "blend_accumulate_multi: null pointer in target %d", i);
        if (t.C < 1 || t.C > rb::BLEND_MAX_C) return fail(RB_ERR_INVALID, "blend_accumulate_multi: target %d needs 1..%d channels", i, rb::BLEND_MAX_C);
        if (t.activation < 0 || t.activation > 2) return fail(RB_ERR_INVALID, "blend_accumulate_multi: bad activation in target %d", i);
        p.t[i].pred = t.pred; p.t[i].sum = t.sum; p.t[i].C = t.C; p.t[i].activation = t.activation;
        vec = vec && aligned16(t.pred) && aligned16(t.sum) && ((long long)PZ * PY * PX) % 4 == 0 && ((long long)VZ * VY * VX) % 4 == 0;
    }
    p.nt = ntargets; p.weight = weight; p.wsum = wsum;
    p.PZ = PZ; p.PY = PY; p.PX = PX; p.VZ = VZ; p.VY = VY; p.VX = VX; p.z0 = z0; p.y0 = y0; p.x0 = x0;
    const long long PS = (long long)PZ * PY * PX;
    if (vec) rb::blend_accumulate_multi_kernel<4><<<grid_for(PS / 4, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
    else rb::blend_accumulate_multi_kernel<1><<<grid_for(PS, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("blend_accumulate_multi_kernel");
}

int rb_blend_finalize_cast2(const float* sum, long long sum_cstride, const float* wsum, void* out, float* favg, long long V, int C,
                            int kind, void* stream) {
    if (!sum || !wsum || !out) return fail(RB_ERR_INVALID, "blend_finalize_cast2: null pointer");
    if (kind != 0 && kind != 1) return fail(RB_ERR_INVALID, "blend_finalize_cast2: kind must be 0 or 1");
    if (C < 1 || C > rb::BLEND_MAX_C || V < 1 || sum_cstride < V) return fail(RB_ERR_INVALID, "blend_finalize_cast2: bad shape");
    rb::FinalizeCast2Params p{sum, wsum, out, favg, V, sum_cstride, C, kind};
    const bool vec = V % 4 == 0 && sum_cstride % 4 == 0 && aligned16(sum) && aligned16(wsum) && aligned16(favg) &&
                     (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
    if (vec) rb::blend_finalize_cast2_kernel<4><<<grid_for(V / 4, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
    else rb::blend_finalize_cast2_kernel<1><<<grid_for(V, 256, 16), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("blend_finalize_cast2_kernel");
}

int rb_extract_patches(const void* vol, int is_u16, int VZ, int VY, int VX, const int* origins, int nb, int PZ, int PY, int PX,
                       int standardize, double* stats, float* out, void* stream) {
    if (!vol || !out || !origins || (standardize && !stats)) return fail(RB_ERR_INVALID, "extract_patches: null pointer");
    if (nb < 1 || nb > rb::EXTRACT_MAX_BATCH) return fail(RB_ERR_INVALID, "extract_patches: need 1..%d patches per call", rb::EXTRACT_MAX_BATCH);
    if ((long long)PZ * PY * PX >= (1LL << 31)) return fail(RB_ERR_INVALID, "extract_patches: patch exceeds 2^31 voxels");
    cudaStream_t st = (cudaStream_t)stream;
    rb::ExtractBatchParams p;
    memset(&p, 0, sizeof(p));
    p.vol = vol; p.is_u16 = is_u16; p.VZ = VZ; p.VY = VY; p.VX = VX; p.PZ = PZ; p.PY = PY; p.PX = PX; p.nb = nb;
    p.stats = stats; p.out = out; p.standardize = standardize;
    for (int b = 0; b < nb; ++b) {
        const int z0 = origins[3 * b], y0 = origins[3 * b + 1], x0 = origins[3 * b + 2];
        if (z0 < 0 || y0 < 0 || x0 < 0 || z0 + PZ > VZ || y0 + PY > VY || x0 + PX > VX)
            return fail(RB_ERR_INVALID, "extract_patches: patch %d outside the volume", b);
        p.z0[b] = z0; p.y0[b] = y0; p.x0[b] = x0;
    }
    const long long PS = (long long)PZ * PY * PX;
    if (standardize) {
        RB_CUDA(cudaMemsetAsync(stats, 0, (size_t)nb * 2 * sizeof(double), st));
        rb::patch_stats_batch_kernel<<<dim3(grid_for(PS, 256, 2), nb), 256, 0, st>>>(p);
        int rc = check_launch("patch_stats_batch_kernel");
        if (rc) return rc;
    }
    rb::patch_write_batch_kernel<<<dim3(grid_for(PS, 256, 8), nb), 256, 0, st>>>(p);
    return check_launch("patch_write_batch_kernel");
}

int rb_blend_add(float* dst, const float* src, long long n, void* stream) {
    if (!dst || !src || n < 0) return fail(RB_ERR_INVALID, "blend_add: bad arguments");
    if (n == 0) return RB_OK;
    rb::blend_add_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    return check_launch("blend_add_kernel");
}

int rb_extract_patch(const void* vol, int is_u16, int VZ, int VY, int VX, int z0, int y0, int x0, int PZ, int PY, int PX,
                     int standardize, double* stats, float* out, void* stream) {
    if (!vol || !out || (standardize && !stats)) return fail(RB_ERR_INVALID, "extract_patch: null pointer");
    if (z0 < 0 || y0 < 0 || x0 < 0 || z0 + PZ > VZ || y0 + PY > VY || x0 + PX > VX) return fail(RB_ERR_INVALID, "extract_patch: patch outside the volume");
    cudaStream_t st = (cudaStream_t)stream;
    rb::ExtractParams p{vol, is_u16, VZ, VY, VX, z0, y0, x0, PZ, PY, PX, stats, out, standardize};
    const long long PS = (long long)PZ * PY * PX;
    if (standardize) {
        RB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
        rb::patch_stats_kernel<<<grid_for(PS, 256, 2), 256, 0, st>>>(p);
        int rc = check_launch("patch_stats_kernel");
        if (rc) return rc;
    }
    rb::patch_write_kernel<<<grid_for(PS, 256), 256, 0, st>>>(p);
    return check_launch("patch_write_kernel");
}


static dim3 loss_grid(long long S, int C, int NB) {
    long long bx = (S + 256 * 8 - 1) / (256 * 8);
    if (bx < 1) bx = 1;
    if (bx > 4096) bx = 4096;
    return dim3((unsigned)bx, (unsigned)C, (unsigned)NB);
}

int rb_loss_bce_dice_reduce(const float* logits, const float* target, double* stats, int NB, int C, long long S, float smoothing,
                            void* stream) {
    if (!logits || !target || !stats) return fail(RB_ERR_INVALID, "loss_bce_dice_reduce: null pointer");
    if (NB <= 0 || C <= 0 || C > 65535 || NB > 65535 || S <= 0) return fail(RB_ERR_INVALID, "loss_bce_dice_reduce: bad shape");
    rb::LossParams p{logits, target, stats, nullptr, nullptr, S, NB, C, 0.f, 0.f, 0.f, smoothing};
    rb::loss_bce_dice_reduce_kernel<<<loss_grid(S, C, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("loss_bce_dice_reduce_kernel");
}

int rb_loss_bce_dice_grad(const float* logits, const float* target, const double* stats, const float* grad_out, float* dlogits,
                          int NB, int C, long long S, float alpha, float beta, float eps, float smoothing, void* stream) {
    if (!logits || !target || !stats || !grad_out || !dlogits) return fail(RB_ERR_INVALID, "loss_bce_dice_grad: null pointer");
    if (NB <= 0 || C <= 0 || C > 65535 || NB > 65535 || S <= 0) return fail(RB_ERR_INVALID, "loss_bce_dice_grad: bad shape");
    rb::LossParams p{logits, target, const_cast<double*>(stats), dlogits, grad_out, S, NB, C, alpha, beta, eps, smoothing};
    rb::loss_bce_dice_grad_kernel<<<loss_grid(S, C, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("loss_bce_dice_grad_kernel");
}

int rb_loss_cosine_reduce(const float* pred, const float* target, double* stats, int NB, long long S, void* stream) {
    if (!pred || !target || !stats) return fail(RB_ERR_INVALID, "loss_cosine_reduce: null pointer");
    if (NB <= 0 || NB > 65535 || S <= 0) return fail(RB_ERR_INVALID, "loss_cosine_reduce: bad shape");
    rb::LossParams p{pred, target, stats, nullptr, nullptr, S, NB, 3, 0.f, 0.f, 0.f, 0.f};
    rb::loss_cosine_reduce_kernel<<<loss_grid(S, 1, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("loss_cosine_reduce_kernel");
}

int rb_loss_cosine_grad(const float* pred, const float* target, const double* stats, const float* grad_out, float* dpred, int NB,
                        long long S, void* stream) {
    if (!pred || !target || !stats || !grad_out || !dpred) return fail(RB_ERR_INVALID, "loss_cosine_grad: null pointer");
    if (NB <= 0 || NB > 65535 || S <= 0) return fail(RB_ERR_INVALID, "loss_cosine_grad: bad shape");
    rb::LossParams p{pred, target, const_cast<double*>(stats), dpred, grad_out, S, NB, 3, 0.f, 0.f, 0.f, 0.f};
    rb::loss_cosine_grad_kernel<<<loss_grid(S, 1, NB), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("loss_cosine_grad_kernel");
}

}  // extern "C"
