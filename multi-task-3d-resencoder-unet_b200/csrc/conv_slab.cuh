// tcgen05 "slab" convolution for the full-resolution 32-channel layers: 3x3x3, stride 1, padding 1, 32 input
// channels (one source) -> 32 output channels, W in {32, 64, 96, 128} (a tile = R = floor(256 / W) full rows = N = R * W
// voxels: 256, or 192 for W = 96).
//
// Why a third kernel: with 32 input channels every (voxel, tap) operand row is only 64 bytes.  The per-tap
// gather kernels (conv_tc5.cuh / conv_tc5t.cuh) re-load a tile's input once per tap, i.e. 27 TMA box rows per
// output voxel, and measured at ~2 cycles per box row that alone is > 2x the tensor-pipe time of the tile
// (profiles/README.md, "32->32 @128^3").  This kernel loads every input row ONCE per tile and never shifts
// an operand by less than a whole W row:
//
//   * a CTA walks along z.  Its shared memory holds a ring of 4 input planes, each (R+2) x W voxels x 32 ch
//     (R = 256 / W output rows per tile, one halo row either side; the zero padding is TMA out-of-bounds fill),
//     plus all 27 x 32 x 32 weights (55 KB, loaded once);
//   * B operand of an MMA = 256 consecutive voxels (R full rows) of plane d+kd-1 starting at row kh: a plain
//     K-major SW64 tile whose start is a multiple of W*64 bytes - no halo columns, no sub-row shifts;
//   * A operand = the weights of the three kw taps of (kd, kh) stacked on M: row kw*32 + co, and the centre tap
//     (kw = 1) once more in rows 96..127.  So 18 MMAs (9 (kd,kh) x 2 K steps) of M=128 x N=256 cover all 27
//     taps: TMEM lanes [32kw, 32kw+32) (and lanes 96..127 again for kw = 1) hold
//         P_kw[co][v] = sum_{kd,kh,ci} W[kd,kh,kw][co][ci] * X[d+kd-1][h+kh-1][w(v)][ci]
//   * the kw shift is applied where it is free, in the epilogue: out[co][(h,w)] = P_0[(h,w-1)] + P_1[(h,w)] +
//     P_2[(h,w+1)], terms with w-1 < 0 or w+1 >= W dropped (that IS the zero padding in W).  Epilogue warp q
//     The two "side" warps (TMEM lane quarters 0 and 2) write their columns, offset by +1 / -1, into a double-
//     buffered staging set in shared memory; the "centre" warp (quarter 1) adds them to its own columns in
//     registers (lane = output channel, so a store of one voxel is 32 consecutive channels) and stores (+ the
//     InstanceNorm sum / sum-of-squares of the stored values, + an optional "+= existing output" for the second
//     source of a virtual concat).  Two such warp triples split the 256 columns of a tile.
//
// Per 256 output voxels: 1 TMA box of (R+2)*W rows instead of 27 boxes of 256 rows, 18 MMAs instead of 54.
#pragma once
#include "common.cuh"

namespace rb {

struct SlabConvParams {
    CUtensorMap mapX;   // rank 5 (32, W, H, D, N), box (32, W, R+2, 1, 1), SWIZZLE_64B
    CUtensorMap mapW;   // rank 3 (Ctot, Mtot, 27), box (32, 32, 1), SWIZZLE_64B
    int c0, m0;         // weight box origin: input-channel offset (source), output-row offset (destination half)
    int W, H, D, NB, lw;
    int R;              // output rows per tile (256 / W)
    int DC;             // output planes per work item
    int hTiles, dChunks;
    FastDiv fdH, fdDC;
    void* out;          // channels-last destination with exactly 32 channels per voxel
    float* stat_sum;    // [NB][statPitch] at channel statC0 + co, or nullptr
    float* stat_sq;
    int statPitch, statC0;
    int debug;   // profiling experiments (RESENC_SLAB_DEBUG bit mask): 1 skip MMAs, 2 skip plane loads, 4 skip epilogue, 8 skip staging, 16 skip output, 64 no prefetch of the existing output (accumulate modes)
};

static constexpr int SLAB_THREADS = 320;   // warps: 0 TMA, 1 MMA, 2..5 and 6..9 epilogue (TMEM lane quarter = warp & 3)
static constexpr int SLAB_WBYTES = 9 * 128 * 64;       // per (kd,kh): rows kw0, kw1, kw2, kw1 again (32 co each), SW64 rows of 32 bf16

static constexpr int SLAB_CHUNK = 16;                        // voxels per staging hand-off
static constexpr int SLAB_SPITCH = SLAB_CHUNK + 4;           // floats per channel row of a staging buffer (+4 pad: conflict-free 16-byte accesses)
static constexpr int SLAB_SBUF = 32 * SLAB_SPITCH;            // one buffer: [32 channels][16 voxels (+pad)]
static constexpr int SLAB_SGROUP = 2 * 2 * SLAB_SBUF;         // per column group: 2 sets x 2 sides
static constexpr int SLAB_OBYTES = 2 * SLAB_SGROUP * 4;   // 2 column groups x 2 sets x 2 sides x [32 ch][16 voxels + 4 pad] fp32

// Side warps of the epilogue (TMEM lane quarter Q = kw = 0 or 2), 64 columns = four 16-voxel chunks per call:
// voxel a + i receives column a + i + Q - 1 of this warp's partial sums (zero where that column is W padding), written
// as [channel = lane][voxel] rows with 16-byte stores into side buffer Q/2 of staging set (cnt & 1).  Even chunks
// (set 0) are consumed by the centre warp of lane quarter 1, odd chunks (set 1) by the one of quarter 3.
template <int Q, int NCOLS, bool YH>
__device__ __forceinline__ void slab_side64(uint32_t t_addr, int a, float* stg, int lane, int Wm, uint32_t sfull0,
                                            uint32_t sempty0, uint32_t& cnt, uint32_t err_flag, bool skip, long long* tWait, const void* yrow) {
    uint32_t v0[32], v1[32], ex[1];
    ex[0] = 0u;
    // accumulate mode (second source of a virtual concat): this side warp also adds the existing fp32 / fp16 (YH) output of half of
    // its 64 voxels (Q = 0: the first 32, Q = 2: the last 32) - the loads overlap the TMEM loads, and the side warps
    // have the slack the centre warps lack
    float y[32];
    if (yrow != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (YH) {
                const unsigned short hv = __ldcs(reinterpret_cast<const unsigned short*>(yrow) + (size_t)((Q == 0 ? 0 : 32) + i) * 32);
                y[i] = f16lo((uint32_t)hv);
            } else {
                y[i] = __ldcs(reinterpret_cast<const float*>(yrow) + (size_t)((Q == 0 ? 0 : 32) + i) * 32);
            }
        }
    }
    tmem_ld_32x32b_x32(t_addr + a, v0);
    tmem_ld_32x32b_x32(t_addr + a + 32, v1);
    if (Q == 0 && a > 0) tmem_ld_32x32b_x1(t_addr + a - 1, ex);
    if (Q == 2 && a + 64 < NCOLS) tmem_ld_32x32b_x1(t_addr + a + 64, ex);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const uint32_t b = cnt & 1u;
        const long long w0 = tWait ? clock64() : 0;
        mbar_wait(sempty0 + 8u * b, ((cnt >> 1) & 1u) ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
        if (tWait) *tWait += clock64() - w0;
        float* dst = stg + b * (2 * SLAB_SBUF) + (Q / 2) * SLAB_SBUF + lane * SLAB_SPITCH;
        if (!skip) {
#pragma unroll
            for (int vec = 0; vec < SLAB_CHUNK / 4; ++vec) {
                float f[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = ch * SLAB_CHUNK + vec * 4 + k;
                    const int c = i + Q - 1;   // column relative to a
                    const uint32_t x = c < 0 ? ex[0] : c < 32 ? v0[c & 31] : c < 64 ? v1[c & 31] : ex[0];
                    f[k] = __uint_as_float(x);
                    // only the first / last voxel of a W row can be padding, and rows start at multiples of 32 (W % 32 == 0).
                    // NCOLS == 256: W is a power of two (mask Wm = W - 1); NCOLS == 192: W = 96, two rows per tile
                    if (NCOLS == 256) {
                        if (Q == 0 && (i & 31) == 0 && ((a + i) & Wm) == 0) f[k] = 0.f;
                        if (Q == 2 && (i & 31) == 31 && ((a + i) & Wm) == Wm) f[k] = 0.f;
                    } else {
                        if (Q == 0 && (i & 31) == 0 && ((a + i) % 96) == 0) f[k] = 0.f;
                        if (Q == 2 && (i & 31) == 31 && ((a + i) % 96) == 95) f[k] = 0.f;
                    }
                    if (yrow != nullptr && (i >> 5) == (Q == 0 ? 0 : 1)) f[k] += y[i & 31];
                }
                *reinterpret_cast<float4*>(dst + vec * 4) = make_float4(f[0], f[1], f[2], f[3]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sfull0 + 8u * b);
        ++cnt;
    }
}

// MODE: 0 = bf16 destination, 1 = fp32 destination, 2 = fp32 destination, out += result,
//       3 = fp16 destination (saturating), 4 = fp16 destination, out += result
// NCOLS: accumulator columns of a tile = R * W (256 for W in {32, 64, 128}, 192 for W = 96)
template <int MODE, int NCOLS = 256>
__global__ void __launch_bounds__(SLAB_THREADS, 1) slab_conv_kernel(const __grid_constant__ SlabConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_al + 192);
    const uint32_t err_flag = smem_u32(smem_al + 200);
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem_al + 200) = 0u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t slotBytes = (uint32_t)(p.R + 2) * p.W * 64u;
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t w_base = smem_u32(smem_al + 1024);
    const uint32_t ring_base = w_base + SLAB_WBYTES;
    float* const O = reinterpret_cast<float*>(smem_al + 1024 + SLAB_WBYTES + 4u * slotBytes);
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 32u + 8u * s; };
    auto tfull_bar = [&](int a) { return bar_base + 64u + 8u * a; };
    auto tempty_bar = [&](int a) { return bar_base + 80u + 8u * a; };
    const uint32_t w_bar = bar_base + 96u;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapX);
        tma_prefetch_desc(&p.mapW);
        for (uint32_t s = 0; s < 4; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        for (uint32_t gb = 0; gb < 2; ++gb)
            for (uint32_t b = 0; b < 2; ++b) {
                mbar_init(bar_base + 104u + 32u * gb + 8u * b, 2);         // staging set full: both side warps
                mbar_init(bar_base + 104u + 32u * gb + 16u + 8u * b, 1);   // staging set free: the centre warp
            }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), 512);   // two 256-column fp32 accumulators
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int totalItems = p.NB * p.dChunks * p.hTiles;
    const bool dbgT = (p.debug & 32) && blockIdx.x == 0;   // per-role cycle counters of CTA 0 (rb_debug_counters)
    // item -> (sample, z chunk, h tile); h tile fastest so that CTAs running together share halo rows in L2
    auto decode = [&](int item, int& n, int& d0, int& nOut, int& h0) {
        uint32_t q, hp, dc, nn;
        fdivmod((uint32_t)item, p.fdH, q, hp);
        fdivmod(q, p.fdDC, nn, dc);
        n = (int)nn;
        d0 = (int)dc * p.DC;
        nOut = min(p.DC, p.D - d0);
        h0 = (int)hp * p.R;
    };

    if (warp == 0) {
        // ===================== TMA producer: the weights once, then one plane per ring slot =====================
        if (elect_one()) {
            mbar_expect_tx(w_bar, SLAB_WBYTES);
            for (int gk = 0; gk < 9; ++gk) {   // rows of group gk: taps 3gk, 3gk+1, 3gk+2, then 3gk+1 (centre) again
                for (int r = 0; r < 4; ++r)
                    tma_load_3d(w_base + (uint32_t)(gk * 128 + r * 32) * 64u, &p.mapW, w_bar, p.c0, p.m0, gk * 3 + (r == 3 ? 1 : r));
            }
        }
        __syncwarp();
        uint32_t idx = 0;   // planes issued so far
        long long tW = 0;
        const long long tA = clock64();
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            int n, d0, nOut, h0;
            decode(item, n, d0, nOut, h0);
            for (int pl = d0 - 1; pl <= d0 + nOut; ++pl, ++idx) {
                const uint32_t slot = idx & 3u;
                const long long w0 = dbgT ? clock64() : 0;
                mbar_wait(empty_bar(slot), ((idx >> 2) & 1u) ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
                if (dbgT) tW += clock64() - w0;
                if (elect_one()) {
                    if ((p.debug & 2)) {
                        mbar_arrive(full_bar(slot));
                    } else {
                        mbar_expect_tx(full_bar(slot), slotBytes);
                        tma_load_5d(ring_base + slot * slotBytes, &p.mapX, full_bar(slot), 0, 0, h0 - 1, pl, n);
                    }
                }
                __syncwarp();
            }
        }
        if (dbgT && lane == 0) { g_dbg[0] = (unsigned long long)tW; g_dbg[1] = (unsigned long long)(clock64() - tA); }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_bf16(128, NCOLS, 0, 0);
        const uint32_t lay = swizzle_layout_code(64);
        const uint32_t rowBytes = (uint32_t)p.W * 64u;
        mbar_wait(w_bar, 0u, DEVERR_WAIT_FULL, err_flag);
        uint32_t kbase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long tWF = 0, tWT = 0;
        const long long tA = clock64();
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            int n, d0, nOut, h0;
            decode(item, n, d0, nOut, h0);
            for (int j = 0; j < nOut; ++j) {
                long long w0 = dbgT ? clock64() : 0;
                if (j == 0) {
                    mbar_wait(full_bar(kbase & 3u), (kbase >> 2) & 1u, DEVERR_WAIT_FULL, err_flag);
                    mbar_wait(full_bar((kbase + 1u) & 3u), ((kbase + 1u) >> 2) & 1u, DEVERR_WAIT_FULL, err_flag);
                }
                const uint32_t newest = kbase + (uint32_t)j + 2u;
                mbar_wait(full_bar(newest & 3u), (newest >> 2) & 1u, DEVERR_WAIT_FULL, err_flag);
                if (dbgT) { const long long w1 = clock64(); tWF += w1 - w0; w0 = w1; }
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, DEVERR_WAIT_TMEM_EMPTY, err_flag);
                if (dbgT) tWT += clock64() - w0;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                    uint32_t first = 0u;
#pragma unroll
                    for (int kd = 0; kd < 3 && !(p.debug & 1); ++kd) {
                        const uint32_t plane = ring_base + ((kbase + (uint32_t)(j + kd)) & 3u) * slotBytes;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint32_t aAddr = w_base + (uint32_t)(kd * 3 + kh) * (128u * 64u);
                            const uint32_t bAddr = plane + (uint32_t)kh * rowBytes;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint64_t da = make_smem_desc(aAddr + k * 32u, 16u, 512u, lay);
                                const uint64_t db = make_smem_desc(bAddr + k * 32u, 16u, 512u, lay);
                                umma_bf16(d_tmem, da, db, idesc, first);
                                first = 1u;
                            }
                        }
                    }
                    umma_commit(tfull_bar(acc));
                    umma_commit(empty_bar((kbase + (uint32_t)j) & 3u));
                    if (j == nOut - 1) {   // the last output of the item releases its two trailing planes as well
                        umma_commit(empty_bar((kbase + (uint32_t)j + 1u) & 3u));
                        umma_commit(empty_bar((kbase + (uint32_t)j + 2u) & 3u));
                    }
                }
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            kbase += (uint32_t)nOut + 2u;
        }
        if (dbgT && lane == 0) {
            g_dbg[2] = (unsigned long long)tWF; g_dbg[3] = (unsigned long long)tWT; g_dbg[4] = (unsigned long long)(clock64() - tA);
        }
    } else {
        // ===================== epilogue: warps 2..5 own columns 0..127 of a tile, warps 6..9 columns 128..255 ==========
        // out[(h,w)] = P_0[(h,w-1)] + P_1[(h,w)] + P_2[(h,w+1)].  TMEM lane quarter q = warp & 3 = kw.  The two side warps
        // (q = 0, 2) hand their shifted columns, 16 voxels at a time, to the centre warps (q = 1 and, through the second copy
        // of the centre weights, q = 3) through one staging set each; a centre warp adds its own columns (lane = output
        // channel), accumulates the statistics and stores.  Two centre quarters = two warp schedulers for the store work.
        const int q = warp & 3;
        const int g = warp >= 6 ? 1 : 0;   // warps 2..5 / 6..9
        const int Wm = p.W - 1;
        constexpr int ncols = NCOLS;           // valid accumulator columns of a tile
        float* const stg = O + g * SLAB_SGROUP;
        const uint32_t sfull0 = bar_base + 104u + 32u * g, sempty0 = sfull0 + 16u;
        uint32_t cnt = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long tWT = 0, tWS = 0, nT = 0, tLd = 0, tSt = 0;
        const long long tA = clock64();
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            int n, d0, nOut, h0;
            decode(item, n, d0, nOut, h0);
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < nOut; ++j) {
                const size_t voxT = (((size_t)n * p.D + (d0 + j)) * p.H + h0) * (size_t)p.W;   // voxel of tile column 0
                if ((MODE == 2 || MODE == 4) && (q == 0 || q == 2) && !(p.debug & 64)) {
                    // accumulate modes: a side warp adds the existing output of 32 voxels per 64-column call (slab_side64).
                    // Those loads used to be issued inside the call, their DRAM latency (two calls per tile) exposed on the
                    // epilogue's critical path: 0.82 ms against 0.49 ms for the plain store mode.  The addresses depend on
                    // the tile only, so both calls' lines are requested here, BEFORE the wait for the tile's MMAs.
                    constexpr int ESZ = MODE == 2 ? 4 : 2;
                    constexpr int LINES = 32 * 32 * ESZ / 128;           // 128-byte lines per call: 16 (fp16) / 32 (fp32)
                    const char* yb = reinterpret_cast<const char*>(p.out) + (voxT + (size_t)(g * 128 + (q == 0 ? 0 : 32))) * (32 * ESZ);
#pragma unroll
                    for (int c = 0; c < 2 * LINES; c += 32) {
                        const int idx = c + lane, call = idx / LINES, line = idx % LINES;
                        if (g * 128 + call * 64 < ncols)
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(yb + (size_t)call * (64 * 32 * ESZ) + (size_t)line * 128));
                    }
                }
                const long long w0 = dbgT ? clock64() : 0;
                mbar_wait(tfull_bar(acc), acc_phase, DEVERR_WAIT_TMEM_FULL, err_flag);
                if (dbgT) { tWT += clock64() - w0; ++nT; }
                tc_fence_after();
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
                if (p.debug & 4) {
                } else if (q == 0) {
                    for (int a = g * 128; a < (g * 128 + 128 < ncols ? g * 128 + 128 : ncols); a += 64)
                        slab_side64<0, NCOLS, MODE == 4>(t_addr, a, stg, lane, Wm, sfull0, sempty0, cnt, err_flag, (p.debug & 8) != 0, dbgT ? &tWS : nullptr,
                                       MODE == 2 ? (const void*)(reinterpret_cast<const float*>(p.out) + (voxT + a) * 32 + lane) :
                                       MODE == 4 ? (const void*)(reinterpret_cast<const unsigned short*>(p.out) + (voxT + a) * 32 + lane) : nullptr);
                } else if (q == 2) {
                    for (int a = g * 128; a < (g * 128 + 128 < ncols ? g * 128 + 128 : ncols); a += 64)
                        slab_side64<2, NCOLS, MODE == 4>(t_addr, a, stg, lane, Wm, sfull0, sempty0, cnt, err_flag, (p.debug & 8) != 0, dbgT ? &tWS : nullptr,
                                       MODE == 2 ? (const void*)(reinterpret_cast<const float*>(p.out) + (voxT + a) * 32 + lane) :
                                       MODE == 4 ? (const void*)(reinterpret_cast<const unsigned short*>(p.out) + (voxT + a) * 32 + lane) : nullptr);
                } else {
                    // centre warp of lane quarter 1 (even chunks, staging set 0) or 3 (odd chunks, set 1): both quarters hold
                    // the kw = 1 partial sums because the centre weights sit in M rows 32..63 and again in rows 96..127
                    const uint32_t b = q == 1 ? 0u : 1u;
                    for (int a = g * 128; a < (g * 128 + 128 < ncols ? g * 128 + 128 : ncols); a += 64) {
                        uint32_t va[16], vb[16];
                        const long long l0 = dbgT ? clock64() : 0;
                        tmem_ld_32x32b_x16(t_addr + a + (int)b * SLAB_CHUNK, va);
                        tmem_ld_32x32b_x16(t_addr + a + (int)b * SLAB_CHUNK + 32, vb);
                        tmem_ld_wait();
                        if (dbgT) tLd += clock64() - l0;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const long long w1 = dbgT ? clock64() : 0;
                            mbar_wait(sfull0 + 8u * b, cnt & 1u, DEVERR_WAIT_FULL, err_flag);
                            if (dbgT) tWS += clock64() - w1;
                            const float* L = stg + b * (2 * SLAB_SBUF) + lane * SLAB_SPITCH;
                            // destinations are 32-channel tensors: voxel (a + half*32 + b*16 + k), channel lane
                            const size_t e0 = (voxT + a + half * 32 + (int)b * SLAB_CHUNK) * 32 + lane;
                            const long long c0 = dbgT ? clock64() : 0;
                            if (!(p.debug & 16)) {
                                float x[SLAB_CHUNK];
#pragma unroll
                                for (int vec = 0; vec < SLAB_CHUNK / 4; ++vec) {
                                    const float4 l = *reinterpret_cast<const float4*>(L + vec * 4);
                                    const float4 r = *reinterpret_cast<const float4*>(L + SLAB_SBUF + vec * 4);
                                    x[vec * 4 + 0] = __uint_as_float(half ? vb[vec * 4 + 0] : va[vec * 4 + 0]) + (l.x + r.x);
                                    x[vec * 4 + 1] = __uint_as_float(half ? vb[vec * 4 + 1] : va[vec * 4 + 1]) + (l.y + r.y);
                                    x[vec * 4 + 2] = __uint_as_float(half ? vb[vec * 4 + 2] : va[vec * 4 + 2]) + (l.z + r.z);
                                    x[vec * 4 + 3] = __uint_as_float(half ? vb[vec * 4 + 3] : va[vec * 4 + 3]) + (l.w + r.w);
                                }
                                if (MODE == 0 || MODE >= 3) {
                                    unsigned short* gp = reinterpret_cast<unsigned short*>(p.out) + e0;
#pragma unroll
                                    for (int ii = 0; ii < SLAB_CHUNK; ++ii) gp[ii * 32] = cvt16(x[ii], MODE >= 3);
                                } else {
                                    float* gp = reinterpret_cast<float*>(p.out) + e0;
#pragma unroll
                                    for (int ii = 0; ii < SLAB_CHUNK; ++ii) gp[ii * 32] = x[ii];
                                }
                                float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                                for (int ii = 0; ii < SLAB_CHUNK; ++ii) {
                                    a1[ii & 3] += x[ii];
                                    a2[ii & 3] += x[ii] * x[ii];
                                }
                                s1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
                                s2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
                            }
                            if (dbgT) tSt += clock64() - c0;
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sempty0 + 8u * b);
                            ++cnt;   // uses of this warp's staging set
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if ((q & 1) && p.stat_sum != nullptr) {
                const int idx = n * p.statPitch + p.statC0 + lane;
                atomicAdd(p.stat_sum + idx, s1);
                atomicAdd(p.stat_sq + idx, s2);
            }
        }
        if (dbgT && lane == 0 && warp == 5) {
            g_dbg[8] = (unsigned long long)tWT; g_dbg[9] = (unsigned long long)tWS; g_dbg[10] = (unsigned long long)(clock64() - tA);
            g_dbg[11] = (unsigned long long)nT; g_dbg[12] = (unsigned long long)tLd; g_dbg[13] = (unsigned long long)tSt;
        }
        if (dbgT && lane == 0 && warp == 4) {
            g_dbg[5] = (unsigned long long)tWT; g_dbg[6] = (unsigned long long)tWS; g_dbg[7] = (unsigned long long)(clock64() - tA);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace rb
