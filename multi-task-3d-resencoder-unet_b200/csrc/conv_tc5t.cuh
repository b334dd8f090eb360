// tcgen05 gather convolution, "weights-stationary-on-M" orientation, for layers with few output channels.
//
//   D^T[co][v] = sum_{tap, c} W[tap][co][c] * X[src(v, tap)][c]        M = co (padded to 128), N = 256 voxels
//
// Why a second orientation: an SS-mode tcgen05.mma reads its A operand from shared memory at about one
// 32-byte row per cycle, so an M=128 instruction costs >= ~128 cycles whatever N is (measured on this path:
// ~170 cycles per M=128,N=32,K=16 MMA, profiles/r1_bottleneck_experiments.md).  With the voxels on M and
// Cout = 32 on N the tensor pipe can therefore never exceed 12 % of its peak.  Putting the (zero padded)
// weights on M and 256 voxels on N makes every MMA a full-rate N=256 instruction: 32/64/128 output channels
// reach 25/50/100 % of the MMA rate instead of 12/25/50 %.
//
// Side effects that help: TMEM lane = output channel, column = voxel, so the epilogue's stores of one column
// are 32 consecutive channels (perfectly coalesced) and the InstanceNorm statistics are plain per-thread sums
// over columns (no shuffles).
//
// Same operand conventions, geometry parameters and fusions as conv_tc5.cuh (mode 0 only).
#pragma once
#include "common.cuh"

namespace rb {

struct Tc5tConvParams {
    CUtensorMap mapX[2];  // rank 5 (C, W, H, D, N), box (KW, tw.., tn) covering 256 output voxels
    CUtensorMap mapW;     // rank 3 (Ctot, Nout, taps), box (KW, wRows, 1); with several M tiles rows >= Nout are zero filled
    int nsrc, srcC[2];
    int KW;
    int tapD, tapH, tapW, offD, offH, offW, istrD, istrH, istrW;
    int lw, lh, ld;       // log2 of the tile box extents (tw, th, td); tn = 256 >> (lw + lh + ld)
    int tilesW, tilesH, tilesD, tilesNB, tilesM;
    int OW, OH, OD, NB;
    int Nout;
    int ostrD, ostrH, ostrW, ooffD, ooffH, ooffW;
    int FD, FH, FW;
    void* out0;
    void* out1;
    int outC0, outC1;
    int outF32;
    int stages;
    int hm;               // 1 = h-major tile (32 w x 8 h): X box (KW, 8 h, 34 w), smem rows ordered [w][h], so a kw shift is 8
                          // rows = one swizzle atom and the three kw taps of a (kd, kh) share ONE X load (1/3 of the TMA rows)
    int wRows;            // rows of the weight box (= min(128, Nout)): rows beyond it are never loaded - the MMA reads
                          // stale shared memory there and fills TMEM lanes no epilogue warp reads
    float* stat_sum;
    float* stat_sq;
    int statSmem;
    FastDiv fdTilesM, fdTilesW, fdTilesH, fdTilesD;
    // split-K over taps for the deep layers (few voxels, huge K): each work item takes `tapsPer` consecutive taps and
    // stores its fp32 partial tile into ws[slice][voxel][Nout]; split_finish_kernel sums the slices, stores the result
    // and accumulates the InstanceNorm statistics
    int splitK, tapsPer;
    FastDiv fdSplitK;
    float* ws;
    long long wsSlice;    // elements per workspace slice (= voxel tiles x 256 x Nout)
    int debug;   // profiling experiments only (RESENC_TC5T_DEBUG bit mask): 1 skip MMAs, 2 skip TMA loads, 4 skip the epilogue body
};

static constexpr int TC5T_THREADS = 192;
static constexpr int TC5T_VOX = 256;

__global__ void __launch_bounds__(TC5T_THREADS, 1) tc5t_gather_conv_kernel(const __grid_constant__ Tc5tConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_al + 192);
    const uint32_t err_flag = smem_u32(smem_al + 200);
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem_al + 200) = 0u;
    uint8_t* tiles = smem_al + 1024;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t rowB = (uint32_t)p.KW * 2u;
    const uint32_t bytesX = (p.hm ? 34u * 8u : (uint32_t)TC5T_VOX) * rowB;
    // hm: three weight tiles of wRows rows; the MMA of the last one reads 128 rows, so the region is 2*wRows + 128 rows
    const uint32_t bytesW = (p.hm ? (2u * (uint32_t)p.wRows + 128u) : 128u) * rowB;
    const uint32_t txBytes = bytesX + (p.hm ? 3u : 1u) * (uint32_t)p.wRows * rowB;
    const uint32_t stageBytes = bytesX + bytesW;
    const uint32_t tile_base = smem_u32(tiles);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
    auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };

    // per-epilogue-warp statistics slots behind the stages: [quad][2][NB][128 channels of this M tile]... indexed
    // by the global channel: [quad][2][NB * Nout]
    float* statS = reinterpret_cast<float*>(tiles + (size_t)S * stageBytes);
    const int statN = p.NB * p.Nout;
    if (p.stat_sum != nullptr && p.statSmem)
        for (int i = threadIdx.x; i < 8 * statN; i += blockDim.x) statS[i] = 0.f;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapX[0]);
        if (p.nsrc > 1) tma_prefetch_desc(&p.mapX[1]);
        tma_prefetch_desc(&p.mapW);
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
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

    const int tw = 1 << p.lw, th = 1 << p.lh, td = 1 << p.ld;
    const int tn = TC5T_VOX >> (p.lw + p.lh + p.ld);
    const int totalTiles = p.tilesW * p.tilesH * p.tilesD * p.tilesNB * p.tilesM * p.splitK;
    const int Ctot = p.srcC[0] + (p.nsrc > 1 ? p.srcC[1] : 0);
    const int ntaps = p.tapD * p.tapH * p.tapW;
    const bool dbgT = (p.debug & 8) && blockIdx.x == 0;   // per-role cycle counters of CTA 0 (rb_debug_counters)

    auto decode = [&](int tile, uint32_t& mt, uint32_t& tiw, uint32_t& tih, uint32_t& tid, uint32_t& tib, uint32_t& sk) {
        uint32_t sp;
        fdivmod((uint32_t)tile, p.fdSplitK, sp, sk);   // tap slice fastest
        fdivmod(sp, p.fdTilesM, sp, mt);   // M tile fastest: CTAs sharing a voxel tile run together
        fdivmod(sp, p.fdTilesW, sp, tiw);
        fdivmod(sp, p.fdTilesH, sp, tih);
        fdivmod(sp, p.fdTilesD, tib, tid);
    };

    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        long long tW = 0;
        const long long tA = clock64();
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
            uint32_t mt, tiw, tih, tid, tib, sk;
            decode(tile, mt, tiw, tih, tid, tib, sk);
            const int ow0 = tiw * tw, oh0 = tih * th, od0 = tid * td, nb0 = tib * tn;
            const int m0 = mt * 128;
            const int t0 = (int)sk * p.tapsPer, t1 = min(ntaps, t0 + p.tapsPer);
            int t = 0;
            for (int kd = 0; kd < p.tapD; ++kd) {
                const int iz = od0 * p.istrD + p.offD + kd;
                for (int kh = 0; kh < p.tapH; ++kh) {
                    const int iy = oh0 * p.istrH + p.offH + kh;
                    for (int kw = 0; kw < (p.hm ? 1 : p.tapW); ++kw, t += (p.hm ? 3 : 1)) {
                        if (t < t0 || t >= t1) continue;
                        const int ix = ow0 * p.istrW + p.offW + kw;
                        int cbase = 0;
                        for (int s = 0; s < p.nsrc; ++s) {
                            for (int c = 0; c < p.srcC[s]; c += p.KW) {
                                const long long w0 = dbgT ? clock64() : 0;
                                mbar_wait(empty_bar(stage), phase ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
                                if (dbgT) tW += clock64() - w0;
                                if (elect_one()) {
                                    const uint32_t dstX = tile_base + stage * stageBytes;
                                    const uint32_t dstW = dstX + bytesX;
                                    if (p.debug & 2) {
                                        mbar_arrive(full_bar(stage));
                                    } else {
                                        mbar_expect_tx(full_bar(stage), txBytes);
                                        if (p.hm) tma_load_5d(dstX, &p.mapX[s], full_bar(stage), c, iy, ix, iz, nb0);   // dims (c, h, w, d, n)
                                        else tma_load_5d(dstX, &p.mapX[s], full_bar(stage), c, ix, iy, iz, nb0);
                                        tma_load_3d(dstW, &p.mapW, full_bar(stage), cbase + c, m0, t);
                                    }
                                }
                                __syncwarp();
                                if (++stage == S) { stage = 0; phase ^= 1u; }
                            }
                            cbase += p.srcC[s];
                        }
                    }
                }
            }
        }
        if (dbgT && lane == 0) { g_dbg[0] = (unsigned long long)tW; g_dbg[1] = (unsigned long long)(clock64() - tA); }
    } else if (warp == 1) {
        // ===================== MMA issuer: A = weights (M = 128), B = voxels (N = 256) =====================
        const uint32_t idesc = make_idesc_bf16(128, TC5T_VOX, 0, 0);
        const uint32_t lay = swizzle_layout_code(p.KW * 2);
        const uint32_t sbo = 8u * p.KW * 2u;
        const int kPerStep = p.KW / 16;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long tWF = 0, tWT = 0, nTiles = 0;
        const long long tA = clock64();
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
            uint32_t mt_, tiw_, tih_, tid_, tib_, sk;
            decode(tile, mt_, tiw_, tih_, tid_, tib_, sk);
            const int t0 = (int)sk * p.tapsPer, t1 = min(ntaps, t0 + p.tapsPer);
            const int stepsPerTile = (p.hm ? (t1 - t0) / 3 : (t1 - t0)) * (Ctot / p.KW);
            const long long w0 = dbgT ? clock64() : 0;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u, DEVERR_WAIT_TMEM_EMPTY, err_flag);
            if (dbgT) { tWT += clock64() - w0; ++nTiles; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC5T_VOX);
            for (int ks = 0; ks < stepsPerTile; ++ks) {
                const long long w1 = dbgT ? clock64() : 0;
                mbar_wait(full_bar(stage), phase, DEVERR_WAIT_FULL, err_flag);
                if (dbgT) tWF += clock64() - w1;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t xAddr = tile_base + stage * stageBytes;
                    const uint32_t wAddr = xAddr + bytesX;
                    const int nkw = p.hm ? 3 : 1;
                    for (int kw = 0; kw < nkw; ++kw) {
                        // hm: tap kw = weight tile kw, B shifted by 8 rows (one w step of the [w][h] ordered box)
                        const uint32_t wA = wAddr + (uint32_t)kw * (uint32_t)p.wRows * rowB;
                        const uint32_t xA = xAddr + (uint32_t)kw * 8u * rowB;
                        for (int k = 0; k < kPerStep && !(p.debug & 1); ++k) {
                            const uint64_t da = make_smem_desc(wA + k * 32u, 16u, sbo, lay);
                            const uint64_t db = make_smem_desc(xA + k * 32u, 16u, sbo, lay);
                            umma_bf16(d_tmem, da, db, idesc, (ks | k | kw) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (dbgT && lane == 0) {
            g_dbg[2] = (unsigned long long)tWF; g_dbg[3] = (unsigned long long)tWT; g_dbg[4] = (unsigned long long)(clock64() - tA);
            g_dbg[11] = (unsigned long long)nTiles;
        }
    } else {
        // ===================== epilogue (warps 2..5): lane = output channel, column = voxel =====================
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int spatial = tw * th * td;   // voxels of one sample inside a tile (a multiple of 32 or < 32)
        long long tWE = 0;
        const long long tAE = clock64();
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
            uint32_t mt, tiw, tih, tid, tib, sk;
            decode(tile, mt, tiw, tih, tid, tib, sk);
            const int co = (int)mt * 128 + quad * 32 + lane;
            const bool rowValid = co < p.Nout;
            const bool warpHasRows = (int)mt * 128 + quad * 32 < p.Nout;
            const long long we0 = dbgT ? clock64() : 0;
            mbar_wait(tfull_bar(acc), acc_phase, DEVERR_WAIT_TMEM_FULL, err_flag);
            if (dbgT) tWE += clock64() - we0;
            tc_fence_after();
            if (warpHasRows && !(p.debug & 4)) {
                const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC5T_VOX);
                // destination of this lane's channel
                const bool first = co < p.outC0;
                const int cdst = first ? co : co - p.outC0;
                const int cpitch = first ? p.outC0 : p.outC1;
                void* const base = first ? p.out0 : p.out1;
                for (int cg = 0; cg < TC5T_VOX; cg += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_addr + cg, v);
                    tmem_ld_wait();
                    float s1 = 0.f, s2 = 0.f;
                    int nbStat = -1;
                    if (p.hm) {
                        // column n of the tile = voxel (w = n >> 3, h = n & 7); plain stride-1 convolution (checked by the host)
                        const int ow0 = (int)tiw * 32 + (cg >> 3), oh0 = (int)tih * 8, od = (int)tid, nb = (int)tib;
                        nbStat = nb;
                        const size_t e0 = ((((size_t)nb * p.OD + od) * p.OH + oh0) * p.OW + ow0) * cpitch + cdst;
                        const int rs = p.OW * cpitch;
                        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
                        if (oh0 + 8 <= p.OH && ow0 + 4 <= p.OW) {
                            if (p.outF32 == 1) {
                                float* dst = reinterpret_cast<float*>(base) + e0;
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (rowValid) dst[(j & 7) * rs + (j >> 3) * cpitch] = __uint_as_float(v[j]);
                            } else {
                                unsigned short* dst = reinterpret_cast<unsigned short*>(base) + e0;
                                const bool h16 = p.outF32 == 2;
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (rowValid) dst[(j & 7) * rs + (j >> 3) * cpitch] = cvt16(__uint_as_float(v[j]), h16);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float x = __uint_as_float(v[j]);
                                a1[j & 3] += x;
                                a2[j & 3] = fmaf(x, x, a2[j & 3]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (oh0 + (j & 7) < p.OH && ow0 + (j >> 3) < p.OW) {
                                    const float x = __uint_as_float(v[j]);
                                    a1[j & 3] += x;
                                    a2[j & 3] = fmaf(x, x, a2[j & 3]);
                                    if (rowValid) {
                                        const size_t e = e0 + (size_t)((j & 7) * rs + (j >> 3) * cpitch);
                                        if (p.outF32 == 1) reinterpret_cast<float*>(base)[e] = x;
                                        else reinterpret_cast<unsigned short*>(base)[e] = cvt16(x, p.outF32 == 2);
                                    }
                                }
                            }
                        }
                        s1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
                        s2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
                    } else if (p.splitK > 1) {
                        // partial tile of one tap slice -> its own workspace slice, in TILE-LOCAL order
                        // ws[sk][voxel tile][column][Nout]: one address per column group and a constant stride, straight-line
                        // coalesced stores (lane = channel).  Round 1 decomposed every column into (n, d, h, w), bounds-checked
                        // it and added it into ONE buffer with red.global: a dependent 64-bit integer chain + a branch per
                        // column with a single warp per scheduler (~200 cycles per column, 60-70 % of the kernel,
                        // profiles/r2_deep_probe.txt).  split_finish_kernel maps voxels back to (tile, column), sums the
                        // slices and takes the statistics.
                        const size_t tileLin = (((size_t)tib * p.tilesD + tid) * p.tilesH + tih) * p.tilesW + tiw;
                        float* dst = p.ws + (size_t)sk * p.wsSlice + (tileLin * TC5T_VOX + cg) * p.Nout + co;
                        if (rowValid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) dst[j * p.Nout] = __uint_as_float(v[j]);
                        }
                    } else if (p.lw >= 5) {
                        // the 32 columns of this group are 32 consecutive voxels of one W row: one address
                        // computation per group, then a constant stride per column
                        const int r0 = cg;
                        const int iw0 = r0 & (tw - 1);
                        const int ih = (r0 >> p.lw) & (th - 1);
                        const int id = (r0 >> (p.lw + p.lh)) & (td - 1);
                        const int in = r0 >> (p.lw + p.lh + p.ld);
                        const int ow0 = (int)tiw * tw + iw0, oh = (int)tih * th + ih, od = (int)tid * td + id,
                                  nb = (int)tib * tn + in;
                        const bool gvalid = (oh < p.OH) && (od < p.OD) && (nb < p.NB);
                        if (gvalid) {
                            nbStat = nb;
                            const int fd = od * p.ostrD + p.ooffD, fh = oh * p.ostrH + p.ooffH, fw0 = ow0 * p.ostrW + p.ooffW;
                            const size_t vox0 = (((size_t)nb * p.FD + fd) * p.FH + fh) * p.FW + fw0;
                            const size_t e0 = vox0 * cpitch + cdst;
                            const int estep = p.ostrW * cpitch;       // elements between consecutive columns (fits 32 bits)
                            const int nvalid = min(32, p.OW - ow0);   // columns beyond the row end are padding
                            // straight-line stores and four independent statistics chains: per-column branches and a
                            // single dependent FADD/FFMA chain made this loop ~85 cycles per column (cycle counters)
                            float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
                            if (nvalid == 32) {
                                if (p.outF32 == 1) {
                                    float* dst = reinterpret_cast<float*>(base) + e0;
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (rowValid) dst[j * estep] = __uint_as_float(v[j]);
                                } else {
                                    unsigned short* dst = reinterpret_cast<unsigned short*>(base) + e0;
                                const bool h16 = p.outF32 == 2;
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (rowValid) dst[j * estep] = cvt16(__uint_as_float(v[j]), h16);
                                }
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float x = __uint_as_float(v[j]);
                                    a1[j & 3] += x;
                                    a2[j & 3] = fmaf(x, x, a2[j & 3]);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    if (j < nvalid) {
                                        const float x = __uint_as_float(v[j]);
                                        a1[j & 3] += x;
                                        a2[j & 3] = fmaf(x, x, a2[j & 3]);
                                        if (rowValid) {
                                            if (p.outF32 == 1) reinterpret_cast<float*>(base)[e0 + (size_t)j * estep] = x;
                                            else reinterpret_cast<unsigned short*>(base)[e0 + (size_t)j * estep] = cvt16(x, p.outF32 == 2);
                                        }
                                    }
                                }
                            }
                            s1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
                            s2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int r = cg + j;                       // voxel row inside the tile
                        const int iw = r & (tw - 1);
                        const int ih = (r >> p.lw) & (th - 1);
                        const int id = (r >> (p.lw + p.lh)) & (td - 1);
                        const int in = r >> (p.lw + p.lh + p.ld);
                        const int ow = (int)tiw * tw + iw, oh = (int)tih * th + ih, od = (int)tid * td + id,
                                  nb = (int)tib * tn + in;
                        const bool valid = (ow < p.OW) && (oh < p.OH) && (od < p.OD) && (nb < p.NB);
                        const float x = __uint_as_float(v[j]);
                        if (valid) {
                            if (p.stat_sum != nullptr) {
                                if (spatial >= 32) {               // the 32 columns of this group share one sample
                                    s1 += x; s2 += x * x; nbStat = nb;
                                } else if (rowValid) {
                                    atomicAdd(p.stat_sum + nb * p.Nout + co, x);
                                    atomicAdd(p.stat_sq + nb * p.Nout + co, x * x);
                                }
                            }
                            if (rowValid) {
                                const int fd = od * p.ostrD + p.ooffD, fh = oh * p.ostrH + p.ooffH, fw = ow * p.ostrW + p.ooffW;
                                const size_t vox = (((size_t)nb * p.FD + fd) * p.FH + fh) * p.FW + fw;
                                if (p.outF32 == 1) reinterpret_cast<float*>(base)[vox * cpitch + cdst] = x;
                                else reinterpret_cast<unsigned short*>(base)[vox * cpitch + cdst] = cvt16(x, p.outF32 == 2);
                            }
                        }
                    }
                    }
                    if (p.stat_sum != nullptr && spatial >= 32 && nbStat >= 0 && rowValid) {
                        const int idx = nbStat * p.Nout + co;
                        if (p.statSmem) {
                            float* slot = statS + (size_t)quad * 2 * statN + idx;
                            slot[0] += s1;
                            slot[statN] += s2;
                        } else {
                            atomicAdd(p.stat_sum + idx, s1);
                            atomicAdd(p.stat_sq + idx, s2);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (dbgT && quad == 0 && lane == 0) { g_dbg[8] = (unsigned long long)tWE; g_dbg[10] = (unsigned long long)(clock64() - tAE); }
    }

    tc_fence_before();
    __syncthreads();
    if (p.stat_sum != nullptr && p.statSmem) {
        for (int i = threadIdx.x; i < statN; i += blockDim.x) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                a += statS[(size_t)w * 2 * statN + i];
                b += statS[(size_t)w * 2 * statN + statN + i];
            }
            if (a != 0.f || b != 0.f) {
                atomicAdd(p.stat_sum + i, a);
                atomicAdd(p.stat_sq + i, b);
            }
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace rb
