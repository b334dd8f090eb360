// tcgen05 / TMEM / TMA implicit-GEMM "gather convolution" for sm_100a.
//
// One kernel covers every dense contraction of the ResEnc U-Net forward and data-gradient:
//   out[m, n] = sum_{tap, c} A[src(m, tap), c] * W[tap][n][c]
// where m runs over a grid of output voxels (N, OD, OH, OW), the input coordinate of a tap
// is  i = o * istr + off + k  (out-of-range => 0, provided by TMA zero fill), A is the
// channels-last bf16 activation (optionally two tensors concatenated along C) and W the
// packed bf16 weight [taps][Nout][Ctot].
//
// Reference ops realised through it (reference file:line):
//   Conv3d k3 s1/s2 fprop        simple_conv_blocks.py:43-51   (istr = stride, off = -1)
//   Conv3d k1 fprop (skip, bottleneck)  resblocks.py:97-100,187-195
//   Conv3d dgrad (stride 1: flipped taps; stride 2: 8 output-parity classes)
//   ConvTranspose3d k=s=2 fprop   decoder.py:110-113 (pixel-shuffle epilogue, N = 8*Cout)
//   ConvTranspose3d dgrad         (= k2 s2 p0 conv)
//   cat((up, skip), 1)            decoder.py:147 (two A sources, no copy)
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2..5 epilogue.
// A tile  = 128 output voxels (tw x th x td x tn box) x KW channels, landed by one 5-D TMA
// box per (tap, channel chunk) in the canonical K-major swizzled layout UMMA consumes.
// Accumulators live in TMEM, double buffered (2 x Ntile fp32 columns) so the epilogue of tile
// i overlaps the MMAs of tile i+1.  Persistent grid, static round-robin tile schedule.
#pragma once
#include "common.cuh"

namespace rb {

struct Tc5ConvParams {
    CUtensorMap mapA[2];  // rank 5 (C, W, H, D, N)
    CUtensorMap mapB;     // rank 3 (Ctot, Nout, taps)
    int nsrc, srcC[2];
    int KW;  // K elements per pipeline step (16/32/64) == swizzle span / 2
    int tapD, tapH, tapW;
    int offD, offH, offW;
    int istrD, istrH, istrW;
    int tw, th, td, tn;  // tile box, product 128
    int tilesW, tilesH, tilesD, tilesNB;
    int OW, OH, OD, NB;  // output class grid
    int Nout, Ntile, nTilesN;
    int mode;  // 0 direct, 1 pixel shuffle (column = parity * psC + channel)
    int ostrD, ostrH, ostrW, ooffD, ooffH, ooffW;
    int FD, FH, FW;  // full output spatial dims
    void* out0;
    void* out1;
    int outC0, outC1;  // channel split of the destination (out1 may be null)
    int psC;
    int psD, psH, psW;  // pixel-shuffle factors per dim (1 or 2)
    int stages;
    float* stat_sum;  // optional [NB][Nout] per-(n,c) sum of outputs   (fp32, atomics)
    float* stat_sq;   // optional [NB][Nout] per-(n,c) sum of squares
    int outF32;       // destination element type: 0 bf16, 1 fp32 (full accumulator), 2 fp16 (pre-norm activations)
    int statSmem;     // 1: statistics are accumulated in shared memory per CTA and flushed once at the end
    FastDiv fdTilesN, fdTilesW, fdTilesH, fdTilesD;   // tile index decode without integer division
    FastDiv fdTw, fdTwTh, fdTwThTd;                    // row -> (iw, ih, id, in) inside a tile
    int tps;          // taps (along W) per pipeline stage: 1, or tapW for narrow-K layers so that one mbarrier round
                      // trip feeds tps * KW/16 MMAs instead of KW/16
    int debug;        // profiling experiments only: 1 = skip the MMAs, 2 = skip the TMA loads (results are garbage)
};

// Column sums over the 32 rows held by the lanes of a warp: v[j] (lane = row) -> lane j returns
// sum_rows v[j].  Butterfly with halving: 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

static constexpr int TC5_THREADS = 192;

__global__ void __launch_bounds__(TC5_THREADS, 1) tc5_gather_conv_kernel(const __grid_constant__ Tc5ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: barriers first (small), then 1024-aligned tiles
    // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
    uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al);
    // full[8], empty[8], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_al + 192);
    const uint32_t err_flag = smem_u32(smem_al + 200);   // CTA-local "a wait timed out" flag
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem_al + 200) = 0u;
    uint8_t* tiles = smem_al + 1024;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t bytesA = 128u * p.KW * 2u;
    const uint32_t bytesB = (uint32_t)p.Ntile * p.KW * 2u;
    const uint32_t stageBytes = (bytesA + bytesB) * (uint32_t)p.tps;   // [A tap 0 .. tps-1][B tap 0 .. tps-1]
    const uint32_t tile_base = smem_u32(tiles);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
    auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };

    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * p.Ntile) tmem_cols <<= 1;
    // optional statistics accumulators behind the pipeline stages: one PRIVATE [2][NB][Nout] fp32 slot per
    // epilogue warp (plain read-modify-write by the owning lane, no atomics), summed and flushed once at the end
    float* statS = reinterpret_cast<float*>(tiles + (size_t)S * stageBytes);
    const int statN = p.NB * p.Nout;
    if (p.stat_sum != nullptr && p.statSmem)
        for (int i = threadIdx.x; i < 8 * statN; i += blockDim.x) statS[i] = 0.f;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapA[0]);
        if (p.nsrc > 1) tma_prefetch_desc(&p.mapA[1]);
        tma_prefetch_desc(&p.mapB);
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
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int spatialTiles = p.tilesW * p.tilesH * p.tilesD * p.tilesNB;
    const int totalTiles = spatialTiles * p.nTilesN;
    const int ntaps = p.tapD * p.tapH * p.tapW;
    const int Ctot = p.srcC[0] + (p.nsrc > 1 ? p.srcC[1] : 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop (warp-uniform control flow lets ptxas keep the pipeline state in uniform
        // registers); one elected lane issues.  Issuing from inside an `if (lane == 0)` region instead costs
        // ~230 cycles per TMA / MMA instruction (measured, profiles/r1_bottleneck_experiments.md).
        {
            int stage = 0;
            uint32_t phase = 0;
            const bool dbgT = (p.debug & 8) && blockIdx.x == 0;
            long long tWait = 0, tAll0 = clock64();
            for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
                // n-tile fastest so CTAs that share an activation tile run close in time (L2 reuse)
                uint32_t sp, nt, tiw, tih, tid, tib;
                fdivmod((uint32_t)tile, p.fdTilesN, sp, nt);
                fdivmod(sp, p.fdTilesW, sp, tiw);
                fdivmod(sp, p.fdTilesH, sp, tih);
                fdivmod(sp, p.fdTilesD, tib, tid);
                const int ow0 = tiw * p.tw, oh0 = tih * p.th, od0 = tid * p.td, nb0 = tib * p.tn;
                const int n0 = nt * p.Ntile;
                int t = 0;
                for (int kd = 0; kd < p.tapD; ++kd) {
                    const int iz = od0 * p.istrD + p.offD + kd;
                    for (int kh = 0; kh < p.tapH; ++kh) {
                        const int iy = oh0 * p.istrH + p.offH + kh;
                        for (int kw = 0; kw < p.tapW; kw += p.tps, t += p.tps) {
                            const int ix = ow0 * p.istrW + p.offW + kw;
                            int cbase = 0;
                            for (int s = 0; s < p.nsrc; ++s) {
                                for (int c = 0; c < p.srcC[s]; c += p.KW) {
                                    const long long w0 = dbgT ? clock64() : 0;
                                    mbar_wait(empty_bar(stage), phase ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
                                    if (dbgT) tWait += clock64() - w0;
                                    const uint32_t dstA = tile_base + stage * stageBytes;
                                    const uint32_t dstB = dstA + bytesA * p.tps;
                                    if (elect_one()) {
                                        if ((p.debug & 7) == 2) {
                                            mbar_arrive(full_bar(stage));
                                        } else {
                                            mbar_expect_tx(full_bar(stage), stageBytes);
                                            for (int j = 0; j < p.tps; ++j)
                                                tma_load_5d(dstA + j * bytesA, &p.mapA[s], full_bar(stage), c, ix + j, iy, iz, nb0);
                                            tma_load_3d(dstB, &p.mapB, full_bar(stage), cbase + c, n0, t);   // box depth = tps taps
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
            if (dbgT && lane == 0) { g_dbg[0] = (unsigned long long)tWait; g_dbg[1] = (unsigned long long)(clock64() - tAll0); }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        {
            const bool dbgT = (p.debug & 8) && blockIdx.x == 0;
            long long tWaitFull = 0, tWaitAcc = 0, tAll0 = clock64();
            const uint32_t idesc = make_idesc_bf16(128, p.Ntile, 0, 0);
            const uint32_t lay = swizzle_layout_code(p.KW * 2);
            const uint32_t sbo = 8u * p.KW * 2u;  // 8 rows of one swizzle span
            const int kPerStep = p.KW / 16;
            const int stepsPerTile = (ntaps / p.tps) * (Ctot / p.KW);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
                long long w0 = dbgT ? clock64() : 0;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, DEVERR_WAIT_TMEM_EMPTY, err_flag);
                if (dbgT) tWaitAcc += clock64() - w0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Ntile);
                for (int ks = 0; ks < stepsPerTile; ++ks) {
                    w0 = dbgT ? clock64() : 0;
                    mbar_wait(full_bar(stage), phase, DEVERR_WAIT_FULL, err_flag);
                    if (dbgT) tWaitFull += clock64() - w0;
                    tc_fence_after();
                    const uint32_t aAddr = tile_base + stage * stageBytes;
                    const uint32_t bAddr = aAddr + bytesA * p.tps;
                    if (elect_one()) {
                        for (int j = 0; j < p.tps && (p.debug & 7) != 1; ++j) {
                            for (int k = 0; k < kPerStep; ++k) {
                                const uint64_t da = make_smem_desc(aAddr + j * bytesA + k * 32u, 16u, sbo, lay);
                                const uint64_t db = make_smem_desc(bAddr + j * bytesB + k * 32u, 16u, sbo, lay);
                                umma_bf16(d_tmem, da, db, idesc, (ks | j | k) ? 1u : 0u);
                            }
                        }
                        umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (dbgT && lane == 0) {
                g_dbg[2] = (unsigned long long)tWaitFull; g_dbg[3] = (unsigned long long)tWaitAcc;
                g_dbg[4] = (unsigned long long)(clock64() - tAll0);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int iw = row % p.tw;
        const int ih = (row / p.tw) % p.th;
        const int id = (row / (p.tw * p.th)) % p.td;
        const int in = row / (p.tw * p.th * p.td);   // once per kernel
        const bool warpUniformSample = ((p.tw * p.th * p.td) & 31) == 0;
        const bool dbgT = (p.debug & 8) && blockIdx.x == 0 && warp == 2 && lane == 0;
        long long tWaitE = 0, tAllE0 = clock64();
        int nTilesE = 0;
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
            uint32_t sp, nt, tiw, tih, tid, tib;
            fdivmod((uint32_t)tile, p.fdTilesN, sp, nt);
            fdivmod(sp, p.fdTilesW, sp, tiw);
            fdivmod(sp, p.fdTilesH, sp, tih);
            fdivmod(sp, p.fdTilesD, tib, tid);
            const int ow = (int)tiw * p.tw + iw, oh = (int)tih * p.th + ih, od = (int)tid * p.td + id, nb = (int)tib * p.tn + in;
            const bool valid = (ow < p.OW) && (oh < p.OH) && (od < p.OD) && (nb < p.NB);
            const int n0 = (int)nt * p.Ntile;

            const long long we0 = dbgT ? clock64() : 0;
            mbar_wait(tfull_bar(acc), acc_phase, DEVERR_WAIT_TMEM_FULL, err_flag);
            if (dbgT) { tWaitE += clock64() - we0; ++nTilesE; }
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.Ntile);
            for (int cg = 0; cg < p.Ntile; cg += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_addr + cg, v);
                tmem_ld_wait();
                const int col0 = n0 + cg;
                if (col0 < p.Nout) {
                    if (p.stat_sum != nullptr && (p.debug & 7) != 3) {
                        if (warpUniformSample) {
                            // all 32 rows of this warp belong to sample nb: butterfly column sums, lane j ends up
                            // with column col0 + j
                            float a[32], b[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float x = valid ? __uint_as_float(v[j]) : 0.f;
                                a[j] = x;
                                b[j] = x * x;
                            }
                            const float s1 = warp_colsum32(a, lane);
                            const float s2 = warp_colsum32(b, lane);
                            if (col0 + lane < p.Nout && nb < p.NB) {
                                const int idx = nb * p.Nout + col0 + lane;
                                if (p.statSmem) {
                                    float* slot = statS + (size_t)quad * 2 * statN + idx;
                                    slot[0] += s1;
                                    slot[statN] += s2;
                                } else {
                                    atomicAdd(p.stat_sum + idx, s1);
                                    atomicAdd(p.stat_sq + idx, s2);
                                }
                            }
                        } else if (valid) {
                            // batch-folded tiles of tiny grids (< 32 voxels per sample): rows of a warp may belong to
                            // different samples
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (col0 + j < p.Nout) {
                                    const float x = __uint_as_float(v[j]);
                                    const int idx = nb * p.Nout + col0 + j;
                                    atomicAdd(p.stat_sum + idx, x);
                                    atomicAdd(p.stat_sq + idx, x * x);
                                }
                            }
                        }
                    }
                    if (valid && (p.debug & 7) != 4) {
                        int fd, fh, fw, ch0;
                        if (p.mode == 1) {
                            const int par = col0 / p.psC;
                            ch0 = col0 - par * p.psC;
                            const int pw = par % p.psW;
                            const int ph = (par / p.psW) % p.psH;
                            const int pd = par / (p.psW * p.psH);
                            fd = od * p.ostrD + pd; fh = oh * p.ostrH + ph; fw = ow * p.ostrW + pw;
                        } else {
                            ch0 = col0;
                            fd = od * p.ostrD + p.ooffD; fh = oh * p.ostrH + p.ooffH; fw = ow * p.ostrW + p.ooffW;
                        }
                        const size_t vox = (((size_t)nb * p.FD + fd) * p.FH + fh) * p.FW + fw;
                        int lim;      // channels available from ch0 in the chosen destination
                        size_t eoff;  // element offset into it
                        void* base;
                        if (ch0 < p.outC0) { base = p.out0; eoff = vox * p.outC0 + ch0; lim = p.outC0 - ch0; }
                        else { base = p.out1; eoff = vox * p.outC1 + (ch0 - p.outC0); lim = p.outC1 - (ch0 - p.outC0); }
                        if (p.mode == 1) lim = min(lim, p.psC - ch0);
                        if (p.outF32 == 1) {
                            float* dst = reinterpret_cast<float*>(base) + eoff;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (q * 4 < lim)
                                    *reinterpret_cast<uint4*>(dst + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                            }
                        } else {
                            bf16* dst = reinterpret_cast<bf16*>(base) + eoff;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q * 8 < lim) {
                                    uint4 o;
                                    const bool h = p.outF32 == 2;
                                    o.x = pack16(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]), h);
                                    o.y = pack16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]), h);
                                    o.z = pack16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]), h);
                                    o.w = pack16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]), h);
                                    *reinterpret_cast<uint4*>(dst + q * 8) = o;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (dbgT) {
            g_dbg[5] = (unsigned long long)tWaitE; g_dbg[6] = (unsigned long long)(clock64() - tAllE0);
            g_dbg[7] = (unsigned long long)nTilesE;
        }
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
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace rb
