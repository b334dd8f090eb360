// tcgen05 / TMEM / TMA weight-gradient kernel for sm_100a.
//
//   dW[tap][a][b] += sum_m P[m][a] * Q[gather(m, tap)][b]
//
// GEMM view: M = a (channels of P), N = b (channels of Q, one or two concatenated sources), K = voxels.
// Both operands are channels-last, i.e. "MN-major" (the M / N index is the contiguous one), which
// tcgen05.mma consumes directly: a TMA box [channels x 64 voxels] lands as 64 rows (K) of one swizzle
// span (channels), exactly the canonical MN-major layout
//     ((T, span/T, atoms), (8, k)) : ((1, T, LBO), (span, SBO))     T = 8 bf16 per 16 bytes
// with SBO = 8 rows * span bytes and LBO = the distance between channel atoms (one TMA box each).
//
// Work item = (a tile of 128, b tile of <= 256, tap, voxel split); the K loop walks 64-voxel spatial
// boxes (so the tap shift and the zero padding are TMA coordinates / out-of-bounds fill, as in the
// forward kernel).  Accumulators sit in TMEM (double buffered when 2 * N <= 512 columns); the epilogue
// adds the fp32 tile into dW with red.global (split-K over voxels and taps share nothing else).
//
// Replaces the weight half of aten::convolution_backward for Conv3d / ConvTranspose3d
// (reference: train.py:224 backward of builders/simple_conv_blocks.py:43-51, builders/decoder.py:110-113).
#pragma once
#include "common.cuh"

namespace rb {

struct Tc5WgradParams {
    CUtensorMap mapP;      // rank 5 (C, W, H, D, N) over the P tensor, box (aw, cw, ch, cd, cn)
    CUtensorMap mapQ[2];   // rank 5 over the Q tensor(s), box (bw, strided extents...)
    int PC, QC[2], nq;
    int aw, bw;            // channel atom widths (16 / 32 / 64) == swizzle span / 2
    int aAtoms;            // 128 / aw
    int bn;                // N tile (multiple of bw, <= 256)
    int aTiles, bTiles;
    int tapD, tapH, tapW, offD, offH, offW, istrD, istrH, istrW;
    int cw, ch, cd, cn;    // voxel chunk box, product 64
    int chunksW, chunksH, chunksD, chunksN;
    int splits, chunksPerSplit;
    int stages;
    int accBufs;           // 1 or 2 TMEM accumulator buffers
    float* dw;             // [taps][PC][QCtot]
    // Narrow layers (QCtot in {32, 64}): "tap stacking".  The roles are swapped: the 128 MMA rows are
    // tpi = 128 / QCtot taps x QCtot channels of Q (one TMA box per tap and channel atom), the N side is P
    // (N = PC <= 256).  One P box then serves tpi taps and no MMA row is padding.
    int swap;
    int tpi;               // taps per item (swap mode)
    int tapGroups;         // ceil(ntaps / tpi)
    FastDiv fdChunksW, fdChunksH, fdChunksD;   // chunk index decode without integer division
    int kbox;              // voxels per pipeline stage (64 or 128)
};

static constexpr int TW5_THREADS = 192;
static constexpr int TW5_KBOX = 64;

__global__ void __launch_bounds__(TW5_THREADS, 1) tc5_wgrad_kernel(const __grid_constant__ Tc5WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_al + 192);
    const uint32_t err_flag = smem_u32(smem_al + 200);   // CTA-local "a wait timed out" flag
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem_al + 200) = 0u;
    uint8_t* tiles = smem_al + 1024;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t atomA = (uint32_t)p.kbox * p.aw * 2u;   // bytes of one A atom (kbox rows x span)
    const uint32_t atomB = (uint32_t)p.kbox * p.bw * 2u;
    const int bAtoms = p.bn / p.bw;   // swap mode: aw = Q atom width, bw = P atom width, bn = PC
    const uint32_t bytesA = atomA * p.aAtoms;
    const uint32_t bytesB = atomB * bAtoms;
    const uint32_t stageBytes = bytesA + bytesB;
    const uint32_t tile_base = smem_u32(tiles);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
    auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.accBufs * p.bn)) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapP);
        tma_prefetch_desc(&p.mapQ[0]);
        if (p.nq > 1) tma_prefetch_desc(&p.mapQ[1]);
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

    const int ntaps = p.tapD * p.tapH * p.tapW;
    const int nChunks = p.chunksW * p.chunksH * p.chunksD * p.chunksN;
    const int tapSlots = p.swap ? p.tapGroups : ntaps;
    const int totalItems = p.aTiles * p.bTiles * tapSlots * p.splits;
    const int QCtot = p.QC[0] + (p.nq > 1 ? p.QC[1] : 0);
    const int qAtomsPerTap = p.swap ? QCtot / p.aw : 0;

    // item -> (split, tap, bt, at); a-tile fastest so neighbouring CTAs share the Q boxes in L2
    auto decode = [&](int item, int& at, int& bt, int& tap, int& sp) {
        at = item % p.aTiles; item /= p.aTiles;
        bt = item % p.bTiles; item /= p.bTiles;
        tap = item % tapSlots; item /= tapSlots;   // swap mode: tap group index
        sp = item;
    };

    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
                int at, bt, tap, sp;
                decode(item, at, bt, tap, sp);
                const int c0 = sp * p.chunksPerSplit;
                const int c1 = min(nChunks, c0 + p.chunksPerSplit);
                // per-item atom table (once per item: divisions are fine here, not per chunk)
                int atX[8], atY[8], atZ[8], atC[8], atSrc[8];
                const int nA = p.aAtoms;
                if (!p.swap) {
                    const int kw = tap % p.tapW, kh = (tap / p.tapW) % p.tapH, kd = tap / (p.tapW * p.tapH);
                    for (int j = 0; j < 8; ++j) {     // B side (Q) atoms
                        const int cb = bt * p.bn + j * p.bw;
                        const bool second = p.nq > 1 && cb >= p.QC[0];
                        atX[j] = p.offW + kw; atY[j] = p.offH + kh; atZ[j] = p.offD + kd;
                        atC[j] = second ? cb - p.QC[0] : cb;
                        atSrc[j] = second ? 1 : 0;
                    }
                } else {
                    for (int j = 0; j < 8; ++j) {     // A side (Q) atoms: tap-major, then channel atoms
                        int tt = tap * p.tpi + j / qAtomsPerTap;
                        if (tt >= ntaps) tt = ntaps - 1;   // rows of taps beyond the kernel are ignored by the epilogue
                        const int tw_ = tt % p.tapW, th_ = (tt / p.tapW) % p.tapH, td_ = tt / (p.tapW * p.tapH);
                        const int cb = (j % qAtomsPerTap) * p.aw;
                        const bool second = p.nq > 1 && cb >= p.QC[0];
                        atX[j] = p.offW + tw_; atY[j] = p.offH + th_; atZ[j] = p.offD + td_;
                        atC[j] = second ? cb - p.QC[0] : cb;
                        atSrc[j] = second ? 1 : 0;
                    }
                }
                for (int c = c0; c < c1; ++c) {
                    uint32_t t, iw, ih, id, in;
                    fdivmod((uint32_t)c, p.fdChunksW, t, iw);
                    fdivmod(t, p.fdChunksH, t, ih);
                    fdivmod(t, p.fdChunksD, in, id);
                    const int gw0 = iw * p.cw, gh0 = ih * p.ch, gd0 = id * p.cd, n0 = in * p.cn;
                    const int qx0 = gw0 * p.istrW, qy0 = gh0 * p.istrH, qz0 = gd0 * p.istrD;
                    mbar_wait(empty_bar(stage), phase ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
                    const uint32_t dstA = tile_base + stage * stageBytes;
                    const uint32_t dstB = dstA + bytesA;
                    if (elect_one()) {
                    mbar_expect_tx(full_bar(stage), stageBytes);
                    if (!p.swap) {
                        for (int j = 0; j < nA; ++j)
                            tma_load_5d(dstA + j * atomA, &p.mapP, full_bar(stage), at * 128 + j * p.aw, gw0, gh0, gd0, n0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < bAtoms)
                                tma_load_5d(dstB + j * atomB, &p.mapQ[atSrc[j]], full_bar(stage), atC[j], qx0 + atX[j],
                                            qy0 + atY[j], qz0 + atZ[j], n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nA)
                                tma_load_5d(dstA + j * atomA, &p.mapQ[atSrc[j]], full_bar(stage), atC[j], qx0 + atX[j],
                                            qy0 + atY[j], qz0 + atZ[j], n0);
                        for (int j = 0; j < bAtoms; ++j)
                            tma_load_5d(dstB + j * atomB, &p.mapP, full_bar(stage), j * p.bw, gw0, gh0, gd0, n0);
                    }
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        {
            const uint32_t idesc = make_idesc_bf16(128, p.bn, 1, 1);   // both operands MN-major
            const uint32_t layA = swizzle_layout_code(p.aw * 2), layB = swizzle_layout_code(p.bw * 2);
            const uint32_t sboA = 8u * p.aw * 2u, sboB = 8u * p.bw * 2u;   // 8 K rows of one span
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
                int at, bt, tap, sp;
                decode(item, at, bt, tap, sp);
                const int c0 = sp * p.chunksPerSplit;
                const int c1 = min(nChunks, c0 + p.chunksPerSplit);
                if (c0 >= c1) continue;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, DEVERR_WAIT_TMEM_EMPTY, err_flag);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.bn);
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(full_bar(stage), phase, DEVERR_WAIT_FULL, err_flag);
                    tc_fence_after();
                    const uint32_t aAddr = tile_base + stage * stageBytes;
                    const uint32_t bAddr = aAddr + bytesA;
                    if (elect_one()) {
                        for (int k = 0; k < p.kbox / 16; ++k) {
                            // 16 voxels = two 8-row groups further down the K direction
                            const uint64_t da = make_smem_desc(aAddr + k * 2u * sboA, atomA, sboA, layA);
                            const uint64_t db = make_smem_desc(bAddr + k * 2u * sboB, atomB, sboB, layB);
                            umma_bf16(d_tmem, da, db, idesc, (c > c0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) umma_commit(tfull_bar(acc));
                __syncwarp();
                if (p.accBufs == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1u; } }
                else acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5): dW tile += TMEM =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            int at, bt, tap, sp;
            decode(item, at, bt, tap, sp);
            const int c0 = sp * p.chunksPerSplit;
            const int c1 = min(nChunks, c0 + p.chunksPerSplit);
            if (c0 >= c1) continue;
            mbar_wait(tfull_bar(acc), acc_phase, DEVERR_WAIT_TMEM_FULL, err_flag);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.bn);
            if (!p.swap) {
                const int a = at * 128 + row;
                float* drow = p.dw + ((size_t)tap * p.PC + a) * QCtot + (size_t)bt * p.bn;
                for (int cg = 0; cg < p.bn; cg += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_addr + cg, v);
                    tmem_ld_wait();
                    if (a < p.PC) {
                        if (p.splits == 1 && bt * p.bn + cg + 32 <= QCtot) {
                            // the only writer of these elements: plain 16-byte stores, no zero-fill needed
#pragma unroll
                            for (int q4 = 0; q4 < 8; ++q4)
                                *reinterpret_cast<uint4*>(drow + cg + q4 * 4) =
                                    make_uint4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
                        } else if (p.splits == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (bt * p.bn + cg + j < QCtot) drow[cg + j] = __uint_as_float(v[j]);
                        } else if (bt * p.bn + cg + 32 <= QCtot) {
                            // split-K partial tile: a lane owns 32 consecutive floats of one dW row, so vector reductions
                            // (16 bytes = one L2 sector per instruction and lane) instead of 32 scalar ones that each touch
                            // their own sector - the scalar form ran at 70 % L2 throughput and 2 % tensor pipe (ncu)
#pragma unroll
                            for (int q4 = 0; q4 < 8; ++q4)
                                red_add_v4(drow + cg + q4 * 4, __uint_as_float(v[q4 * 4]), __uint_as_float(v[q4 * 4 + 1]),
                                           __uint_as_float(v[q4 * 4 + 2]), __uint_as_float(v[q4 * 4 + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int b = bt * p.bn + cg + j;
                                if (b < QCtot) atomicAdd(drow + cg + j, __uint_as_float(v[j]));
                            }
                        }
                    }
                }
            } else {
                // row = (tap within the group, Q channel b); column = P channel a
                const int tt = tap * p.tpi + row / QCtot;
                const int b = row % QCtot;
                float* dcol = p.dw + (size_t)tt * p.PC * QCtot + b;
                for (int cg = 0; cg < p.bn; cg += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_addr + cg, v);
                    tmem_ld_wait();
                    if (tt < ntaps) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int a = cg + j;
                            if (a < p.PC) atomicAdd(dcol + (size_t)a * QCtot, __uint_as_float(v[j]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (p.accBufs == 2) { if (++acc == 2) { acc = 0; acc_phase ^= 1u; } }
            else acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace rb
