// tcgen05 weight gradient with TWO-SIDED TAP STACKING, for stride-1 convolutions.
//
//   dW[kd,kh,kw][a][b] = sum_g P[g][a] * Q[g + off + (kd,kh,kw)][b]
//                      = sum_u P[u - (offD+kd, offH+kh, 0)][a] * Q[u + (0, 0, offW+kw)][b]        (u = g + (offD+kd, offH+kh, 0))
//
// Shifting P by the (kd, kh) part of the tap and Q by the kw part lets one MMA cover tapW x (tapD*tapH) taps:
//   M rows    = (kw, b)        : tapW shifted TMA boxes of Q      (3 x 32 channels = 96 rows, 3 x 64 = 192 -> 2 groups)
//   N columns = ((kd,kh), a)   : tapD*tapH shifted TMA boxes of P (9 x 32 channels = 288 columns -> groups of <= 256)
//   K         = voxels u
// For the 32-channel full-resolution layers this turns 27 MMAs of 128 x 32 (12 % of the MMA rate, one P/Q box pair
// each) into two MMAs of 128 x 256 / 128 x 32 per K step with 12 boxes instead of 35 per voxel chunk.
// Zero padding of both operands is TMA out-of-bounds fill; out-of-grid u positions contribute nothing because the
// corresponding Q rows are padding.
//
// Operand layout, pipeline and epilogue as in wgrad_tc5.cuh (MN-major operands straight from the channels-last
// tensors, fp32 accumulators in TMEM, red.global into dW[tap][a][b]).
#pragma once
#include "common.cuh"

namespace rb {

struct Tc5Wgrad2Params {
    CUtensorMap mapP;      // rank 5 (C, W, H, D, N) over P, box (pw, cw, ch, cd, cn)
    CUtensorMap mapQ[2];   // rank 5 over the Q tensor(s), box (qw, cw, ch, cd, cn)
    int PC, QC[2], nq;
    int pw, qw;            // channel atom widths of P and Q (32 or 64)
    int mAtomsTotal;       // tapW * QCtot / qw
    int nAtomsTotal;       // tapD * tapH * PC / pw
    int mPerGroup;         // 128 / qw
    int nPerGroup;         // 256 / pw
    int mGroups, nGroups;
    int merged;            // all N atoms (<= 512 columns) in one item: the Q boxes are loaded once per voxel chunk and two
                           // MMAs (256 + rest columns) share them; single TMEM accumulator of nAtomsTotal*pw columns
    int tapD, tapH, tapW, offD, offH, offW;
    int kbox;              // voxels per stage (64 or 128)
    int cw, ch, cd, cn;    // voxel chunk box, product kbox
    int chunksW, chunksH, chunksD, chunksN;
    FastDiv fdChunksW, fdChunksH, fdChunksD, fdMGroups, fdNGroups;
    int splits, chunksPerSplit;
    int stages;
    float* dw;             // [taps][PC][QCtot]
};

static constexpr int TW52_THREADS = 192;

// MERGED: separate instantiation, so that the two-group path keeps its 8-entry atom tables and single MMA per K step
template <bool MERGED>
__global__ void __launch_bounds__(TW52_THREADS, 1) tc5_wgrad2_kernel(const __grid_constant__ Tc5Wgrad2Params p) {
    constexpr int NBMAX = MERGED ? 16 : 8;
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
    const uint32_t atomA = (uint32_t)p.kbox * p.qw * 2u;   // one Q atom: kbox voxel rows x one swizzle span
    const uint32_t atomB = (uint32_t)p.kbox * p.pw * 2u;
    const uint32_t bytesA = atomA * p.mPerGroup;            // 128 MMA rows
    const uint32_t bytesB = atomB * p.nPerGroup;            // up to 256 MMA columns
    const uint32_t stageBytes = bytesA + bytesB;
    const uint32_t tile_base = smem_u32(tiles);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
    auto tfull_bar = [&](int a) { return bar_base + 128u + 8u * a; };
    auto tempty_bar = [&](int a) { return bar_base + 144u + 8u * a; };

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
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nChunks = p.chunksW * p.chunksH * p.chunksD * p.chunksN;
    const int totalItems = p.mGroups * p.nGroups * p.splits;
    const int QCtot = p.QC[0] + (p.nq > 1 ? p.QC[1] : 0);
    const int qAtomsPerTap = QCtot / p.qw;    // M atom j -> kw = j / qAtomsPerTap
    const int pAtomsPerTap = p.PC / p.pw;     // N atom j -> (kd,kh) = j / pAtomsPerTap

    auto decode = [&](int item, uint32_t& mg, uint32_t& ng, uint32_t& sp) {
        uint32_t t;
        fdivmod((uint32_t)item, p.fdMGroups, t, mg);
        fdivmod(t, p.fdNGroups, sp, ng);
    };
    auto group_counts = [&](uint32_t mg, uint32_t ng, int& nA, int& nB) {
        nA = min(p.mPerGroup, p.mAtomsTotal - (int)mg * p.mPerGroup);
        nB = min(p.nPerGroup, p.nAtomsTotal - (int)ng * p.nPerGroup);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            uint32_t mg, ng, sp;
            decode(item, mg, ng, sp);
            int nA, nB;
            group_counts(mg, ng, nA, nB);
            const int c0 = sp * p.chunksPerSplit;
            const int c1 = min(nChunks, c0 + p.chunksPerSplit);
            // per-item atom tables (divisions once per item)
            int aX[8], aC[8], aSrc[8], bY[NBMAX], bZ[NBMAX], bC[NBMAX];
#pragma unroll
            for (int j = 0; j < NBMAX; ++j) {
                const int ja = (int)mg * p.mPerGroup + (j & 7);
                const int kw = ja / qAtomsPerTap;
                const int cb = (ja - kw * qAtomsPerTap) * p.qw;
                const bool second = p.nq > 1 && cb >= p.QC[0];
                aX[j & 7] = p.offW + kw;
                aC[j & 7] = second ? cb - p.QC[0] : cb;
                aSrc[j & 7] = second ? 1 : 0;
                const int jb = (int)ng * p.nPerGroup + j;
                const int s = jb / pAtomsPerTap;
                const int kd = s / p.tapH, kh = s - kd * p.tapH;
                bZ[j] = -(p.offD + kd);
                bY[j] = -(p.offH + kh);
                bC[j] = (jb - s * pAtomsPerTap) * p.pw;
            }
            const uint32_t txBytes = (uint32_t)nA * atomA + (uint32_t)nB * atomB;
            for (int c = c0; c < c1; ++c) {
                uint32_t t, iw, ih, id, in;
                fdivmod((uint32_t)c, p.fdChunksW, t, iw);
                fdivmod(t, p.fdChunksH, t, ih);
                fdivmod(t, p.fdChunksD, in, id);
                const int uw = iw * p.cw, uh = ih * p.ch, ud = id * p.cd, n0 = in * p.cn;
                mbar_wait(empty_bar(stage), phase ^ 1u, DEVERR_WAIT_EMPTY, err_flag);
                if (elect_one()) {
                    const uint32_t dstA = tile_base + stage * stageBytes;
                    const uint32_t dstB = dstA + bytesA;
                    mbar_expect_tx(full_bar(stage), txBytes);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < nA)
                            tma_load_5d(dstA + j * atomA, &p.mapQ[aSrc[j]], full_bar(stage), aC[j], uw + aX[j], uh, ud, n0);
#pragma unroll
                    for (int j = 0; j < NBMAX; ++j)
                        if (j < nB)
                            tma_load_5d(dstB + j * atomB, &p.mapP, full_bar(stage), bC[j], uw, uh + bY[j], ud + bZ[j], n0);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t layA = swizzle_layout_code(p.qw * 2), layB = swizzle_layout_code(p.pw * 2);
        const uint32_t sboA = 8u * p.qw * 2u, sboB = 8u * p.pw * 2u;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            uint32_t mg, ng, sp;
            decode(item, mg, ng, sp);
            int nA, nB;
            group_counts(mg, ng, nA, nB);
            const int c0 = sp * p.chunksPerSplit;
            const int c1 = min(nChunks, c0 + p.chunksPerSplit);
            if (c0 >= c1) continue;
            const int ncol = nB * p.pw;
            const int n1 = ncol > 256 ? 256 : ncol, n2 = ncol - n1;
            const uint32_t idesc = make_idesc_bf16(128, n1, 1, 1);   // both operands MN-major
            const uint32_t idesc2 = make_idesc_bf16(128, n2 > 0 ? n2 : 8, 1, 1);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u, DEVERR_WAIT_TMEM_EMPTY, err_flag);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
            for (int c = c0; c < c1; ++c) {
                mbar_wait(full_bar(stage), phase, DEVERR_WAIT_FULL, err_flag);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t aAddr = tile_base + stage * stageBytes;
                    const uint32_t bAddr = aAddr + bytesA;
                    for (int k = 0; k < p.kbox / 16; ++k) {
                        const uint64_t da = make_smem_desc(aAddr + k * 2u * sboA, atomA, sboA, layA);
                        const uint64_t db = make_smem_desc(bAddr + k * 2u * sboB, atomB, sboB, layB);
                        umma_bf16(d_tmem, da, db, idesc, (c > c0 || k > 0) ? 1u : 0u);
                        if (MERGED && n2 > 0) {
                            const uint64_t db2 = make_smem_desc(bAddr + (uint32_t)(256 / p.pw) * atomB + k * 2u * sboB, atomB, sboB, layB);
                            umma_bf16(d_tmem + 256u, da, db2, idesc2, (c > c0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));
            __syncwarp();
            if (MERGED) acc_phase ^= 1u;                       // one accumulator: same barrier pair every item
            else if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ===================== epilogue: row = (kw, b), column = ((kd,kh), a) =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
            uint32_t mg, ng, sp;
            decode(item, mg, ng, sp);
            int nA, nB;
            group_counts(mg, ng, nA, nB);
            const int c0 = sp * p.chunksPerSplit;
            const int c1 = min(nChunks, c0 + p.chunksPerSplit);
            if (c0 >= c1) continue;
            // this thread's row: M atom, tap kw and Q channel b
            const int ja = row / p.qw;
            const int jaG = (int)mg * p.mPerGroup + ja;
            const int kw = jaG / qAtomsPerTap;
            const int b = (jaG - kw * qAtomsPerTap) * p.qw + (row - ja * p.qw);
            const bool rowValid = ja < nA;
            mbar_wait(tfull_bar(acc), acc_phase, DEVERR_WAIT_TMEM_FULL, err_flag);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
            const int ncols = nB * p.pw;
            for (int cg = 0; cg < ncols; cg += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_addr + cg, v);
                tmem_ld_wait();
                if (rowValid) {
                    // the 32 columns of a group share one N atom when pw >= 32: one (kd,kh), consecutive a
                    const int jb = (int)ng * p.nPerGroup + cg / p.pw;
                    const int s = jb / pAtomsPerTap;
                    const int a0 = (jb - s * pAtomsPerTap) * p.pw + (cg % p.pw);
                    const int t = s * p.tapW + kw;      // tap index (kd*tapH + kh)*tapW + kw
                    float* d = p.dw + ((size_t)t * p.PC + a0) * QCtot + b;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(d + (size_t)j * QCtot, __uint_as_float(v[j]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (MERGED) acc_phase ^= 1u;
            else if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
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
