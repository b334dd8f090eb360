// Shape-generic implicit-GEMM gather convolution and weight-gradient kernels on the legacy
// warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate).  They take any channel
// count that is a multiple of 8 and any grid, and serve (a) shapes the tcgen05 kernels do not
// cover (Cout < 32, C % 16 != 0, bottleneck widths) and (b) as the on-device cross-check the
// GPU tests run the tcgen05 kernels against.  Same operand conventions as conv_tc5.cuh.
#pragma once
#include "common.cuh"

namespace rb {

struct GConvParams {
    const bf16* src[2];
    int srcC[2];
    int nsrc;
    int ID, IH, IW, NB;  // input grid
    const bf16* w;       // [taps][Nout][Ctot]
    int tapD, tapH, tapW, offD, offH, offW, istrD, istrH, istrW;
    int OD, OH, OW;  // output class grid
    int Nout;
    int mode, ostrD, ostrH, ostrW, ooffD, ooffH, ooffW, FD, FH, FW;
    void* out0;
    void* out1;
    int outC0, outC1, psC, psD, psH, psW;
    int outF32;  // destination element type: 0 bf16, 1 fp32, 2 fp16
    int splitK;  // > 1: blockIdx.z takes a contiguous slice of the (tap, k-chunk) loop and adds fp32
    float* ws;   //      partials into ws[m][Nout] (zeroed by the host); gather_finish_kernel stores
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

static constexpr int GC_BM = 128, GC_BN = 64, GC_BK = 32, GC_PITCH = 40, GC_THREADS = 256;

__global__ void __launch_bounds__(GC_THREADS) gather_conv_mma_kernel(const GConvParams p) {
    __shared__ __align__(16) bf16 As[GC_BM * GC_PITCH];
    __shared__ __align__(16) bf16 Bs[GC_BN * GC_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp & 3, warp_n = warp >> 2;
    const long long Mtot = (long long)p.NB * p.OD * p.OH * p.OW;
    const long long m0 = (long long)blockIdx.x * GC_BM;
    const int n0 = blockIdx.y * GC_BN;
    const int Ctot = p.srcC[0] + (p.nsrc > 1 ? p.srcC[1] : 0);
    const int ntaps = p.tapD * p.tapH * p.tapW;
    const int kchunks = (Ctot + GC_BK - 1) / GC_BK;
    const int nItAll = ntaps * kchunks;
    int itBeg = 0, nIt = nItAll;
    if (p.splitK > 1) {
        const int per = (nItAll + p.splitK - 1) / p.splitK;
        itBeg = blockIdx.z * per;
        nIt = min(nItAll, itBeg + per);
        if (itBeg >= nIt) return;
    }

    // the two A rows this thread stages, decomposed once
    int rnb[2], rod[2], roh[2], row_[2];
    bool rvalid[2];
    const int kvec = tid & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 2) + i * 64;
        long long m = m0 + r;
        rvalid[i] = m < Mtot;
        if (!rvalid[i]) m = 0;
        row_[i] = (int)(m % p.OW); m /= p.OW;
        roh[i] = (int)(m % p.OH); m /= p.OH;
        rod[i] = (int)(m % p.OD); m /= p.OD;
        rnb[i] = (int)m;
    }
    const int bn = tid >> 2;  // B row staged by this thread

    uint4 ra[2], rb_;
    auto prefetch = [&](int it) {
        const int t = it / kchunks, kc = it - t * kchunks;
        const int kw = t % p.tapW, kh = (t / p.tapW) % p.tapH, kd = t / (p.tapW * p.tapH);
        const int c = kc * GC_BK + kvec * 8;
        const bf16* sp = p.src[0];
        int cs = c, sC = p.srcC[0];
        if (p.nsrc > 1 && c >= p.srcC[0]) { sp = p.src[1]; cs = c - p.srcC[0]; sC = p.srcC[1]; }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int iz = rod[i] * p.istrD + p.offD + kd;
            const int iy = roh[i] * p.istrH + p.offH + kh;
            const int ix = row_[i] * p.istrW + p.offW + kw;
            const bool ok = rvalid[i] && c < Ctot && iz >= 0 && iz < p.ID && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
            ra[i] = make_uint4(0, 0, 0, 0);
            if (ok) {
                const size_t vox = (((size_t)rnb[i] * p.ID + iz) * p.IH + iy) * p.IW + ix;
                ra[i] = __ldg(reinterpret_cast<const uint4*>(sp + vox * sC + cs));
            }
        }
        rb_ = make_uint4(0, 0, 0, 0);
        if (n0 + bn < p.Nout && c < Ctot)
            rb_ = __ldg(reinterpret_cast<const uint4*>(p.w + ((size_t)t * p.Nout + n0 + bn) * Ctot + c));
    };

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    prefetch(itBeg);
    for (int it = itBeg; it < nIt; ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            *reinterpret_cast<uint4*>(&As[((tid >> 2) + i * 64) * GC_PITCH + kvec * 8]) = ra[i];
        *reinterpret_cast<uint4*>(&Bs[bn * GC_PITCH + kvec * 8]) = rb_;
        __syncthreads();
        if (it + 1 < nIt) prefetch(it + 1);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = warp_m * 32 + mi * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
                const int c = ks * 16 + (lane >> 4) * 8;
                ldsm_x4(smem_u32(&As[r * GC_PITCH + c]), af[mi][0], af[mi][1], af[mi][2], af[mi][3]);
            }
            uint32_t bfr[4][2];
#pragma unroll
            for (int nj = 0; nj < 2; ++nj) {
                const int r = warp_n * 32 + nj * 16 + (lane >> 4) * 8 + (lane & 7);
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(smem_u32(&Bs[r * GC_PITCH + c]), bfr[nj * 2][0], bfr[nj * 2][1], bfr[nj * 2 + 1][0],
                        bfr[nj * 2 + 1][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) mma_bf16_16816(acc[mi][ni], af[mi], bfr[ni][0], bfr[ni][1]);
        }
        __syncthreads();
    }

    // epilogue
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = warp_m * 32 + mi * 16 + half * 8 + g;
            long long m = m0 + r;
            if (m >= Mtot) continue;
            if (p.splitK > 1) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int col = n0 + warp_n * 32 + ni * 8 + q * 2;
                    if (col >= p.Nout) continue;
                    float* d = p.ws + (size_t)m * p.Nout + col;
                    atomicAdd(d, acc[mi][ni][half * 2]);
                    atomicAdd(d + 1, acc[mi][ni][half * 2 + 1]);
                }
                continue;
            }
            const int ow = (int)(m % p.OW); m /= p.OW;
            const int oh = (int)(m % p.OH); m /= p.OH;
            const int od = (int)(m % p.OD); m /= p.OD;
            const int nb = (int)m;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int col = n0 + warp_n * 32 + ni * 8 + q * 2;
                if (col >= p.Nout) continue;
                int fd, fh, fw, ch;
                if (p.mode == 1) {
                    const int par = col / p.psC;
                    ch = col - par * p.psC;
                    const int pw = par % p.psW, ph = (par / p.psW) % p.psH, pd = par / (p.psW * p.psH);
                    fd = od * p.ostrD + pd; fh = oh * p.ostrH + ph; fw = ow * p.ostrW + pw;
                } else {
                    ch = col;
                    fd = od * p.ostrD + p.ooffD; fh = oh * p.ostrH + p.ooffH; fw = ow * p.ostrW + p.ooffW;
                }
                const size_t vox = (((size_t)nb * p.FD + fd) * p.FH + fh) * p.FW + fw;
                void* base = (ch < p.outC0) ? p.out0 : p.out1;
                const size_t eoff = (ch < p.outC0) ? vox * p.outC0 + ch : vox * p.outC1 + (ch - p.outC0);
                if (p.outF32 == 1)
                    *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + eoff) =
                        make_float2(acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1]);
                else
                    *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(base) + eoff) =
                        pack16(acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1], p.outF32 == 2);
            }
        }
    }
}

// Split-K finish: ws[m][Nout] fp32 -> bf16 destination(s) with the same addressing as above.
__global__ void __launch_bounds__(256) gather_finish_kernel(const GConvParams p) {
    const long long Mtot = (long long)p.NB * p.OD * p.OH * p.OW;
    const int half = p.Nout >> 1;
    const long long total = Mtot * half;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % half) * 2;
        long long m = i / half;
        const float2 v = *reinterpret_cast<const float2*>(p.ws + (size_t)m * p.Nout + col);
        const int ow = (int)(m % p.OW); m /= p.OW;
        const int oh = (int)(m % p.OH); m /= p.OH;
        const int od = (int)(m % p.OD); m /= p.OD;
        const int nb = (int)m;
        int fd, fh, fw, ch;
        if (p.mode == 1) {
            const int par = col / p.psC;
            ch = col - par * p.psC;
            const int pw = par % p.psW, ph = (par / p.psW) % p.psH, pd = par / (p.psW * p.psH);
            fd = od * p.ostrD + pd; fh = oh * p.ostrH + ph; fw = ow * p.ostrW + pw;
        } else {
            ch = col;
            fd = od * p.ostrD + p.ooffD; fh = oh * p.ostrH + p.ooffH; fw = ow * p.ostrW + p.ooffW;
        }
        const size_t vox = (((size_t)nb * p.FD + fd) * p.FH + fh) * p.FW + fw;
        void* base = (ch < p.outC0) ? p.out0 : p.out1;
        const size_t eoff = (ch < p.outC0) ? vox * p.outC0 + ch : vox * p.outC1 + (ch - p.outC0);
        if (p.outF32 == 1)
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + eoff) = v;
        else
            *reinterpret_cast<uint32_t*>(reinterpret_cast<bf16*>(base) + eoff) = pack16(v.x, v.y, p.outF32 == 2);
    }
}

// Finish pass of the tap-split tcgen05 convolution (conv_tc5t.cuh, deep 4^3 / 8^3 / 16^3 layers):
//   out[m][c] = sum_slices ws[slice][tile(m)][column(m)][c]   (bf16 or fp32 destination, plain stride-1 convolution: voxel m
//   of the output grid is voxel m of the destination), plus the InstanceNorm statistics sum / sum of squares per
//   (sample, channel) of the stored values - the statistics the unsplit kernels take in their epilogue.
// A warp owns 32 consecutive channels (lane = channel: 128-byte rows of the workspace) and `vpw` consecutive voxels of
// one sample, all slice loads of a voxel in flight at once; the 8 warps of a block cover 8 consecutive voxel runs of the
// same (sample, channel group), so the statistics are combined in shared memory and cost 64 atomics per block.
struct SplitFinishParams {
    const float* ws;      // [slices][voxel tiles][256][Nout]
    long long sliceStride;
    int slices;
    int S;                // voxels per sample
    int NB, Nout;
    int OW, OH, OD;
    int lw, lh, ld;       // log2 tile box extents; tn = 256 >> (lw + lh + ld) samples per tile
    int tilesW, tilesH, tilesD;
    void* out0;
    void* out1;
    int outC0, outC1, outF32;
    float* stat_sum;      // [NB][Nout] or null
    float* stat_sq;
    int vpw;              // voxels per warp
    int runsPerSample;    // ceil(S / (warps per block * vpw))
};

__global__ void __launch_bounds__(256) split_finish_kernel(const SplitFinishParams p) {
    // lane = (voxel sub-index lane >> 3, channel quad lane & 7): one warp instruction moves 4 voxels x 32 channels (512 B)
    __shared__ float red[2][8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;                   // 8 warps, or 2 for the tiny 4^3 planes (more blocks: latency bound)
    const int cgroups = p.Nout >> 5;
    int b = blockIdx.x;
    const int cgI = b % cgroups; b /= cgroups;
    const int run = b % p.runsPerSample;
    const int nb = b / p.runsPerSample;
    const int c = cgI * 32 + (lane & 7) * 4;          // first of this lane's 4 channels
    const int vbeg = (run * nw + warp) * p.vpw;
    const int vend = min(p.S, vbeg + p.vpw);
    const bool first = c < p.outC0;
    const int cdst = first ? c : c - p.outC0;
    const int cpitch = first ? p.outC0 : p.outC1;
    void* const base = first ? p.out0 : p.out1;
    const int ls = p.lw + p.lh + p.ld;
    const int tib = nb >> (8 - ls), in = nb & ((256 >> ls) - 1);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int v = vbeg + (lane >> 3); v < vend; v += 4) {
        const int ow = v % p.OW, t = v / p.OW;
        const int oh = t % p.OH, od = t / p.OH;
        const int tiw = ow >> p.lw, tih = oh >> p.lh, tid = od >> p.ld;
        const int r = (ow & ((1 << p.lw) - 1)) | ((oh & ((1 << p.lh) - 1)) << p.lw) | ((od & ((1 << p.ld) - 1)) << (p.lw + p.lh)) | (in << ls);
        const size_t tileLin = (((size_t)tib * p.tilesD + tid) * p.tilesH + tih) * p.tilesW + tiw;
        const float4* src = reinterpret_cast<const float4*>(p.ws + (tileLin * 256 + r) * p.Nout + c);
        const size_t ss = (size_t)p.sliceStride >> 2;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        int sl = 0;
        for (; sl + 4 <= p.slices; sl += 4) {
            const float4 x0 = __ldcs(src + (size_t)sl * ss), x1 = __ldcs(src + (size_t)(sl + 1) * ss);
            const float4 x2 = __ldcs(src + (size_t)(sl + 2) * ss), x3 = __ldcs(src + (size_t)(sl + 3) * ss);
            a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
            a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
            a2.x += x2.x; a2.y += x2.y; a2.z += x2.z; a2.w += x2.w;
            a3.x += x3.x; a3.y += x3.y; a3.z += x3.z; a3.w += x3.w;
        }
        for (; sl < p.slices; ++sl) {
            const float4 x0 = __ldcs(src + (size_t)sl * ss);
            a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
        }
        float x[4] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                      (a0.w + a1.w) + (a2.w + a3.w)};
        const size_t m = (size_t)nb * p.S + v;
        if (p.outF32 == 1) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + m * cpitch + cdst) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(base) + m * cpitch + cdst) =
                make_uint2(pack16(x[0], x[1], p.outF32 == 2), pack16(x[2], x[3], p.outF32 == 2));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { s1[k] += x[k]; s2[k] = fmaf(x[k], x[k], s2[k]); }
    }
    if (p.stat_sum != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 8);  s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 8);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16); s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 16);
        }
        if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { red[0][warp][lane * 4 + k] = s1[k]; red[1][warp][lane * 4 + k] = s2[k]; }
        }
        __syncthreads();
        if (warp < 2) {
            float a = 0.f;
            for (int w = 0; w < nw; ++w) a += red[warp][w][lane];
            atomicAdd((warp == 0 ? p.stat_sum : p.stat_sq) + (size_t)nb * p.Nout + cgI * 32 + lane, a);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Weight gradient:  dW[tap][a][b] += sum_m P[m][a] * Q[gather(m, tap)][b]
//   conv wgrad : P = dy (a = Cout) on the output grid, Q = x (b = Cin, two sources allowed)
//   convT wgrad: P = x  (a = Cin) on the input grid,  Q = dy gathered at 2i+p (b = Cout)
// fp32 atomics into a zero-initialised [taps][A][B] buffer (split over the voxel dimension).
// ---------------------------------------------------------------------------------------
struct GWgradParams {
    const bf16* P; int PC;        // [NB, GD, GH, GW, PC]
    const bf16* Q[2]; int QC[2]; int nq;   // gathered operand(s), grid QD x QH x QW
    int NB, GD, GH, GW, QD, QH, QW;
    int tapD, tapH, tapW, offD, offH, offW, istrD, istrH, istrW;
    float* dw;                    // [taps][PC][QCtot]
    int mPerSplit;                // voxels handled by one blockIdx.z
};

static constexpr int GW_BA = 64, GW_BB = 64, GW_BK = 32, GW_PITCH = 72;

__global__ void __launch_bounds__(256) gather_wgrad_mma_kernel(const GWgradParams p) {
    __shared__ __align__(16) bf16 Ps[GW_BK * GW_PITCH];
    __shared__ __align__(16) bf16 Qs[GW_BK * GW_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_a = warp & 3, warp_b = warp >> 2;
    const int QCtot = p.QC[0] + (p.nq > 1 ? p.QC[1] : 0);
    const int tilesB = (QCtot + GW_BB - 1) / GW_BB;
    const int a0 = (blockIdx.x / tilesB) * GW_BA;
    const int b0 = (blockIdx.x % tilesB) * GW_BB;
    const int t = blockIdx.y;
    const int kw = t % p.tapW, kh = (t / p.tapW) % p.tapH, kd = t / (p.tapW * p.tapH);
    const long long Mtot = (long long)p.NB * p.GD * p.GH * p.GW;
    const long long mBeg = (long long)blockIdx.z * p.mPerSplit;
    const long long mEnd = min(Mtot, mBeg + p.mPerSplit);
    if (mBeg >= mEnd) return;

    const int lr = tid >> 3, lv = tid & 7;  // staged row (voxel within chunk) and 16-byte vector
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    uint4 rp, rq;
    auto prefetch = [&](long long mc) {
        const long long m = mc + lr;
        rp = make_uint4(0, 0, 0, 0);
        rq = make_uint4(0, 0, 0, 0);
        if (m < mEnd) {
            long long mm = m;
            const int gw = (int)(mm % p.GW); mm /= p.GW;
            const int gh = (int)(mm % p.GH); mm /= p.GH;
            const int gd = (int)(mm % p.GD); mm /= p.GD;
            const int nb = (int)mm;
            const int ca = a0 + lv * 8;
            if (ca < p.PC) rp = __ldg(reinterpret_cast<const uint4*>(p.P + (size_t)m * p.PC + ca));
            const int qz = gd * p.istrD + p.offD + kd, qy = gh * p.istrH + p.offH + kh, qx = gw * p.istrW + p.offW + kw;
            const int cb = b0 + lv * 8;
            if (cb < QCtot && qz >= 0 && qz < p.QD && qy >= 0 && qy < p.QH && qx >= 0 && qx < p.QW) {
                const bf16* qp = p.Q[0];
                int cs = cb, qc = p.QC[0];
                if (p.nq > 1 && cb >= p.QC[0]) { qp = p.Q[1]; cs = cb - p.QC[0]; qc = p.QC[1]; }
                const size_t vox = (((size_t)nb * p.QD + qz) * p.QH + qy) * p.QW + qx;
                rq = __ldg(reinterpret_cast<const uint4*>(qp + vox * qc + cs));
            }
        }
    };

    prefetch(mBeg);
    for (long long mc = mBeg; mc < mEnd; mc += GW_BK) {
        *reinterpret_cast<uint4*>(&Ps[lr * GW_PITCH + lv * 8]) = rp;
        *reinterpret_cast<uint4*>(&Qs[lr * GW_PITCH + lv * 8]) = rq;
        __syncthreads();
        if (mc + GW_BK < mEnd) prefetch(mc + GW_BK);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[4];
            {
                // A'[a][k] stored as Ps[k][a]: transposed 8x8 loads
                const int kr = ks * 16 + (lane >> 4) * 8 + (lane & 7);
                const int ac = warp_a * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(smem_u32(&Ps[kr * GW_PITCH + ac]), af[0], af[1], af[2], af[3]);
            }
#pragma unroll
            for (int nj = 0; nj < 2; ++nj) {
                uint32_t b00, b01, b10, b11;
                const int kr = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
                const int bc = warp_b * 32 + nj * 16 + (lane >> 4) * 8;
                ldsm_x4_t(smem_u32(&Qs[kr * GW_PITCH + bc]), b00, b01, b10, b11);
                mma_bf16_16816(acc[nj * 2], af, b00, b01);
                mma_bf16_16816(acc[nj * 2 + 1], af, b10, b11);
            }
        }
        __syncthreads();
    }

    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int a = a0 + warp_a * 16 + half * 8 + g;
            const int b = b0 + warp_b * 32 + ni * 8 + q * 2;
            if (a < p.PC) {
                float* d = p.dw + ((size_t)t * p.PC + a) * QCtot + b;
                if (b < QCtot) atomicAdd(d, acc[ni][half * 2]);
                if (b + 1 < QCtot) atomicAdd(d + 1, acc[ni][half * 2 + 1]);
            }
        }
    }
}

}  // namespace rb
