// Split-precision activations of the fp32-accurate inference tier ("bf16x3").
//
// A value v is carried as hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits together) and a weight likewise, and
//     v * w  ~=  hi_v*hi_w + lo_v*hi_w + hi_v*lo_w            (the dropped lo*lo term is ~2^-18 relative)
// is ONE implicit GEMM with fp32 accumulation when the activation row of a voxel is laid out as the 3C channels
// [ hi(C) | lo(C) | hi(C) ] and the weight rows as [ hi_w | hi_w | lo_w ]: the existing gather-conv kernels run it
// unchanged with Cin' = 3*Cin.  The kernels below are the only new device code the tier needs: they produce that
// layout from fp32 pre-norm tensors (norm / gate / residual / LeakyReLU applied in fp32), from the raw network input
// (stem im2col) and through the ResNet-D average pool.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace rb {

// hi / lo halves of 8 fp32 values as two packed bf16x8 vectors
__device__ __forceinline__ void split8(const float (&a)[8], uint4& hi, uint4& lo) {
    hi = pack8(a);
    float h[8], l[8];
    unpack8(hi, h);
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = a[j] - h[j];
    lo = pack8(l);
}

// z[hi|lo|hi] = act( y * scale[n,c] + shift[n,c] + (res_hi + res_lo) ), all in fp32.
// y: [NB][S][C] fp32; res, z: [NB][S][3C] bf16; scale / shift: [NB][C] fp32 or both null (identity).
struct SplitApplyParams {
    const float* y;
    const bf16* res;   // may be null
    bf16* z;
    const float* scale;
    const float* shift;
    long long S;
    int NB, C, act;
    float slope;
};

__global__ void __launch_bounds__(256) split_apply_kernel(const SplitApplyParams p) {
    // grid = (blocks, NB); 32-bit index math inside one sample (S * C/8 < 2^31, checked on the host)
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % cg;
        const uint32_t v = i / cg;
        float a[8];
        load8_prenorm(p.y, ((size_t)nb * per + i) * 8, 1, a);
        if (p.scale != nullptr) {
            const size_t cidx = (size_t)nb * p.C + g * 8;
            float sc[8], sh[8];
            *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx));
            *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + 4));
            *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx));
            *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], sc[j], sh[j]);
        }
        const size_t row = ((size_t)nb * (size_t)p.S + v) * 3 * (size_t)p.C + g * 8;   // element offset of hi
        if (p.res != nullptr) {
            float rh[8], rl[8];
            unpack8(ld_stream(reinterpret_cast<const uint4*>(p.res + row)), rh);
            unpack8(ld_stream(reinterpret_cast<const uint4*>(p.res + row + p.C)), rl);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += rh[j] + rl[j];
        }
        if (p.act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] > 0.f ? a[j] : a[j] * p.slope;
        }
        uint4 hi, lo;
        split8(a, hi, lo);
        *reinterpret_cast<uint4*>(p.z + row) = hi;
        *reinterpret_cast<uint4*>(p.z + row + p.C) = lo;
        *reinterpret_cast<uint4*>(p.z + row + 2 * (size_t)p.C) = hi;
    }
}

// AvgPool3d(kernel = stride) on split activations: the window mean of (hi + lo) in fp32, re-split.
// in: [NB, D, H, W, 3C], out: [NB, D/sd, H/sh, W/sw, 3C]; C = logical channels.
__global__ void __launch_bounds__(256) avgpool_split_kernel(const PoolParams p) {
    const int cg = p.C >> 3;
    const int OD = p.D / p.sd, OH = p.H / p.sh, OW = p.W / p.sw;
    const long long total = (long long)p.NB * OD * OH * OW * cg;
    const float inv = 1.f / (float)(p.sd * p.sh * p.sw);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int g = (int)(t % cg); t /= cg;
        const int ow = (int)(t % OW); t /= OW;
        const int oh = (int)(t % OH); t /= OH;
        const int od = (int)(t % OD); t /= OD;
        const int nb = (int)t;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int a = 0; a < p.sd; ++a)
            for (int b = 0; b < p.sh; ++b)
                for (int c = 0; c < p.sw; ++c) {
                    const size_t vox = (((size_t)nb * p.D + od * p.sd + a) * p.H + oh * p.sh + b) * p.W + ow * p.sw + c;
                    const bf16* src = p.in + vox * 3 * (size_t)p.C + g * 8;
                    float fh[8], fl[8];
                    unpack8(ld_stream(reinterpret_cast<const uint4*>(src)), fh);
                    unpack8(ld_stream(reinterpret_cast<const uint4*>(src + p.C)), fl);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += fh[j] + fl[j];
                }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv;
        uint4 hi, lo;
        split8(acc, hi, lo);
        const size_t ovox = (((size_t)nb * OD + od) * OH + oh) * OW + ow;
        bf16* dst = p.out + ovox * 3 * (size_t)p.C + g * 8;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + p.C) = lo;
        *reinterpret_cast<uint4*>(dst + 2 * (size_t)p.C) = hi;
    }
}

// Stem im2col of the raw NCDHW fp32 input into split columns: col [NB, D, H, W, 3*Kp] = [hi(Kp) | lo(Kp) | hi(Kp)],
// column k = tap * Cin + ci, zero padded to Kp.
__global__ void __launch_bounds__(256) stem_im2col_split_kernel(const Im2colParams p) {
    const int kg = p.Kp >> 3;
    const long long S = (long long)p.D * p.H * p.W;
    const long long total = (long long)p.NB * S * kg;
    const int K = p.kd * p.kh * p.kw * p.Cin;
    const int pd = (p.kd - 1) / 2, ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % kg);
        long long v = i / kg;
        const long long vox = v;
        const int w = (int)(v % p.W); v /= p.W;
        const int h = (int)(v % p.H); v /= p.H;
        const int d = (int)(v % p.D); v /= p.D;
        const int nb = (int)v;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = g * 8 + j;
            float val = 0.f;
            if (k < K) {
                const int ci = k % p.Cin;
                const int t = k / p.Cin;
                const int tw = t % p.kw, th = (t / p.kw) % p.kh, td = t / (p.kw * p.kh);
                const int z = d + td - pd, y = h + th - ph, x = w + tw - pw;
                if (z >= 0 && z < p.D && y >= 0 && y < p.H && x >= 0 && x < p.W)
                    val = __ldg(p.x + (((size_t)nb * p.Cin + ci) * p.D + z) * p.H * p.W + (size_t)y * p.W + x);
            }
            f[j] = val;
        }
        uint4 hi, lo;
        split8(f, hi, lo);
        bf16* dst = p.col + (size_t)vox * 3 * (size_t)p.Kp + g * 8;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + p.Kp) = lo;
        *reinterpret_cast<uint4*>(dst + 2 * (size_t)p.Kp) = hi;
    }
}

}  // namespace rb
