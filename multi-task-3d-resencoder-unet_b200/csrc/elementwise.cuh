// HBM-bound kernels of the ResEnc U-Net hot path: InstanceNorm statistics, the fused
// normalise + affine/SE-gate + residual + LeakyReLU pass and its backward, AvgPool, the
// 1x1x1 task heads, stem im2col and weight (un)packing.  All activations are channels-last
// (NDHWC) bf16 with C % 8 == 0; every access is a 16-byte vector, coalesced along C.
//
// Reference arithmetic (file:line in /root/reference):
//   InstanceNorm3d(affine=False|True, eps)          build_network_from_config.py:172, simple_conv_blocks.py:58-60
//   LeakyReLU(0.01), out += residual                 resblocks.py:76,113-114
//   SqueezeExcite gate (DNA, not vendored)           resblocks.py:86-87,111-112
//   AvgPool3d(stride, stride) in the ResNet-D skip   resblocks.py:92-95
//   seg head Conv3d(C, classes, 1, bias=True)        decoder.py:131,151-152
#pragma once
#include "common.cuh"

namespace rb {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    return o;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 8 consecutive channels of a pre-norm tensor: f32 = 0 bf16, 1 fp32, 2 fp16 (elem = element offset)
__device__ __forceinline__ void load8_prenorm(const void* base, size_t elem, int f32, float (&f)[8]) {
    if (f32 == 1) {
        const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
        float4 a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(q));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(q + 1));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else if (f32 == 2) {     // fp16 pre-norm tensor
        const uint4 v = ld_stream(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(base) + elem));
        f[0] = f16lo(v.x); f[1] = f16hi(v.x); f[2] = f16lo(v.y); f[3] = f16hi(v.y);
        f[4] = f16lo(v.z); f[5] = f16hi(v.z); f[6] = f16lo(v.w); f[7] = f16hi(v.w);
    } else {
        unpack8(ld_stream(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + elem)), f);
    }
}

// ---------------------------------------------------------------------------------------
// Per-(sample, [w], channel) reductions.  out is double [NB][G][C][2] with G = 1 or W.
//   kind 0: (sum y, sum y^2)                        -> InstanceNorm statistics / SE squeeze
//   kind 1: (sum g, sum g*y), g = dz * lrelu'(z)    -> InstanceNorm / gate / affine backward
// grid = (blocks per sample, NB); a block strides over voxels, thread = (voxel row, 8 channels).
// ---------------------------------------------------------------------------------------
struct ReduceParams {
    const void* y;   // [NB, S, C]  bf16 or fp32 (yF32)
    const bf16* dz;  // kind 1
    const bf16* z;   // kind 1 with act: sign source (may be null => g = dz, or the sign is recomputed from y)
    const float* sgnA;   // kind 1, z == null: lrelu'(z) from sign(fmaf(y, sgnA[n][c], sgnB[n][c])) - the expression the
    const float* sgnB;   // forward pass rounded to z, so the stored activation need not be read again
    double* out;
    long long S;
    int C, W, perW;
    float slope;
    int kind;
    int yF32;
};

// SGN: the sign-from-pre-norm variant (separate instantiation: its 16 extra registers would cost the default
// path one resident block per SM)
template <bool SGN>
__global__ void __launch_bounds__(256) plane_reduce_kernel(const ReduceParams p) {
    extern __shared__ float red[];  // [rows][cg*16]
    const int cg = p.C >> 3;
    const int rows = 256 / cg > 0 ? 256 / cg : 1;
    const int nb = blockIdx.y;
    // when C/8 > 256 a thread loops over several channel groups
    for (int cbase = 0; cbase < cg; cbase += 256) {
        const int mycg = cbase + (threadIdx.x % (cg < 256 ? cg : 256));
        const int myrow = threadIdx.x / (cg < 256 ? cg : 256);
        const bool active = myrow < rows && mycg < cg;
        if (!p.perW) {
            float s1[8], s2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
            if (active) {
                const size_t base = (size_t)nb * p.S;
                for (long long v = (long long)blockIdx.x * rows + myrow; v < p.S; v += (long long)gridDim.x * rows) {
                    const size_t off = (base + v) * p.C + (size_t)mycg * 8;
                    float a[8], b[8];
                    load8_prenorm(p.y, off, p.yF32, a);
                    if (p.kind == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s1[j] += a[j]; s2[j] += a[j] * a[j]; }
                    } else {
                        unpack8(ld_stream(reinterpret_cast<const uint4*>(p.dz + off)), b);
                        if (p.z != nullptr) {
                            float zz[8];
                            unpack8(ld_stream(reinterpret_cast<const uint4*>(p.z + off)), zz);
#pragma unroll
                            for (int j = 0; j < 8; ++j) b[j] = zz[j] > 0.f ? b[j] : b[j] * p.slope;
                        } else if (SGN) {
                            // scale / shift re-read per row (L1 hits) instead of 16 live registers per thread
                            const size_t cs = (size_t)nb * p.C + (size_t)mycg * 8;
                            float sA[8], sB[8];
                            *reinterpret_cast<float4*>(sA) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs));
                            *reinterpret_cast<float4*>(sA + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs + 4));
                            *reinterpret_cast<float4*>(sB) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs));
                            *reinterpret_cast<float4*>(sB + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs + 4));
#pragma unroll
                            for (int j = 0; j < 8; ++j) b[j] = fmaf(a[j], sA[j], sB[j]) > 0.f ? b[j] : b[j] * p.slope;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s1[j] += b[j]; s2[j] += b[j] * a[j]; }
                    }
                }
            }
            // reduce over rows through shared memory
            const int width = (cg < 256 ? cg : 256) * 16;
            if (active) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    red[myrow * width + (threadIdx.x % (cg < 256 ? cg : 256)) * 16 + j] = s1[j];
                    red[myrow * width + (threadIdx.x % (cg < 256 ? cg : 256)) * 16 + 8 + j] = s2[j];
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                float acc = 0.f;
                for (int r = 0; r < rows; ++r) acc += red[r * width + i];
                const int g = i >> 4, j = i & 15;
                const int c = (cbase + g) * 8 + (j & 7);
                if (c < p.C) atomicAdd(p.out + ((size_t)nb * p.C + c) * 2 + (j >> 3), (double)acc);
            }
            __syncthreads();
        } else {
            // per-w variant: blockIdx.x enumerates w; rows stride over (d,h)
            const int w = blockIdx.x;
            const long long DH = p.S / p.W;
            float s1[8], s2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
            if (active) {
                for (long long dh = myrow; dh < DH; dh += rows) {
                    const size_t off = ((size_t)nb * p.S + dh * p.W + w) * p.C + (size_t)mycg * 8;
                    float a[8], b[8];
                    load8_prenorm(p.y, off, p.yF32, a);
                    if (p.kind == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s1[j] += a[j]; s2[j] += a[j] * a[j]; }
                    } else {
                        unpack8(ld_stream(reinterpret_cast<const uint4*>(p.dz + off)), b);
                        if (p.z != nullptr) {
                            float zz[8];
                            unpack8(ld_stream(reinterpret_cast<const uint4*>(p.z + off)), zz);
#pragma unroll
                            for (int j = 0; j < 8; ++j) b[j] = zz[j] > 0.f ? b[j] : b[j] * p.slope;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s1[j] += b[j]; s2[j] += b[j] * a[j]; }
                    }
                }
            }
            const int width = (cg < 256 ? cg : 256) * 16;
            if (active) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    red[myrow * width + (threadIdx.x % (cg < 256 ? cg : 256)) * 16 + j] = s1[j];
                    red[myrow * width + (threadIdx.x % (cg < 256 ? cg : 256)) * 16 + 8 + j] = s2[j];
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                float acc = 0.f;
                for (int r = 0; r < rows; ++r) acc += red[r * width + i];
                const int g = i >> 4, j = i & 15;
                const int c = (cbase + g) * 8 + (j & 7);
                if (c < p.C) p.out[(((size_t)nb * p.W + w) * p.C + c) * 2 + (j >> 3)] = (double)acc;
            }
            __syncthreads();
        }
    }
}

// Round-2 variant of the per-(n, c) reduction for 2-byte pre-norm tensors (one coefficient set per sample, not the
// per-w variant).  ncu on the kernel above at 32 ch @128^3 x2 (profiles/r2_ncu_norm_passes.csv): 137 us for 537 MB =
// 3.9 TB/s, issue slots 30 % busy, 32 resident warps per SM, every warp waiting on its two 16-byte loads (long
// scoreboard) - ~32 KB in flight per SM, not enough to cover the DRAM latency at 6.5 TB/s.  Here a thread issues the
// loads of TWO voxels before it touches either (raw 16-byte registers, unpacked one voxel at a time), which doubles the
// bytes in flight at +8 registers.
template <bool SGN>
__global__ void __launch_bounds__(256, 4) plane_reduce_u2_kernel(const ReduceParams p) {
    extern __shared__ float red[];  // [rows][cg*16]
    const int cg = p.C >> 3;        // <= 256 (host-checked)
    const int rows = 256 / cg;
    const int nb = blockIdx.y;
    const int mycg = threadIdx.x % cg;
    const int myrow = threadIdx.x / cg;
    const bool active = myrow < rows;
    const bool yh = p.yF32 == 2;
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    if (active) {
        // 32-bit index math inside one sample (S * C/8 < 2^31, checked on the host)
        const size_t base = (size_t)nb * p.S * cg + mycg;
        const uint4* yv = reinterpret_cast<const uint4*>(p.y) + base;
        const uint4* dzv = reinterpret_cast<const uint4*>(p.dz) + base;
        const uint4* zv = reinterpret_cast<const uint4*>(p.z) + base;
        const uint32_t S = (uint32_t)p.S;
        const uint32_t step = gridDim.x * (uint32_t)rows;
        const size_t cs = (size_t)nb * p.C + (size_t)mycg * 8;
        auto accum = [&](const uint4& ry, const uint4& rdz, const uint4& rz) {
            float a[8], b[8];
            if (yh) {
                a[0] = f16lo(ry.x); a[1] = f16hi(ry.x); a[2] = f16lo(ry.y); a[3] = f16hi(ry.y);
                a[4] = f16lo(ry.z); a[5] = f16hi(ry.z); a[6] = f16lo(ry.w); a[7] = f16hi(ry.w);
            } else {
                unpack8(ry, a);
            }
            if (p.kind == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { s1[j] += a[j]; s2[j] = fmaf(a[j], a[j], s2[j]); }
                return;
            }
            unpack8(rdz, b);
            if (!SGN && p.z != nullptr) {
                float zz[8];
                unpack8(rz, zz);
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = zz[j] > 0.f ? b[j] : b[j] * p.slope;
            } else if (SGN) {
                // scale / shift re-read per voxel (L1 hits) instead of 16 live registers per thread
                float sA[8], sB[8];
                *reinterpret_cast<float4*>(sA) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs));
                *reinterpret_cast<float4*>(sA + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs + 4));
                *reinterpret_cast<float4*>(sB) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs));
                *reinterpret_cast<float4*>(sB + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs + 4));
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = fmaf(a[j], sA[j], sB[j]) > 0.f ? b[j] : b[j] * p.slope;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += b[j]; s2[j] = fmaf(b[j], a[j], s2[j]); }
        };
        for (uint32_t v = blockIdx.x * (uint32_t)rows + myrow; v < S; v += 2u * step) {
            const bool two = v + step < S;
            const uint32_t o0 = v * (uint32_t)cg, o1 = (v + step) * (uint32_t)cg;
            uint4 y0, y1, d0, d1, z0, z1;
            y0 = ld_stream(yv + o0);
            if (p.kind != 0) d0 = ld_stream(dzv + o0);
            if (!SGN && p.kind != 0 && p.z != nullptr) z0 = ld_stream(zv + o0);
            if (two) {
                y1 = ld_stream(yv + o1);
                if (p.kind != 0) d1 = ld_stream(dzv + o1);
                if (!SGN && p.kind != 0 && p.z != nullptr) z1 = ld_stream(zv + o1);
            }
            accum(y0, d0, z0);
            if (two) accum(y1, d1, z1);
        }
    }
    const int width = cg * 16;
    if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[myrow * width + mycg * 16 + j] = s1[j];
            red[myrow * width + mycg * 16 + 8 + j] = s2[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < rows; ++r) acc += red[r * width + i];
        const int g = i >> 4, j = i & 15;
        atomicAdd(p.out + ((size_t)nb * p.C + g * 8 + (j & 7)) * 2 + (j >> 3), (double)acc);
    }
}

// ---------------------------------------------------------------------------------------
// Fused apply:   z = act( y * scale[n,(w),c] + shift[n,(w),c] + res )
// scale/shift fold mean, rstd, affine gamma/beta and the SE gate (host glue builds them
// from the statistics; they are [NB][G][C] fp32 with G = 1 or W).
// ---------------------------------------------------------------------------------------
struct ApplyParams {
    const void* y;    // bf16 or fp32 (yF32)
    const bf16* res;  // may be null
    bf16* z;
    const float* scale;
    const float* shift;
    long long S;
    int NB, C, W, perW, act;
    float slope;
    int yF32;
};

__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const ApplyParams p) {
    // grid = (blocks, NB); 32-bit index math inside one sample (S * C/8 < 2^31, checked on the host)
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const uint4* rv = p.res ? reinterpret_cast<const uint4*>(p.res) + (size_t)nb * per : nullptr;
    uint4* zv = reinterpret_cast<uint4*>(p.z) + (size_t)nb * per;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % cg;
        const uint32_t v = i / cg;
        const size_t cidx = p.perW ? (((size_t)nb * p.W + (v % (uint32_t)p.W)) * p.C + g * 8) : ((size_t)nb * p.C + g * 8);
        float a[8], sc[8], sh[8];
        load8_prenorm(p.y, ((size_t)nb * per + i) * 8, p.yF32, a);
        *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx));
        *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + 4));
        *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx));
        *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], sc[j], sh[j]);
        if (rv != nullptr) {
            float r[8];
            unpack8(ld_stream(rv + i), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += r[j];
        }
        if (p.act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] > 0.f ? a[j] : a[j] * p.slope;
        }
        zv[i] = pack8(a);
    }
}

// Two consecutive 8-channel vectors per thread (C % 16 == 0): with a 2-byte pre-norm tensor the kernel above has ONE
// 16-byte streaming load in flight per thread (24 KB per SM: bytes-in-flight bound at ~4.2 TB/s, tools/norm_bench.py);
// here a thread streams 32 contiguous bytes of y (and of the residual) and stores 32 contiguous bytes of z.
__global__ void __launch_bounds__(256) norm_act_fwd_x2_kernel(const ApplyParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;                 // even
    const uint32_t per = (uint32_t)p.S * cg;                // vectors per sample (even)
    const uint32_t pairs = per >> 1;
    const int nb = blockIdx.y;
    const uint4* rv = p.res ? reinterpret_cast<const uint4*>(p.res) + (size_t)nb * per : nullptr;
    uint4* zv = reinterpret_cast<uint4*>(p.z) + (size_t)nb * per;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += gridDim.x * blockDim.x) {
        const uint32_t i = q * 2;
        const uint32_t g = i % cg;
        const uint32_t v = i / cg;
        const size_t cidx = p.perW ? (((size_t)nb * p.W + (v % (uint32_t)p.W)) * p.C + g * 8) : ((size_t)nb * p.C + g * 8);
        float a[2][8], r[2][8];
        load8_prenorm(p.y, ((size_t)nb * per + i) * 8, p.yF32, a[0]);
        load8_prenorm(p.y, ((size_t)nb * per + i + 1) * 8, p.yF32, a[1]);
        if (rv != nullptr) {
            unpack8(ld_stream(rv + i), r[0]);
            unpack8(ld_stream(rv + i + 1), r[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float sc[8], sh[8];
            *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + u * 8));
            *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + u * 8 + 4));
            *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + u * 8));
            *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + u * 8 + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) a[u][j] = fmaf(a[u][j], sc[j], sh[j]);
            if (rv != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[u][j] += r[u][j];
            }
            if (p.act) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[u][j] = a[u][j] > 0.f ? a[u][j] : a[u][j] * p.slope;
            }
            zv[i + u] = pack8(a[u]);
        }
    }
}

// Several vectors in flight per thread for the forward apply pass (2-byte pre-norm tensor): all loads of an iteration
// are issued first and held as raw 16-byte registers, then unpacked and finished one vector at a time.  The one-vector
// kernel keeps 48 warps x one 16-byte load = 24 KB per SM in flight (long-scoreboard bound, issue slots 50 % busy,
// 4.4 TB/s at 32 ch @128^3 x2); U = 4 without a residual / U = 2 with one put 64 KB in flight at 4 blocks per SM.
template <bool YH, bool HAS_RES, int U>
__global__ void __launch_bounds__(256, 4) norm_act_fwd_un_kernel(const ApplyParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const size_t base = (size_t)nb * per;
    const uint4* yv = reinterpret_cast<const uint4*>(p.y) + base;
    const uint4* rv = HAS_RES ? reinterpret_cast<const uint4*>(p.res) + base : nullptr;
    uint4* zv = reinterpret_cast<uint4*>(p.z) + base;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < per; i0 += (uint32_t)U * stride) {
        uint4 ry[U], rr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < per) {
                ry[u] = ld_stream(yv + i);
                if (HAS_RES) rr[u] = ld_stream(rv + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i >= per) break;
            const uint32_t g = i % cg;
            const uint32_t v = i / cg;
            const size_t cidx = p.perW ? (((size_t)nb * p.W + (v % (uint32_t)p.W)) * p.C + g * 8) : ((size_t)nb * p.C + g * 8);
            float a[8], sc[8], sh[8];
            if (YH) {
                a[0] = f16lo(ry[u].x); a[1] = f16hi(ry[u].x); a[2] = f16lo(ry[u].y); a[3] = f16hi(ry[u].y);
                a[4] = f16lo(ry[u].z); a[5] = f16hi(ry[u].z); a[6] = f16lo(ry[u].w); a[7] = f16hi(ry[u].w);
            } else {
                unpack8(ry[u], a);
            }
            *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx));
            *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + 4));
            *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx));
            *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], sc[j], sh[j]);
            if (HAS_RES) {
                float r[8];
                unpack8(rr[u], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] += r[j];
            }
            if (p.act) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = a[j] > 0.f ? a[j] : a[j] * p.slope;
            }
            zv[i] = pack8(a);
        }
    }
}

// Backward of the fused apply + InstanceNorm:
//   g    = dz * lrelu'(z)                 (z = saved output; act==0 => g = dz)
//   dres = g                              (only when dres != null)
//   dy   = g * k1[n,(w),c] + y * k2[n,(w),c] + k3[n,(w),c]
struct ApplyBwdParams {
    const bf16* dz;
    const bf16* z;
    const void* y;    // bf16 or fp32 (yF32)
    bf16* dy;
    bf16* dres;
    const float* k1;
    const float* k2;
    const float* k3;
    long long S;
    int NB, C, W, perW, act;
    float slope;
    int yF32;
    const float* sgnA;   // act && z == null: sign of the activation input recomputed as fmaf(y, sgnA, sgnB) > 0
    const float* sgnB;
};

__global__ void __launch_bounds__(256) norm_act_bwd_kernel(const ApplyBwdParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const size_t base = (size_t)nb * per;
    const uint4* dzv = reinterpret_cast<const uint4*>(p.dz) + base;
    const uint4* zv = p.z ? reinterpret_cast<const uint4*>(p.z) + base : nullptr;
    uint4* dyv = reinterpret_cast<uint4*>(p.dy) + base;
    uint4* drv = p.dres ? reinterpret_cast<uint4*>(p.dres) + base : nullptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
        const uint32_t g = i % cg;
        const uint32_t v = i / cg;
        const size_t c1 = p.perW ? (((size_t)nb * p.W + (v % (uint32_t)p.W)) * p.C + g * 8) : ((size_t)nb * p.C + g * 8);
        const size_t c2 = c1;
        float gd[8], yy[8], a1[8], a2[8], a3[8];
        unpack8(ld_stream(dzv + i), gd);
        load8_prenorm(p.y, (base + i) * 8, p.yF32, yy);
        if (p.act && zv != nullptr) {
            float zz[8];
            unpack8(ld_stream(zv + i), zz);
#pragma unroll
            for (int j = 0; j < 8; ++j) gd[j] = zz[j] > 0.f ? gd[j] : gd[j] * p.slope;
        } else if (p.act) {
            const size_t cs = (size_t)nb * p.C + g * 8;
            float sa[8], sb[8];
            *reinterpret_cast<float4*>(sa) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs));
            *reinterpret_cast<float4*>(sa + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs + 4));
            *reinterpret_cast<float4*>(sb) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs));
            *reinterpret_cast<float4*>(sb + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) gd[j] = fmaf(yy[j], sa[j], sb[j]) > 0.f ? gd[j] : gd[j] * p.slope;
        }
        if (drv != nullptr) drv[i] = pack8(gd);
        *reinterpret_cast<float4*>(a1) = __ldg(reinterpret_cast<const float4*>(p.k1 + c1));
        *reinterpret_cast<float4*>(a1 + 4) = __ldg(reinterpret_cast<const float4*>(p.k1 + c1 + 4));
        *reinterpret_cast<float4*>(a2) = __ldg(reinterpret_cast<const float4*>(p.k2 + c2));
        *reinterpret_cast<float4*>(a2 + 4) = __ldg(reinterpret_cast<const float4*>(p.k2 + c2 + 4));
        *reinterpret_cast<float4*>(a3) = __ldg(reinterpret_cast<const float4*>(p.k3 + c2));
        *reinterpret_cast<float4*>(a3 + 4) = __ldg(reinterpret_cast<const float4*>(p.k3 + c2 + 4));
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(gd[j], a1[j], fmaf(yy[j], a2[j], a3[j]));
        dyv[i] = pack8(o);
    }
}

// Two vectors in flight per thread for the backward apply pass (2-byte pre-norm tensor): the loads of both vectors are
// issued first and held as raw 16-byte registers (8-12 registers instead of the 32-48 floats the hoisted-coefficient
// variant below kept live), unpacked one vector at a time; coefficients stay L1 reads.  ncu on the kernel above at
// 32 ch @128^3 x2: 180 us, 4.5 TB/s, issue slots 38 % busy, long-scoreboard bound with 40 warps x 2 loads in flight.
// ZM: how lrelu'(z) is obtained - 0 no activation, 1 from the stored z, 2 from the sign of fmaf(y, sgnA, sgnB)
template <bool YH, int ZM>
__global__ void __launch_bounds__(256, 4) norm_act_bwd_u2_kernel(const ApplyBwdParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const size_t base = (size_t)nb * per;
    const uint4* dzv = reinterpret_cast<const uint4*>(p.dz) + base;
    const uint4* yv = reinterpret_cast<const uint4*>(p.y) + base;
    constexpr bool hasz = ZM == 1;
    constexpr bool sgn = ZM == 2;
    const uint4* zv = hasz ? reinterpret_cast<const uint4*>(p.z) + base : nullptr;
    uint4* dyv = reinterpret_cast<uint4*>(p.dy) + base;
    uint4* drv = p.dres ? reinterpret_cast<uint4*>(p.dres) + base : nullptr;
    const uint32_t stride = gridDim.x * blockDim.x;
    auto apply = [&](uint32_t i, const uint4& rdz, const uint4& ry, const uint4& rz) {
        const uint32_t g = i % cg;
        const uint32_t v = i / cg;
        const size_t c1 = p.perW ? (((size_t)nb * p.W + (v % (uint32_t)p.W)) * p.C + g * 8) : ((size_t)nb * p.C + g * 8);
        float gd[8], yy[8], a1[8], a2[8], a3[8];
        unpack8(rdz, gd);
        if (YH) {
            yy[0] = f16lo(ry.x); yy[1] = f16hi(ry.x); yy[2] = f16lo(ry.y); yy[3] = f16hi(ry.y);
            yy[4] = f16lo(ry.z); yy[5] = f16hi(ry.z); yy[6] = f16lo(ry.w); yy[7] = f16hi(ry.w);
        } else {
            unpack8(ry, yy);
        }
        if (hasz) {
            float zz[8];
            unpack8(rz, zz);
#pragma unroll
            for (int j = 0; j < 8; ++j) gd[j] = zz[j] > 0.f ? gd[j] : gd[j] * p.slope;
        } else if (sgn) {
            const size_t cs = (size_t)nb * p.C + g * 8;
            float sa[8], sb[8];
            *reinterpret_cast<float4*>(sa) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs));
            *reinterpret_cast<float4*>(sa + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnA + cs + 4));
            *reinterpret_cast<float4*>(sb) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs));
            *reinterpret_cast<float4*>(sb + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnB + cs + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) gd[j] = fmaf(yy[j], sa[j], sb[j]) > 0.f ? gd[j] : gd[j] * p.slope;
        }
        if (drv != nullptr) drv[i] = pack8(gd);
        *reinterpret_cast<float4*>(a1) = __ldg(reinterpret_cast<const float4*>(p.k1 + c1));
        *reinterpret_cast<float4*>(a1 + 4) = __ldg(reinterpret_cast<const float4*>(p.k1 + c1 + 4));
        *reinterpret_cast<float4*>(a2) = __ldg(reinterpret_cast<const float4*>(p.k2 + c1));
        *reinterpret_cast<float4*>(a2 + 4) = __ldg(reinterpret_cast<const float4*>(p.k2 + c1 + 4));
        *reinterpret_cast<float4*>(a3) = __ldg(reinterpret_cast<const float4*>(p.k3 + c1));
        *reinterpret_cast<float4*>(a3 + 4) = __ldg(reinterpret_cast<const float4*>(p.k3 + c1 + 4));
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(gd[j], a1[j], fmaf(yy[j], a2[j], a3[j]));
        dyv[i] = pack8(o);
    };
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < per; i0 += 2u * stride) {
        const uint32_t i1 = i0 + stride;
        const bool two = i1 < per;
        uint4 d0, y0, z0, d1, y1, z1;
        d0 = ld_stream(dzv + i0);
        y0 = ld_stream(yv + i0);
        if (hasz) z0 = ld_stream(zv + i0);
        if (two) {
            d1 = ld_stream(dzv + i1);
            y1 = ld_stream(yv + i1);
            if (hasz) z1 = ld_stream(zv + i1);
        }
        apply(i0, d0, y0, z0);
        if (two) apply(i1, d1, y1, z1);
    }
}

// ---------------------------------------------------------------------------------------
// Round-2 variants of the two apply passes for the common case (one coefficient set per (n, c), C/8 a power of two
// <= 64): the per-channel coefficients are loaded ONCE per thread (a thread's channel group is the same in every
// grid-stride iteration because the stride is a multiple of C/8) instead of 4-10 float4 L1 loads per 8 elements, and
// U = 2 independent 8-element groups are in flight per iteration.  The round-1 kernels above ran the 128^3 launches at
// 4.5 TB/s (bwd) / 5.4 TB/s (fwd with a 2-byte pre-norm tensor): load-instruction bound, not DRAM bound.
// MEASURED SLOWER than the round-1 kernels (tools/norm_bench.py, profiles/r2_norm_bench.txt: 137 vs 127 us forward,
// 353 vs 279 us backward at 32 ch @128^3 x2 - 62 / 113 registers cost more occupancy than the saved L1 loads return),
// so they are NOT the default (RESENC_NORM_VARIANT=1 selects them).
// ---------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) norm_act_fwd_v1_kernel(const ApplyParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint32_t g = tid & (cg - 1);
    const uint4* rv = p.res ? reinterpret_cast<const uint4*>(p.res) + (size_t)nb * per : nullptr;
    uint4* zv = reinterpret_cast<uint4*>(p.z) + (size_t)nb * per;
    float sc[8], sh[8];
    {
        const size_t cidx = (size_t)nb * p.C + g * 8;
        *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx));
        *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cidx + 4));
        *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx));
        *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cidx + 4));
    }
    for (uint32_t i0 = tid; i0 < per; i0 += U * stride) {
        float a[U][8], r[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < per) {
                load8_prenorm(p.y, ((size_t)nb * per + i) * 8, p.yF32, a[u]);
                if (rv != nullptr) unpack8(ld_stream(rv + i), r[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < per) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[u][j] = fmaf(a[u][j], sc[j], sh[j]);
                if (rv != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[u][j] += r[u][j];
                }
                if (p.act) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[u][j] = a[u][j] > 0.f ? a[u][j] : a[u][j] * p.slope;
                }
                zv[i] = pack8(a[u]);
            }
        }
    }
}

template <int U>
__global__ void __launch_bounds__(256) norm_act_bwd_v1_kernel(const ApplyBwdParams p) {
    const uint32_t cg = (uint32_t)p.C >> 3;
    const uint32_t per = (uint32_t)p.S * cg;
    const int nb = blockIdx.y;
    const size_t base = (size_t)nb * per;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint32_t g = tid & (cg - 1);
    const uint4* dzv = reinterpret_cast<const uint4*>(p.dz) + base;
    const uint4* zv = p.z ? reinterpret_cast<const uint4*>(p.z) + base : nullptr;
    uint4* dyv = reinterpret_cast<uint4*>(p.dy) + base;
    uint4* drv = p.dres ? reinterpret_cast<uint4*>(p.dres) + base : nullptr;
    const bool sgn = p.act && zv == nullptr;
    float a1[8], a2[8], a3[8], sa[8], sb[8];
    {
        const size_t c = (size_t)nb * p.C + g * 8;
        *reinterpret_cast<float4*>(a1) = __ldg(reinterpret_cast<const float4*>(p.k1 + c));
        *reinterpret_cast<float4*>(a1 + 4) = __ldg(reinterpret_cast<const float4*>(p.k1 + c + 4));
        *reinterpret_cast<float4*>(a2) = __ldg(reinterpret_cast<const float4*>(p.k2 + c));
        *reinterpret_cast<float4*>(a2 + 4) = __ldg(reinterpret_cast<const float4*>(p.k2 + c + 4));
        *reinterpret_cast<float4*>(a3) = __ldg(reinterpret_cast<const float4*>(p.k3 + c));
        *reinterpret_cast<float4*>(a3 + 4) = __ldg(reinterpret_cast<const float4*>(p.k3 + c + 4));
        if (sgn) {
            *reinterpret_cast<float4*>(sa) = __ldg(reinterpret_cast<const float4*>(p.sgnA + c));
            *reinterpret_cast<float4*>(sa + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnA + c + 4));
            *reinterpret_cast<float4*>(sb) = __ldg(reinterpret_cast<const float4*>(p.sgnB + c));
            *reinterpret_cast<float4*>(sb + 4) = __ldg(reinterpret_cast<const float4*>(p.sgnB + c + 4));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
        }
    }
    for (uint32_t i0 = tid; i0 < per; i0 += U * stride) {
        float gd[U][8], yy[U][8], zz[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < per) {
                unpack8(ld_stream(dzv + i), gd[u]);
                load8_prenorm(p.y, (base + i) * 8, p.yF32, yy[u]);
                if (p.act && zv != nullptr) unpack8(ld_stream(zv + i), zz[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < per) {
                if (p.act && zv != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) gd[u][j] = zz[u][j] > 0.f ? gd[u][j] : gd[u][j] * p.slope;
                } else if (sgn) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) gd[u][j] = fmaf(yy[u][j], sa[j], sb[j]) > 0.f ? gd[u][j] : gd[u][j] * p.slope;
                }
                if (drv != nullptr) drv[i] = pack8(gd[u]);
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(gd[u][j], a1[j], fmaf(yy[u][j], a2[j], a3[j]));
                dyv[i] = pack8(o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// AvgPool3d(kernel = stride) forward / backward, window (sd, sh, sw) in {1,2}^3.
// ---------------------------------------------------------------------------------------
struct PoolParams {
    const bf16* in;
    bf16* out;
    int NB, D, H, W, C;  // dims of the full-resolution tensor
    int sd, sh, sw;
};

// IDX: unsigned int when the element count fits 31 bits (64-bit div/mod chains made the pooling kernels run at 1 TB/s)
template <typename IDX>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const PoolParams p) {
    const int cg = p.C >> 3;
    const int OD = p.D / p.sd, OH = p.H / p.sh, OW = p.W / p.sw;
    const IDX total = (IDX)p.NB * OD * OH * OW * cg;
    const float inv = 1.f / (float)(p.sd * p.sh * p.sw);
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        IDX t = i;
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
                    float f[8];
                    unpack8(ld_stream(reinterpret_cast<const uint4*>(p.in + vox * p.C + g * 8)), f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += f[j];
                }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv;
        reinterpret_cast<uint4*>(p.out)[i] = pack8(acc);
    }
}

// in = d(pooled) [NB, D/sd, H/sh, W/sw, C]; out = d(full) [NB, D, H, W, C]
template <typename IDX>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const PoolParams p) {
    const int cg = p.C >> 3;
    const int OD = p.D / p.sd, OH = p.H / p.sh, OW = p.W / p.sw;
    const IDX total = (IDX)p.NB * p.D * p.H * p.W * cg;
    const float inv = 1.f / (float)(p.sd * p.sh * p.sw);
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        IDX t = i;
        const int g = (int)(t % cg); t /= cg;
        const int w = (int)(t % p.W); t /= p.W;
        const int h = (int)(t % p.H); t /= p.H;
        const int d = (int)(t % p.D); t /= p.D;
        const int nb = (int)t;
        const size_t vox = (((size_t)nb * OD + d / p.sd) * OH + h / p.sh) * OW + w / p.sw;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.in + vox * p.C + g * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= inv;
        reinterpret_cast<uint4*>(p.out)[i] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------
// Task head: 1x1x1 conv with bias, K <= 8 classes, output NCDHW fp32 (what the trainer's
// losses and the blend consume).  act: 0 none, 1 sigmoid, 2 softmax over classes
// (build_network_from_config.py:6-18,322-323).
// ---------------------------------------------------------------------------------------
static constexpr int HEAD_MAXK = 8;
struct HeadParams {
    const bf16* x;    // [NB, S, C]
    const float* w;   // [K][C]
    const float* b;   // [K]
    float* out;       // [NB][K][S]
    long long S;
    int NB, C, K, act;
};

__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadParams p) {
    extern __shared__ float hw[];  // [K][C] + [K]
    for (int i = threadIdx.x; i < p.K * p.C; i += blockDim.x) hw[i] = p.w[i];
    for (int i = threadIdx.x; i < p.K; i += blockDim.x) hw[p.K * p.C + i] = p.b ? p.b[i] : 0.f;
    __syncthreads();
    const long long total = (long long)p.NB * p.S;
    const int cg = p.C >> 3;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        float acc[HEAD_MAXK];
#pragma unroll
        for (int k = 0; k < HEAD_MAXK; ++k) acc[k] = k < p.K ? hw[p.K * p.C + k] : 0.f;
        for (int g = 0; g < cg; ++g) {
            float f[8];
            unpack8(ld_stream(reinterpret_cast<const uint4*>(p.x + (size_t)v * p.C) + g), f);
#pragma unroll
            for (int k = 0; k < HEAD_MAXK; ++k) {
                if (k < p.K) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], hw[k * p.C + g * 8 + j], acc[k]);
                }
            }
        }
        if (p.act == 1) {
#pragma unroll
            for (int k = 0; k < HEAD_MAXK; ++k) acc[k] = 1.f / (1.f + expf(-acc[k]));
        } else if (p.act == 2) {
            float mx = -INFINITY, sum = 0.f;
#pragma unroll
            for (int k = 0; k < HEAD_MAXK; ++k) if (k < p.K) mx = fmaxf(mx, acc[k]);
#pragma unroll
            for (int k = 0; k < HEAD_MAXK; ++k) if (k < p.K) { acc[k] = expf(acc[k] - mx); sum += acc[k]; }
#pragma unroll
            for (int k = 0; k < HEAD_MAXK; ++k) acc[k] /= sum;
        }
        const long long nb = v / p.S, s = v % p.S;
#pragma unroll
        for (int k = 0; k < HEAD_MAXK; ++k)
            if (k < p.K) p.out[((size_t)nb * p.K + k) * p.S + s] = acc[k];
    }
}

// ---------------------------------------------------------------------------------------
// Fused tail of a task decoder in inference (no autograd):
//     logits = head( act( y * scale[n,c] + shift[n,c] + res ) )          decoder.py:144-152 (last stage -> seg_layers[-1])
// The last activation z of a decoder is consumed only by its 1x1x1 head, so writing it (2 B / element) and reading it
// back (2 B) is pure HBM traffic: this kernel reads the pre-norm tensor once and writes K fp32 logit planes.  z is
// rounded to bf16 in registers and the plain head (rb_head_fwd) runs on this same kernel (NORM = 0, the launcher's choice
// when scale == nullptr), so the logits are bit-identical to the two-kernel path.
// Thread = (voxel, 8-channel group): a warp instruction reads 512 contiguous bytes; the C/8 lanes of a voxel combine
// their partial dot products with a butterfly (C/8 a power of two <= 32, host-checked).
// ---------------------------------------------------------------------------------------
struct NormHeadParams {
    const void* y;        // [NB, S, C] pre-norm (yF32: 0 bf16, 1 fp32, 2 fp16), or the stored activation when scale == null
    const bf16* res;      // may be null
    const float* scale;   // [NB][C] or null
    const float* shift;
    const float* w;       // [K][C]
    const float* b;       // [K] or null
    float* out;           // [NB][K][S]
    long long S;
    int NB, C, K, act, head_act;
    float slope;
    int yF32;
};

// Template flags keep every per-element decision out of the instruction stream: the first version tested the pre-norm
// type, residual and activation per element and needed ~300 instructions per 16-byte vector - issue bound at 1.9 TB/s.
//   NORM: 0 = plain head on a stored bf16 activation, 1 = normalise + activation first;  YMODE: element type of y
static constexpr int NH_U = 4;   // voxel sets (16-byte loads) in flight per thread: 3 blocks x 8 warps x 4 x 512 B = 48 KB per SM
template <int KMAX, int NORM, int YMODE, int HAS_RES>
__global__ void __launch_bounds__(256, 3) norm_act_head_fwd_kernel(const NormHeadParams p) {
    extern __shared__ float hw[];  // [K][C] + [K]
    for (int i = threadIdx.x; i < p.K * p.C; i += blockDim.x) hw[i] = p.w[i];
    for (int i = threadIdx.x; i < p.K; i += blockDim.x) hw[p.K * p.C + i] = p.b ? p.b[i] : 0.f;
    __syncthreads();
    const uint32_t cg = (uint32_t)p.C >> 3;          // lanes per voxel
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t g = lane & (cg - 1u);
    const uint32_t vsub = lane / cg;                 // voxel of this lane within the warp's group
    const uint32_t vpw = 32u / cg;                   // voxels per warp instruction
    const int nb = blockIdx.y;
    const uint32_t S = (uint32_t)p.S;
    const size_t base = (size_t)nb * p.S;
    const uint32_t stride = gridDim.x * 8u * vpw;
    const float slope = p.act ? p.slope : 1.f;       // act == 0: the select below is the identity
    const float* wrow = hw + g * 8;
    const size_t cs = (size_t)nb * p.C + g * 8;
    // the loop variable is warp-uniform (first voxel of the warp's group): the shuffles below run with all lanes
    for (uint32_t v0 = (blockIdx.x * 8u + (threadIdx.x >> 5)) * vpw; v0 < S; v0 += (uint32_t)NH_U * stride) {
        float a[NH_U][8], r[NH_U][8];
        bool ok[NH_U];
#pragma unroll
        for (int u = 0; u < NH_U; ++u) {
            const uint32_t v = v0 + u * stride + vsub;
            ok[u] = v < S;
            const size_t off = (base + (ok[u] ? v : 0u)) * p.C + g * 8;
            load8_prenorm(p.y, off, NORM ? YMODE : 0, a[u]);
            if (HAS_RES) unpack8(ld_stream(reinterpret_cast<const uint4*>(p.res + off)), r[u]);
        }
#pragma unroll
        for (int u = 0; u < NH_U; ++u) {
            if (u > 0 && v0 + u * stride >= S) break;   // warp-uniform
            if (NORM) {
                // coefficients re-read per vector (L1 hits) rather than held in 16 registers
                float sc[8], sh[8];
                *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + cs));
                *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + cs + 4));
                *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + cs));
                *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + cs + 4));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = fmaf(a[u][j], sc[j], sh[j]);
                    if (HAS_RES) t += r[u][j];
                    a[u][j] = t > 0.f ? t : t * slope;
                }
                unpack8(pack8(a[u]), a[u]);          // what the stored activation would hold: one rounding to bf16
            }
            float acc[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                acc[k] = 0.f;
                if (k < p.K) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wrow + k * p.C);
                    const float4 w1 = *reinterpret_cast<const float4*>(wrow + k * p.C + 4);
                    acc[k] = fmaf(a[u][0], w0.x, fmaf(a[u][1], w0.y, fmaf(a[u][2], w0.z, a[u][3] * w0.w))) +
                             fmaf(a[u][4], w1.x, fmaf(a[u][5], w1.y, fmaf(a[u][6], w1.z, a[u][7] * w1.w)));
                }
            }
            for (uint32_t m = 1; m < cg; m <<= 1) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], m);
            }
            if (g == 0 && ok[u]) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) acc[k] += k < p.K ? hw[p.K * p.C + k] : 0.f;
                if (p.head_act == 1) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) acc[k] = 1.f / (1.f + expf(-acc[k]));
                } else if (p.head_act == 2) {
                    float mx = -INFINITY, sum = 0.f;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < p.K) mx = fmaxf(mx, acc[k]);
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < p.K) { acc[k] = expf(acc[k] - mx); sum += acc[k]; }
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) acc[k] /= sum;
                }
                float* o = p.out + (size_t)nb * p.K * p.S + (v0 + u * stride + vsub);
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    if (k < p.K) o[(size_t)k * p.S] = acc[k];
            }
        }
    }
}

// Backward of the head (on raw logits): dx = dl . W (bf16), dW += dl^T x, db += sum dl.
// thread = (voxel, 8-channel group); block-level reduction of dW / db, then fp32 atomics.
struct HeadBwdParams {
    const bf16* x;
    const float* w;
    const float* dl;  // [NB][K][S]
    bf16* dx;         // [NB, S, C]
    float* dw;        // [K][C]   (accumulated)
    float* db;        // [K]
    long long S;
    int NB, C, K;
};

// KMAX: compile-time bound of the task's channel count (1, 2, 4 or 8) - it sizes the per-thread dW accumulators
template <int KMAX>
__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadBwdParams p) {
    extern __shared__ float sm[];  // [K][C] weights, then [K][C] + [K] block accumulators
    float* wS = sm;
    float* accS = sm + p.K * p.C;
    for (int i = threadIdx.x; i < p.K * p.C; i += blockDim.x) wS[i] = p.w[i];
    for (int i = threadIdx.x; i < p.K * p.C + p.K; i += blockDim.x) accS[i] = 0.f;
    __syncthreads();
    const int cg = p.C >> 3;
    const long long total = (long long)p.NB * p.S * cg;
    float dwl[KMAX][8];
    float dbl[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        dbl[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) dwl[k][j] = 0.f;
    }
    // all threads of a block keep the same channel group across iterations when the stride is a
    // multiple of cg, so round the stride
    const long long stride = ((long long)gridDim.x * blockDim.x / cg) * cg;
    const int g = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % cg);
    // (sample, voxel) of this thread's element: divided once, then advanced by the loop stride (64-bit divisions in the
    // loop body used to dominate this kernel)
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long vstep = stride / cg;
    long long nb = (i0 / cg) / p.S, s = (i0 / cg) % p.S;
    // four elements per iteration, all loads issued before the arithmetic: with one 16-byte load in flight per thread
    // this kernel ran at 1.6 TB/s (latency bound)
    constexpr int U = 4;
    for (long long i = i0; i < total; i += U * stride) {
        uint4 xv[U];
        float dv[U][KMAX];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long iu = i + u * stride;
            ok[u] = iu < total;
            while (s >= p.S) { s -= p.S; ++nb; }
            if (ok[u]) {
                xv[u] = ld_stream(reinterpret_cast<const uint4*>(p.x) + iu);
#pragma unroll
                for (int k = 0; k < KMAX; ++k)
                    dv[u][k] = k < p.K ? __ldg(p.dl + ((size_t)nb * p.K + k) * p.S + s) : 0.f;
            }
            s += vstep;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            float f[8], o[8];
            unpack8(xv[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < p.K) {
                    const float d = dv[u][k];
                    if (g == 0) dbl[k] += d;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j] = fmaf(d, wS[k * p.C + g * 8 + j], o[j]);
                        dwl[k][j] = fmaf(d, f[j], dwl[k][j]);
                    }
                }
            }
            reinterpret_cast<uint4*>(p.dx)[i + u * stride] = pack8(o);
        }
    }
    // Shared-memory fp32 atomics are compare-and-swap loops; 64 threads of a block share each accumulator.  When the
    // channel-group count divides 32 (every lane l of a warp then has group l % cg), the lanes of one group are
    // summed with shuffles first and one lane per group issues the atomic.
    const bool warpReduce = (cg <= 32) && (32 % cg == 0) && (blockDim.x % 32 == 0) && (stride % 32 == 0);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < p.K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = dwl[k][j];
                if (warpReduce) {
                    for (int o = cg; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane < cg) atomicAdd(&accS[k * p.C + g * 8 + j], v);
                } else {
                    atomicAdd(&accS[k * p.C + g * 8 + j], v);
                }
            }
            float b = dbl[k];
            if (warpReduce) {
                for (int o = cg; o < 32; o <<= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
                if (lane == 0) atomicAdd(&accS[p.K * p.C + k], b);
            } else if (g == 0) {
                atomicAdd(&accS[p.K * p.C + k], b);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.K * p.C; i += blockDim.x) atomicAdd(p.dw + i, accS[i]);
    if (p.db != nullptr)
        for (int i = threadIdx.x; i < p.K; i += blockDim.x) atomicAdd(p.db + i, accS[p.K * p.C + i]);
}

// ---------------------------------------------------------------------------------------
// Stem im2col: x NCDHW fp32 (the trainer's input layout) -> col [NB, D, H, W, Kp] bf16 with
// column index k = tap * Cin + ci (zero padded to Kp, a multiple of 16), so the stem
// convolution (encoder.py:81-86) and its weight gradient become 1x1x1 GEMMs.
// ---------------------------------------------------------------------------------------
struct Im2colParams {
    const float* x;
    bf16* col;
    int NB, Cin, D, H, W, kd, kh, kw, Kp;
};

__global__ void __launch_bounds__(256) stem_im2col_kernel(const Im2colParams p) {
    const int kg = p.Kp >> 3;
    const long long S = (long long)p.D * p.H * p.W;
    const long long total = (long long)p.NB * S * kg;
    const int K = p.kd * p.kh * p.kw * p.Cin;
    const int pd = (p.kd - 1) / 2, ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % kg);
        long long v = i / kg;
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
        reinterpret_cast<uint4*>(p.col)[i] = pack8(f);
    }
}

// Fixed-shape fast path (the reference stem: 3x3x3 taps on a 1-channel volume): one thread per voxel, all taps unrolled,
// the thread writes its whole Kp-wide row (coalesced 16-byte stores); the shape-generic kernel above spends its time in
// run-time divisions.
template <int KD, int KH, int KW, int CIN, int KP>
__global__ void __launch_bounds__(256) stem_im2col_fixed_kernel(const Im2colParams p) {
    const long long S = (long long)p.D * p.H * p.W;
    const long long total = (long long)p.NB * S;
    const size_t HW = (size_t)p.H * p.W;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        long long r = v;
        const int w = (int)(r % p.W); r /= p.W;
        const int h = (int)(r % p.H); r /= p.H;
        const int d = (int)(r % p.D); r /= p.D;
        const int nb = (int)r;
        float f[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) f[k] = 0.f;
#pragma unroll
        for (int td = 0; td < KD; ++td) {
            const int z = d + td - (KD - 1) / 2;
#pragma unroll
            for (int th = 0; th < KH; ++th) {
                const int y = h + th - (KH - 1) / 2;
                const bool rowOk = z >= 0 && z < p.D && y >= 0 && y < p.H;
#pragma unroll
                for (int tw = 0; tw < KW; ++tw) {
                    const int x = w + tw - (KW - 1) / 2;
                    if (rowOk && x >= 0 && x < p.W) {
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci)
                            f[((td * KH + th) * KW + tw) * CIN + ci] =
                                __ldg(p.x + (((size_t)nb * CIN + ci) * p.D + z) * HW + (size_t)y * p.W + x);
                    }
                }
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(p.col) + v * (KP / 8);
#pragma unroll
        for (int g = 0; g < KP / 8; ++g) {
            float q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = f[g * 8 + j];
            dst[g] = pack8(q);
        }
    }
}

// ---------------------------------------------------------------------------------------
// InstanceNorm finalisation (tiny, one thread per (n, c)).
//   fwd: sums[n][c] = (sum y, sum y^2) over S voxels (double) -> mean, rstd and the folded
//        scale = gamma * rstd, shift = beta - mean * scale consumed by norm_act_fwd_kernel.
//   bwd: red[n][c] = (G1 = sum g, Gy = sum g*y) -> coefficients of
//        dy = g*k1 + y*k2 + k3  (closed-form InstanceNorm backward, biased variance) and the
//        affine parameter gradients dgamma[c] = sum_n sum g*xhat, dbeta[c] = sum_n sum g.
// ---------------------------------------------------------------------------------------
struct FinalizeParams {
    const double* sums;   // [NB][C][2]  (or null when fsum/fsq are given)
    const float* fsum;    // [NB][C] fp32 statistics from the conv epilogue
    const float* fsq;
    const float* gamma;   // [C] or null
    const float* beta;    // [C] or null
    float* mean;          // [NB][C]
    float* rstd;          // [NB][C]
    float* scale;         // [NB][C]
    float* shift;         // [NB][C]
    int NB, C;
    double S;
    double eps;
};

__global__ void __launch_bounds__(256) in_finalize_fwd_kernel(const FinalizeParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.NB * p.C) return;
    const int c = i % p.C;
    const double s1 = p.sums ? p.sums[2 * (size_t)i] : (double)p.fsum[i];
    const double s2 = p.sums ? p.sums[2 * (size_t)i + 1] : (double)p.fsq[i];
    const double m = s1 / p.S;
    double var = s2 / p.S - m * m;
    if (var < 0.0) var = 0.0;
    const double r = 1.0 / sqrt(var + p.eps);
    const double ga = p.gamma ? (double)p.gamma[c] : 1.0;
    const double be = p.beta ? (double)p.beta[c] : 0.0;
    p.mean[i] = (float)m;
    p.rstd[i] = (float)r;
    p.scale[i] = (float)(ga * r);
    p.shift[i] = (float)(be - m * ga * r);
}

struct FinalizeBwdParams {
    const double* red;    // [NB][C][2] = (sum g, sum g*y)
    const float* mean;
    const float* rstd;
    const float* gamma;   // or null
    float* k1;            // [NB][C]
    float* k2;
    float* k3;
    float* dgamma;        // [C] accumulated (or null)
    float* dbeta;         // [C] accumulated (or null)
    int NB, C;
    double S;
};

__global__ void __launch_bounds__(256) in_finalize_bwd_kernel(const FinalizeBwdParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    double dg = 0.0, db = 0.0;
    for (int n = 0; n < p.NB; ++n) {
        const size_t i = (size_t)n * p.C + c;
        const double G1 = p.red[2 * i], Gy = p.red[2 * i + 1];
        const double m = p.mean[i], r = p.rstd[i];
        const double G2 = r * (Gy - m * G1);  // sum g * xhat
        const double a = (p.gamma ? (double)p.gamma[c] : 1.0) * r;
        p.k1[i] = (float)a;
        p.k2[i] = (float)(-a * r * G2 / p.S);
        p.k3[i] = (float)(-a * G1 / p.S + a * m * r * G2 / p.S);
        dg += G2;
        db += G1;
    }
    if (p.dgamma) p.dgamma[c] += (float)dg;
    if (p.dbeta) p.dbeta[c] += (float)db;
}

// ---------------------------------------------------------------------------------------
// Layout conversion at the network boundary: NCDHW fp32 <-> NDHWC bf16 (C % 8 == 0).
// One thread = one voxel x 8 channels; reads are coalesced along the voxel index per channel,
// writes are 16-byte vectors.
// ---------------------------------------------------------------------------------------
struct LayoutParams {
    const float* f32;  // [NB][C][S]
    bf16* cl;          // [NB][S][C]
    long long S;
    int NB, C;
};

__global__ void __launch_bounds__(256) ncdhw_to_cl_kernel(const LayoutParams p) {
    const int cg = p.C >> 3;
    const long long total = (long long)p.NB * cg * p.S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i % p.S;
        long long t = i / p.S;
        const int g = (int)(t % cg);
        const int nb = (int)(t / cg);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __ldg(p.f32 + ((size_t)nb * p.C + g * 8 + j) * p.S + s);
        *reinterpret_cast<uint4*>(p.cl + ((size_t)nb * p.S + s) * p.C + g * 8) = pack8(f);
    }
}

struct LayoutBackParams {
    const bf16* cl;
    float* f32;
    long long S;
    int NB, C;
};

__global__ void __launch_bounds__(256) cl_to_ncdhw_kernel(const LayoutBackParams p) {
    const int cg = p.C >> 3;
    const long long total = (long long)p.NB * cg * p.S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i % p.S;
        long long t = i / p.S;
        const int g = (int)(t % cg);
        const int nb = (int)(t / cg);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.cl + ((size_t)nb * p.S + s) * p.C + g * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) p.f32[((size_t)nb * p.C + g * 8 + j) * p.S + s] = f[j];
    }
}

// ---------------------------------------------------------------------------------------
// Weight (un)packing between the canonical parameter layout and the kernels' layouts, as tiled transposes
// through shared memory (both sides coalesced).  T = taps (kd*kh*kw), canonical w[co][ci][t] fp32.
//   pack  : out_f[t][co][ci] = bf16(w[co][ci][t])                       (fprop operand)
//           out_d[T-1-t][ci][co] = bf16(w[co][ci][t])                   (stride-1 data-gradient operand: flipped taps)
//   unpack: grad[co][ci][t] = dwp[t][co][ci]                            (weight-gradient result -> canonical)
// One block = a 32 x 32 (co, ci) tile with all taps (T <= 27).
// ---------------------------------------------------------------------------------------
struct WPackParams {
    const float* w;   // [Cout][Cin][T]
    bf16* out_f;      // [T][Cout][Cin] or null
    bf16* out_d;      // [T][Cin][Cout] (taps flipped) or null
    int Cout, Cin, T;
};

__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const WPackParams p) {
    extern __shared__ bf16 tileW[];   // [32 co][32 ci][T] (+1 padding on the ci stride)
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int T = p.T;
    const int pitch = 32 * T + 2;     // elements per co row of the tile (padding breaks bank alignment)
    // load: for each co of the tile, 32 ci x T floats are contiguous in the canonical layout
    for (int r = threadIdx.x >> 5; r < 32; r += 8) {
        const int co = co0 + r;
        if (co >= p.Cout) continue;
        const int nci = min(32, p.Cin - ci0);
        const float* src = p.w + ((size_t)co * p.Cin + ci0) * T;
        for (int i = threadIdx.x & 31; i < nci * T; i += 32) tileW[r * pitch + i] = __float2bfloat16_rn(__ldg(src + i));
    }
    __syncthreads();
    // fprop pack: [t][co][ci]: 32 consecutive ci per (t, co)
    if (p.out_f != nullptr) {
        for (int j = threadIdx.x >> 5; j < 32 * T; j += 8) {
            const int t = j / 32, r = j - t * 32;
            const int co = co0 + r, ci = ci0 + (threadIdx.x & 31);
            if (co < p.Cout && ci < p.Cin)
                p.out_f[((size_t)t * p.Cout + co) * p.Cin + ci] = tileW[r * pitch + (threadIdx.x & 31) * T + t];
        }
    }
    // dgrad pack: [T-1-t][ci][co]: 32 consecutive co per (t, ci)
    if (p.out_d != nullptr) {
        for (int j = threadIdx.x >> 5; j < 32 * T; j += 8) {
            const int t = j / 32, c = j - t * 32;
            const int ci = ci0 + c, co = co0 + (threadIdx.x & 31);
            if (co < p.Cout && ci < p.Cin)
                p.out_d[((size_t)(T - 1 - t) * p.Cin + ci) * p.Cout + co] = tileW[(threadIdx.x & 31) * pitch + c * T + t];
        }
    }
}

// Store phase shared by the pack kernels: bf16 tile [32 co][32 ci][T] (pitch 32*T + 2) in shared memory -> both operand
// layouts, two 64-byte rows per warp instruction.
__device__ __forceinline__ void wpack_store_tile(const bf16* tileW, bf16* out_f, bf16* out_d, int Cout, int Cin, int T,
                                                 int co0, int ci0) {
    const int pitch = 32 * T + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned short* tw = reinterpret_cast<const unsigned short*>(tileW);
    const int half = lane >> 4, l16 = lane & 15;
    // fprop pack [t][co][ci]: a warp instruction stores two (t, co) rows of 32 ci (2 x 64 B)
    if (out_f != nullptr) {
        for (int j = warp * 2 + half; j < 32 * T; j += 16) {
            const int t = j >> 5, r = j & 31;
            const int co = co0 + r;
            if (co < Cout) {
                const unsigned a = tw[r * pitch + (2 * l16) * T + t], b = tw[r * pitch + (2 * l16 + 1) * T + t];
                reinterpret_cast<uint32_t*>(out_f + ((size_t)t * Cout + co) * Cin + ci0)[l16] = a | (b << 16);
            }
        }
    }
    // dgrad pack [T-1-t][ci][co]: two (t, ci) rows of 32 co per warp instruction
    if (out_d != nullptr) {
        const bool pairOk = (Cout & 1) == 0;
        for (int j = warp * 2 + half; j < 32 * T; j += 16) {
            const int t = j >> 5, c = j & 31;
            const int ci = ci0 + c;
            const int coA = co0 + 2 * l16;
            bf16* row = out_d + ((size_t)(T - 1 - t) * Cin + ci) * Cout;
            if (pairOk && coA + 1 < Cout) {
                const unsigned a = tw[(2 * l16) * pitch + c * T + t], b = tw[(2 * l16 + 1) * pitch + c * T + t];
                *reinterpret_cast<uint32_t*>(row + coA) = a | (b << 16);
            } else {
                if (coA < Cout) reinterpret_cast<unsigned short*>(row)[coA] = tw[(2 * l16) * pitch + c * T + t];
                if (coA + 1 < Cout) reinterpret_cast<unsigned short*>(row)[coA + 1] = tw[(2 * l16 + 1) * pitch + c * T + t];
            }
        }
    }
}

// Vectorised variant for Cin % 32 == 0 (every layer but the stem): 16-byte global loads with the whole row batch in
// flight (the scalar kernel above keeps ~4 x 4 B per thread in flight and ran the 512 x 512 x 27 layers at 0.5 TB/s),
// 4-byte (bf16x2) shared-memory and global stores, two output rows per warp instruction.
__global__ void __launch_bounds__(256) pack_conv_weights_vec_kernel(const WPackParams p) {
    extern __shared__ bf16 tileW[];   // [32 co][32 ci][T] (+2 padding: odd word pitch => conflict-free column reads)
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int T = p.T;
    const int pitch = 32 * T + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n4 = 8 * T;             // float4 per (co, 32 ci) row
    for (int rr = 0; rr < 4; rr += 2) {
        float4 v[2][7];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int r = warp + 8 * (rr + k);
            const int co = co0 + r;
            const float4* src = reinterpret_cast<const float4*>(p.w + ((size_t)co * p.Cin + ci0) * T);
#pragma unroll
            for (int it = 0; it < 7; ++it) {
                const int idx = lane + 32 * it;
                if (co < p.Cout && idx < n4) {
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[k][it].x), "=f"(v[k][it].y), "=f"(v[k][it].z), "=f"(v[k][it].w) : "l"(src + idx));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int r = warp + 8 * (rr + k);
            if (co0 + r >= p.Cout) continue;
            uint32_t* dst = reinterpret_cast<uint32_t*>(tileW + r * pitch);
#pragma unroll
            for (int it = 0; it < 7; ++it) {
                const int idx = lane + 32 * it;
                if (idx < n4) {
                    dst[2 * idx] = pack_bf16(v[k][it].x, v[k][it].y);
                    dst[2 * idx + 1] = pack_bf16(v[k][it].z, v[k][it].w);
                }
            }
        }
    }
    __syncthreads();
    wpack_store_tile(tileW, p.out_f, p.out_d, p.Cout, p.Cin, T, co0, ci0);
}

// Data-gradient operand of a STRIDED conv as one pixel-shuffle gather (ops._merged_dgrad_plan): out[window tap
// (ud, uh, uw)][(rd, rh, rw, ci)][co] = w[co][ci][kd][kh][kw] where the kernel index of an axis follows from the output
// parity r and the window tap u of that axis (kidx[axis][r][u], -1 = this (parity, tap) pair has no kernel index: a
// zero block).  One launch per weight instead of the ~17 torch launches (flip, zero fill, one strided slice copy and one
// cast per parity class) that built the same tensor; reads are 4-byte gathers served by the L2 (the 27 taps of a
// (co, ci) pair share their sectors), stores are coalesced bf16 pairs along co.
struct DMergePackParams {
    const float* w;      // [Cout][Cin][K0][K1][K2]
    bf16* out;           // [nt0*nt1*nt2][s0*s1*s2*Cin][Cout]
    int Cout, Cin, K0, K1, K2;
    int nt0, nt1, nt2, s0, s1, s2;
    signed char kidx[3][2][4];
};

__global__ void __launch_bounds__(256) pack_dgrad_merged_kernel(const __grid_constant__ DMergePackParams p) {
    const int half = p.Cout >> 1;
    const long long total = (long long)p.nt0 * p.nt1 * p.nt2 * p.s0 * p.s1 * p.s2 * p.Cin * half;
    const int T = p.K0 * p.K1 * p.K2;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        long long r = o;
        const int co = (int)(r % half) * 2; r /= half;
        const int ci = (int)(r % p.Cin); r /= p.Cin;
        const int rw = (int)(r % p.s2); r /= p.s2;
        const int rh = (int)(r % p.s1); r /= p.s1;
        const int rd = (int)(r % p.s0); r /= p.s0;
        const int uw = (int)(r % p.nt2); r /= p.nt2;
        const int uh = (int)(r % p.nt1); r /= p.nt1;
        const int ud = (int)r;
        const int kd = p.kidx[0][rd][ud], kh = p.kidx[1][rh][uh], kw = p.kidx[2][rw][uw];
        uint32_t v = 0u;
        if (kd >= 0 && kh >= 0 && kw >= 0) {
            const size_t tap = (size_t)(kd * p.K1 + kh) * p.K2 + kw;
            const float a = __ldg(p.w + ((size_t)co * p.Cin + ci) * T + tap);
            const float b = __ldg(p.w + ((size_t)(co + 1) * p.Cin + ci) * T + tap);
            v = pack_bf16(a, b);
        }
        reinterpret_cast<uint32_t*>(p.out)[o] = v;
    }
}

struct WUnpackParams {
    const float* dwp;  // [T][A][B]
    float* grad;       // [A][B][T]
    int A, B, T;
};

// TT = compile-time tap count (27, 8) so that the load loop is fully unrolled (27 independent 128-byte rows in flight per
// warp instead of the ~4 the runtime-bound loop kept: 1.7 TB/s on the 512 x 512 x 27 gradients), 0 = runtime T.
template <int TT>
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const WUnpackParams p) {
    extern __shared__ float tileG[];  // [8 a][32 b][T] (+1): small tiles, many resident blocks (latency bound otherwise)
    const int a0 = blockIdx.y * 8, b0 = blockIdx.x * 32;
    const int T = TT ? TT : p.T;
    const int pitch = 32 * T + 1;
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = a0 + r, b = b0 + lane;
    if (TT) {
        float v[TT ? TT : 1];
#pragma unroll
        for (int t = 0; t < (TT ? TT : 1); ++t)
            v[t] = (a < p.A && b < p.B) ? __ldcs(p.dwp + ((size_t)t * p.A + a) * p.B + b) : 0.f;
#pragma unroll
        for (int t = 0; t < (TT ? TT : 1); ++t) tileG[r * pitch + lane * T + t] = v[t];
    } else {
        for (int t = 0; t < T; ++t)
            if (a < p.A && b < p.B) tileG[r * pitch + lane * T + t] = __ldg(p.dwp + ((size_t)t * p.A + a) * p.B + b);
    }
    __syncthreads();
    if (a < p.A) {
        const int nb = min(32, p.B - b0);
        float* dst = p.grad + ((size_t)a * p.B + b0) * T;
        if (TT && nb == 32) {
#pragma unroll
            for (int i = 0; i < (TT ? TT : 1); ++i) dst[lane + 32 * i] = tileG[r * pitch + lane + 32 * i];
        } else {
            for (int i = lane; i < nb * T; i += 32) dst[i] = tileG[r * pitch + i];
        }
    }
}

}  // namespace rb
