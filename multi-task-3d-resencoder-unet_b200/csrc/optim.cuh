// Optimiser step of the training hot path: global-norm gradient clip + AdamW as two multi-tensor passes.
//
// Reference (file:line in /root/reference): train.py:79-83 builds torch.optim.AdamW(lr, weight_decay), train.py:227
// clips with torch.nn.utils.clip_grad_norm_(model.parameters(), 3) and train.py:228 steps.  Through PyTorch that is
// three multi-tensor sweeps per step over the 235 M parameters of the 128^3 network (profiles/r2_launches_bench_steps2.csv):
// the per-tensor norms (0.94 GB read), the in-place scaling of every gradient (1.9 GB) and the fused AdamW update
// (6.6 GB) - 1.9 ms of a 25 ms step.  Here the clip coefficient is applied to the gradient as it is read by the update
// (the scaling sweep disappears) and both passes keep several 16-byte loads in flight per thread.
//
// Arithmetic (fp32, the order of torch's fused AdamW kernel):
//     coef  = min(1, max_norm / (sqrt(sum g^2) + 1e-6))          clip_grad_norm_
//     g     = g * coef
//     p     = p - lr * wd * p
//     m     = m + (1 - b1) * (g - m)                              lerp
//     v     = b2 * v + (1 - b2) * g * g
//     p     = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once
#include "common.cuh"

namespace rb {

static constexpr int OPT_MAX_TENSORS = 48;   // per launch: 48 x (4 pointers + count) = 1.9 KB of kernel parameters

struct OptTensorList {
    float* p[OPT_MAX_TENSORS];
    const float* g[OPT_MAX_TENSORS];
    float* m[OPT_MAX_TENSORS];
    float* v[OPT_MAX_TENSORS];
    long long n[OPT_MAX_TENSORS];
};

struct OptHyper {
    const float* lr;        // device scalar
    const float* step;      // device scalar: t of THIS update (already incremented)
    const double* sumsq;    // device scalar: sum of squares of all gradients, or null (no clipping)
    float beta1, beta2, eps, weight_decay, max_norm;
};

__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// sum of squares of every listed gradient -> one double (atomicAdd per block).  grid = (blocks, tensors)
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const __grid_constant__ OptTensorList L, double* out) {
    const int t = blockIdx.y;
    const long long n = L.n[t];
    const float* g = L.g[t];
    float acc = 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const long long nv = n >> 2;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        long long i = tid;
        for (; i + 3 * stride < nv; i += 4 * stride) {           // four 16-byte loads in flight per thread
            const float4 x0 = ld_f4_stream(g + 4 * i), x1 = ld_f4_stream(g + 4 * (i + stride));
            const float4 x2 = ld_f4_stream(g + 4 * (i + 2 * stride)), x3 = ld_f4_stream(g + 4 * (i + 3 * stride));
            a0 += x0.x * x0.x + x0.y * x0.y + x0.z * x0.z + x0.w * x0.w;
            a1 += x1.x * x1.x + x1.y * x1.y + x1.z * x1.z + x1.w * x1.w;
            a2 += x2.x * x2.x + x2.y * x2.y + x2.z * x2.z + x2.w * x2.w;
            a3 += x3.x * x3.x + x3.y * x3.y + x3.z * x3.z + x3.w * x3.w;
        }
        for (; i < nv; i += stride) {
            const float4 x0 = ld_f4_stream(g + 4 * i);
            a0 += x0.x * x0.x + x0.y * x0.y + x0.z * x0.z + x0.w * x0.w;
        }
        acc = (a0 + a1) + (a2 + a3);
        for (long long j = (nv << 2) + tid; j < n; j += stride) acc += g[j] * g[j];
    } else {
        for (long long j = tid; j < n; j += stride) acc += g[j] * g[j];
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0 && s != 0.f) atomicAdd(out, (double)s);
    }
}

struct AdamCoef {
    float coef, lr_wd, one_m_b1, b2, one_m_b2, step_size, inv_bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_coef_init(const OptHyper& h, AdamCoef& cs) {
    const float lr = *h.lr;
    const double t = (double)*h.step;
    const double bc1 = 1.0 - pow((double)h.beta1, t);
    const double bc2 = 1.0 - pow((double)h.beta2, t);
    float coef = 1.f;
    if (h.sumsq != nullptr) {
        const float total = (float)sqrt(*h.sumsq);
        coef = fminf(h.max_norm / (total + 1e-6f), 1.f);
    }
    cs.coef = coef;
    cs.lr_wd = lr * h.weight_decay;
    cs.one_m_b1 = 1.f - h.beta1;
    cs.b2 = h.beta2;
    cs.one_m_b2 = 1.f - h.beta2;
    cs.step_size = (float)((double)lr / bc1);
    cs.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    cs.eps = h.eps;
}

// explicit rounding intrinsics: the compiler may not contract or reassociate them, so the two kernels that inline this
// function (multi-tensor update, update fused with the operand packs) produce the same bits
__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamCoef& c) {
    g = __fmul_rn(g, c.coef);
    p = __fmaf_rn(-c.lr_wd, p, p);
    m = __fmaf_rn(c.one_m_b1, __fsub_rn(g, m), m);
    v = __fmaf_rn(__fmul_rn(c.one_m_b2, g), g, __fmul_rn(c.b2, v));
    const float denom = __fmaf_rn(__fsqrt_rn(v), c.inv_bc2_sqrt, c.eps);
    p = __fmaf_rn(-c.step_size, __fdiv_rn(m, denom), p);
}

__device__ __forceinline__ void adamw_vec(float4& p, const float4& g, float4& m, float4& v, const AdamCoef& c) {
    adamw_elem(p.x, g.x, m.x, v.x, c);
    adamw_elem(p.y, g.y, m.y, v.y, c);
    adamw_elem(p.z, g.z, m.z, v.z, c);
    adamw_elem(p.w, g.w, m.w, v.w, c);
}

// grid = (blocks, tensors); a block strides over its tensor, two float4 per stream in flight per thread
__global__ void __launch_bounds__(256) adamw_clip_kernel(const __grid_constant__ OptTensorList L, const OptHyper h) {
    __shared__ AdamCoef cs;
    if (threadIdx.x == 0) adam_coef_init(h, cs);
    __syncthreads();
    const AdamCoef c = cs;
    const int t = blockIdx.y;
    const long long n = L.n[t];
    float* p = L.p[t];
    const float* g = L.g[t];
    float* m = L.m[t];
    float* v = L.v[t];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
    if (vec) {
        const long long nv = n >> 2;
        long long i = tid;
        for (; i + stride < nv; i += 2 * stride) {
            const long long j = i + stride;
            float4 p0 = ld_f4(p + 4 * i), m0 = ld_f4(m + 4 * i), v0 = ld_f4(v + 4 * i);
            const float4 g0 = ld_f4_stream(g + 4 * i);
            float4 p1 = ld_f4(p + 4 * j), m1 = ld_f4(m + 4 * j), v1 = ld_f4(v + 4 * j);
            const float4 g1 = ld_f4_stream(g + 4 * j);
            adamw_vec(p0, g0, m0, v0, c);
            *reinterpret_cast<float4*>(p + 4 * i) = p0;
            *reinterpret_cast<float4*>(m + 4 * i) = m0;
            *reinterpret_cast<float4*>(v + 4 * i) = v0;
            adamw_vec(p1, g1, m1, v1, c);
            *reinterpret_cast<float4*>(p + 4 * j) = p1;
            *reinterpret_cast<float4*>(m + 4 * j) = m1;
            *reinterpret_cast<float4*>(v + 4 * j) = v1;
        }
        for (; i < nv; i += stride) {
            float4 p0 = ld_f4(p + 4 * i), m0 = ld_f4(m + 4 * i), v0 = ld_f4(v + 4 * i);
            const float4 g0 = ld_f4_stream(g + 4 * i);
            adamw_vec(p0, g0, m0, v0, c);
            *reinterpret_cast<float4*>(p + 4 * i) = p0;
            *reinterpret_cast<float4*>(m + 4 * i) = m0;
            *reinterpret_cast<float4*>(v + 4 * i) = v0;
        }
        for (long long j = (nv << 2) + tid; j < n; j += stride) adamw_elem(p[j], g[j], m[j], v[j], c);
    } else {
        for (long long j = tid; j < n; j += stride) adamw_elem(p[j], g[j], m[j], v[j], c);
    }
}


// ---------------------------------------------------------------------------------------
// AdamW update of ONE conv weight fused with the bf16 operand packs of the next step (elementwise.cuh, "Weight
// (un)packing"): the update already streams the fp32 weight through registers, so the [taps][Cout][Cin] fprop operand
// and the flipped [taps][Cin][Cout] data-gradient operand are written from there instead of by a second kernel that
// re-reads the weight (pack_conv_weights_vec_kernel: 0.94 GB read + 0.94 GB written per step on the 128^3 network).
// One block = a 32 x 32 (co, ci) tile with all taps; Cin % 32 == 0, taps <= 27.  The arithmetic is adamw_elem, so the
// result is bit-identical to adamw_clip_kernel followed by the pack kernel.
// MEASURED SLOWER in the whole step (interleaved 30-step runs of bench.py on one B200: 25.60 / 25.81 ms with the fused
// kernel against 25.39 / 25.33 ms with the multi-tensor update + pack kernel): 32 x 32 tiles give the 512-channel layers
// 256 blocks on 148 SMs at two blocks per SM, and a block alternates between its streaming phase and its transposing
// store phase instead of overlapping them.  Opt-in only (ClippedAdamW(manage_packs=True)).
// ---------------------------------------------------------------------------------------
struct AdamPackParams {
    float* w;         // [Cout][Cin][T]
    const float* g;
    float* m;
    float* v;
    bf16* out_f;      // [T][Cout][Cin] or null
    bf16* out_d;      // [T][Cin][Cout] (taps flipped) or null
    int Cout, Cin, T;
};

__global__ void __launch_bounds__(256) adamw_pack_kernel(const AdamPackParams p, const OptHyper h) {
    extern __shared__ bf16 tileA[];   // [32 co][32 ci][T] (+2 padding), as in pack_conv_weights_vec_kernel
    __shared__ AdamCoef cs;
    if (threadIdx.x == 0) adam_coef_init(h, cs);
    __syncthreads();
    const AdamCoef c = cs;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int T = p.T;
    const int pitch = 32 * T + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n4 = 8 * T;             // float4 per (co, 32 ci) row
    for (int rr = 0; rr < 4; ++rr) {
        const int r = warp + 8 * rr;
        const int co = co0 + r;
        if (co >= p.Cout) continue;
        const size_t row = ((size_t)co * p.Cin + ci0) * T;
        uint32_t* dst = reinterpret_cast<uint32_t*>(tileA + r * pitch);
#pragma unroll
        for (int it0 = 0; it0 < 8; it0 += 4) {           // four float4 per stream in flight: 16 x 16 B per thread
            float4 w4[4], g4[4], m4[4], v4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = lane + 32 * (it0 + k);
                if (idx < n4) {
                    w4[k] = ld_f4(p.w + row + 4 * idx);
                    m4[k] = ld_f4(p.m + row + 4 * idx);
                    v4[k] = ld_f4(p.v + row + 4 * idx);
                    g4[k] = ld_f4_stream(p.g + row + 4 * idx);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = lane + 32 * (it0 + k);
                if (idx < n4) {
                    adamw_vec(w4[k], g4[k], m4[k], v4[k], c);
                    *reinterpret_cast<float4*>(p.w + row + 4 * idx) = w4[k];
                    *reinterpret_cast<float4*>(p.m + row + 4 * idx) = m4[k];
                    *reinterpret_cast<float4*>(p.v + row + 4 * idx) = v4[k];
                    dst[2 * idx] = pack_bf16(w4[k].x, w4[k].y);
                    dst[2 * idx + 1] = pack_bf16(w4[k].z, w4[k].w);
                }
            }
        }
    }
    __syncthreads();
    wpack_store_tile(tileA, p.out_f, p.out_d, p.Cout, p.Cin, T, co0, ci0);
}

}  // namespace rb
