// Fused task losses: the reference trainer's BCEDiceLoss and MaskedCosineLoss
// (training/losses/losses.py:307-318,217-238,105-126,17-43 and :187-215) on the fp32 NCDHW logits the head kernel
// writes.  One reduction pass per loss (double accumulators), a handful of scalar torch ops for the loss value, and
// one pass that writes d(loss)/d(logits) directly for the head backward - instead of ~40 full-resolution elementwise
// and reduction launches.  HBM-bound: 8 B read per element in each pass, 4 B written in the gradient pass.
#pragma once
#include "common.cuh"

namespace rb {

struct LossParams {
    const float* x;      // logits / predicted normals [NB][C][S]
    const float* t;      // targets                    [NB][C][S]
    double* stats;       // BCEDice: [C][4] = sum bce, sum p*t, sum p*p, sum t*t ; cosine: [2] = sum mask*cos, sum mask
    float* dx;           // gradient output (grad kernels)
    const float* gout;   // upstream gradient of the scalar loss (device)
    long long S;
    int NB, C;
    float alpha, beta, eps, smoothing;
};

__device__ __forceinline__ double block_sum_double(double v, double* red) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;   // valid in thread 0
}

// grid (blocks over S, C, NB)
__global__ void __launch_bounds__(256) loss_bce_dice_reduce_kernel(const LossParams p) {
    __shared__ double red[8];
    const int c = blockIdx.y, nb = blockIdx.z;
    const float* x = p.x + ((size_t)nb * p.C + c) * p.S;
    const float* t = p.t + ((size_t)nb * p.C + c) * p.S;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.S; i += (long long)gridDim.x * blockDim.x) {
        const float z = __ldg(x + i), y = __ldg(t + i);
        const float ys = y * (1.f - 2.f * p.smoothing) + p.smoothing;
        const float e = __expf(-fabsf(z));
        a0 += fmaxf(z, 0.f) - z * ys + log1pf(e);
        const float pr = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        a1 += pr * y;
        a2 += pr * pr;
        a3 += y * y;
    }
    const double s0 = block_sum_double((double)a0, red);
    const double s1 = block_sum_double((double)a1, red);
    const double s2 = block_sum_double((double)a2, red);
    const double s3 = block_sum_double((double)a3, red);
    if (threadIdx.x == 0) {
        atomicAdd(p.stats + c * 4 + 0, s0);
        atomicAdd(p.stats + c * 4 + 1, s1);
        atomicAdd(p.stats + c * 4 + 2, s2);
        atomicAdd(p.stats + c * 4 + 3, s3);
    }
}

// d/dz [ alpha * mean bce + beta * (1 - mean_c 2 I_c / max(den_c, eps)) ]
__global__ void __launch_bounds__(256) loss_bce_dice_grad_kernel(const LossParams p) {
    const int c = blockIdx.y, nb = blockIdx.z;
    const float* x = p.x + ((size_t)nb * p.C + c) * p.S;
    const float* t = p.t + ((size_t)nb * p.C + c) * p.S;
    float* dx = p.dx + ((size_t)nb * p.C + c) * p.S;
    const float g = __ldg(p.gout);
    const double I = p.stats[c * 4 + 1], den = p.stats[c * 4 + 2] + p.stats[c * 4 + 3];
    const bool clamped = den < (double)p.eps;
    const double dn = clamped ? (double)p.eps : den;
    // D = 2 I / dn ;  dD/dp = 2 t / dn - (clamped ? 0 : 4 I p / dn^2)
    const float kT = (float)(2.0 / dn), kP = clamped ? 0.f : (float)(4.0 * I / (dn * dn));
    const float wb = g * p.alpha / (float)((double)p.NB * p.C * (double)p.S);
    const float wd = -g * p.beta / (float)p.C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.S; i += (long long)gridDim.x * blockDim.x) {
        const float z = __ldg(x + i), y = __ldg(t + i);
        const float ys = y * (1.f - 2.f * p.smoothing) + p.smoothing;
        const float e = __expf(-fabsf(z));
        const float pr = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        dx[i] = wb * (pr - ys) + wd * (kT * y - kP * pr) * pr * (1.f - pr);
    }
}

// MaskedCosineLoss: 3-channel vectors.  grid (blocks over S, 1, NB)
__global__ void __launch_bounds__(256) loss_cosine_reduce_kernel(const LossParams p) {
    __shared__ double red[8];
    const int nb = blockIdx.z;
    const float* x = p.x + (size_t)nb * 3 * p.S;
    const float* t = p.t + (size_t)nb * 3 * p.S;
    float a0 = 0.f, a1 = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.S; i += (long long)gridDim.x * blockDim.x) {
        const float x0 = __ldg(x + i), x1 = __ldg(x + p.S + i), x2 = __ldg(x + 2 * p.S + i);
        const float t0 = __ldg(t + i), t1 = __ldg(t + p.S + i), t2 = __ldg(t + 2 * p.S + i);
        const float tn = sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
        if (tn > 1e-6f) {
            const float xn = fmaxf(sqrtf(x0 * x0 + x1 * x1 + x2 * x2), 1e-8f);
            const float u0 = x0 / xn, u1 = x1 / xn, u2 = x2 / xn;
            const float un = fmaxf(sqrtf(u0 * u0 + u1 * u1 + u2 * u2), 1e-8f);
            a0 += (u0 * t0 + u1 * t1 + u2 * t2) / (un * fmaxf(tn, 1e-8f));
            a1 += 1.f;
        }
    }
    const double s0 = block_sum_double((double)a0, red);
    const double s1 = block_sum_double((double)a1, red);
    if (threadIdx.x == 0) {
        atomicAdd(p.stats + 0, s0);
        atomicAdd(p.stats + 1, s1);
    }
}

// loss = 1 - S / (M + 1e-8):  d/dx = -(t^ - (u . t^) u) / |x| / (M + 1e-8) on unmasked voxels
__global__ void __launch_bounds__(256) loss_cosine_grad_kernel(const LossParams p) {
    const int nb = blockIdx.z;
    const float* x = p.x + (size_t)nb * 3 * p.S;
    const float* t = p.t + (size_t)nb * 3 * p.S;
    float* dx = p.dx + (size_t)nb * 3 * p.S;
    const float k = -__ldg(p.gout) / (float)(p.stats[1] + 1e-8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.S; i += (long long)gridDim.x * blockDim.x) {
        const float x0 = __ldg(x + i), x1 = __ldg(x + p.S + i), x2 = __ldg(x + 2 * p.S + i);
        const float t0 = __ldg(t + i), t1 = __ldg(t + p.S + i), t2 = __ldg(t + 2 * p.S + i);
        const float tn = sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        const float xn = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
        if (tn > 1e-6f && xn > 1e-8f) {
            const float u0 = x0 / xn, u1 = x1 / xn, u2 = x2 / xn;
            const float h0 = t0 / tn, h1 = t1 / tn, h2 = t2 / tn;
            const float c = u0 * h0 + u1 * h1 + u2 * h2;
            const float s = k / xn;
            d0 = s * (h0 - c * u0);
            d1 = s * (h1 - c * u1);
            d2 = s * (h2 - c * u2);
        }
        dx[i] = d0;
        dx[p.S + i] = d1;
        dx[2 * p.S + i] = d2;
    }
}

}  // namespace rb
