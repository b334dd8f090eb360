// Sliding-window blend kernels (HBM-bound): patch accumulation into the running sum / weight
// volumes and the fused finalise + cast.  fp32 throughout, IEEE round-to-nearest with no FMA
// contraction, so that with weight == nullptr ("uniform") the result is bit-identical to the
// reference's numpy loop on the same predictions.
//
// Reference arithmetic (file:line in /root/reference):
//   accumulate   inference.py:135-157   sum[..., z0:z0+P, y0:, x0:] += pred ; count[...] += 1
//   finalise     inference.py:166-210   "normals" (c == 3): v /= sqrt(v0^2+v1^2+v2^2) + 1e-8 where count > 0
//                                        others: v /= count where count > 0
//   cast         inference.py:213-263   normals: clip((v + 1) / 2 * 65535, 0, 65535) -> uint16 (truncating)
//                                        others : clip(v * 255, 0, 255) -> uint8 (truncating)
//   gaussian     inference/helpers.py:8-68 (importance map, applied only when weight != nullptr)
#pragma once
#include "common.cuh"

namespace rb {

struct BlendParams {
    const float* pred;    // [C][PZ][PY][PX]   one patch (already activated)
    const float* weight;  // [PZ][PY][PX] or null (uniform)
    float* sum;           // [C][VZ][VY][VX]   slab-local volume
    float* wsum;          // [VZ][VY][VX] or null (caller keeps one shared weight volume)
    int C, PZ, PY, PX;
    int VZ, VY, VX;
    int z0, y0, x0;       // patch origin in slab coordinates (may be negative / overhang: clipped)
    int activation;       // 0 none, 1 sigmoid, 2 softmax over C (applied to pred on the fly)
};

// One launch per patch, launches in stream order => each voxel's additions happen in the
// reference's z-major patch order, no atomics, deterministic.
// thread = 4 consecutive x of one (z, y) row when PX % 4 == 0 and x0 % 4 == 0 (vector path).
__global__ void __launch_bounds__(256) blend_accumulate_kernel(const BlendParams p) {
    const long long PS = (long long)p.PZ * p.PY * p.PX;
    const long long VS = (long long)p.VZ * p.VY * p.VX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < PS; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % p.PX);
        long long t = i / p.PX;
        const int py = (int)(t % p.PY);
        const int pz = (int)(t / p.PY);
        const int z = p.z0 + pz, y = p.y0 + py, x = p.x0 + px;
        if (z < 0 || z >= p.VZ || y < 0 || y >= p.VY || x < 0 || x >= p.VX) continue;
        const long long v = ((long long)z * p.VY + y) * p.VX + x;
        const float w = p.weight ? __ldg(p.weight + i) : 1.f;
        if (p.activation == 2) {
            float mx = -INFINITY;
            for (int c = 0; c < p.C; ++c) mx = fmaxf(mx, __ldg(p.pred + c * PS + i));
            float den = 0.f;
            for (int c = 0; c < p.C; ++c) den += expf(__ldg(p.pred + c * PS + i) - mx);
            for (int c = 0; c < p.C; ++c) {
                const float a = expf(__ldg(p.pred + c * PS + i) - mx) / den;
                float* d = p.sum + c * VS + v;
                *d = __fadd_rn(*d, p.weight ? __fmul_rn(a, w) : a);
            }
        } else {
            for (int c = 0; c < p.C; ++c) {
                float a = __ldg(p.pred + c * PS + i);
                if (p.activation == 1) a = 1.f / (1.f + expf(-a));
                float* d = p.sum + c * VS + v;
                *d = __fadd_rn(*d, p.weight ? __fmul_rn(a, w) : a);
            }
        }
        if (p.wsum) p.wsum[v] = __fadd_rn(p.wsum[v], w);
    }
}

struct FinalizeCastParams {
    const float* sum;   // [C][V]
    const float* wsum;  // [V]
    void* out;          // uint8 [C][V] or uint16 [C][V]
    float* favg;        // optional fp32 [C][V] finalised values (null to skip)
    long long V;
    int C;
    int kind;           // 0: average -> uint8 ; 1: normals (C == 3 renormalise, else untouched) -> uint16
};

__device__ __forceinline__ float clipf(float v, float lo, float hi) {
    // numpy.clip semantics (NaN propagates; irrelevant after the cast but kept identical)
    return v < lo ? lo : (v > hi ? hi : v);
}

__global__ void __launch_bounds__(256) blend_finalize_cast_kernel(const FinalizeCastParams p) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < p.V; v += (long long)gridDim.x * blockDim.x) {
        const float cnt = __ldg(p.wsum + v);
        const bool mask = cnt > 0.f;
        if (p.kind == 1) {
            float s[3];
            if (p.C == 3) {
                s[0] = __ldg(p.sum + v); s[1] = __ldg(p.sum + p.V + v); s[2] = __ldg(p.sum + 2 * p.V + v);
                // sum_block[0]**2 + sum_block[1]**2 + sum_block[2]**2, left to right, each op rounded
                const float mag = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(s[0], s[0]), __fmul_rn(s[1], s[1])),
                                                                  __fmul_rn(s[2], s[2]))), 1e-8f);
                if (mask) { s[0] = __fdiv_rn(s[0], mag); s[1] = __fdiv_rn(s[1], mag); s[2] = __fdiv_rn(s[2], mag); }
            }
            for (int c = 0; c < p.C; ++c) {
                const float a = (p.C == 3) ? s[c] : __ldg(p.sum + c * p.V + v);
                if (p.favg) p.favg[c * p.V + v] = a;
                float q = __fmul_rn(__fdiv_rn(__fadd_rn(a, 1.0f), 2.0f), 65535.0f);
                q = clipf(q, 0.f, 65535.f);
                reinterpret_cast<uint16_t*>(p.out)[c * p.V + v] = (uint16_t)q;
            }
        } else {
            for (int c = 0; c < p.C; ++c) {
                float a = __ldg(p.sum + c * p.V + v);
                if (mask) a = __fdiv_rn(a, cnt);
                if (p.favg) p.favg[c * p.V + v] = a;
                float q = clipf(__fmul_rn(a, 255.0f), 0.f, 255.f);
                reinterpret_cast<uint8_t*>(p.out)[c * p.V + v] = (uint8_t)q;
            }
        }
    }
}

// Neighbour-slab merge for the z-sharded sweep: dst[0:n] += src[0:n] (fp32, exact order:
// lower slab's partial first, so the result equals the single-GPU sum when each voxel got at
// most one contribution per slab... in general sums are associative only up to rounding and the
// multi-GPU result is compared with a tolerance).
__global__ void __launch_bounds__(256) blend_add_kernel(float* dst, const float* src, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __fadd_rn(dst[i], src[i]);
}

// ---------------------------------------------------------------------------------------
// Patch extraction + per-patch standardisation on the device
// (dataloading/inference_dataset.py:62-75: scale /255 or /65535, then pytorch3dunet
//  Standardize(channelwise=False): (m - mean) / clip(std, 1e-10); std is the population std).
// Pass 1 reduces sum / sum^2 in double; pass 2 writes the standardised fp32 NCDHW patch.
// ---------------------------------------------------------------------------------------
struct ExtractParams {
    const void* vol;   // [VZ][VY][VX] uint8 or uint16
    int is_u16;
    int VZ, VY, VX;
    int z0, y0, x0, PZ, PY, PX;
    double* stats;     // [2]  (sum, sum of squares) of the scaled patch
    float* out;        // [PZ][PY][PX]
    int standardize;
};

__device__ __forceinline__ float extract_load(const ExtractParams& p, long long i) {
    const int px = (int)(i % p.PX);
    long long t = i / p.PX;
    const int py = (int)(t % p.PY);
    const int pz = (int)(t / p.PY);
    const long long v = ((long long)(p.z0 + pz) * p.VY + (p.y0 + py)) * p.VX + (p.x0 + px);
    if (p.is_u16) return __fdiv_rn((float)reinterpret_cast<const uint16_t*>(p.vol)[v], 65535.0f);
    return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(p.vol)[v], 255.0f);
}

__global__ void __launch_bounds__(256) patch_stats_kernel(const ExtractParams p) {
    const long long PS = (long long)p.PZ * p.PY * p.PX;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < PS; i += (long long)gridDim.x * blockDim.x) {
        const double a = (double)extract_load(p, i);
        s1 += a;
        s2 += a * a;
    }
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        atomicAdd(p.stats, a);
        atomicAdd(p.stats + 1, b);
    }
}

__global__ void __launch_bounds__(256) patch_write_kernel(const ExtractParams p) {
    const long long PS = (long long)p.PZ * p.PY * p.PX;
    float mean = 0.f, inv = 1.f;
    if (p.standardize) {
        const double m = p.stats[0] / (double)PS;
        double var = p.stats[1] / (double)PS - m * m;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        if (sd < 1e-10) sd = 1e-10;
        mean = (float)m;
        inv = (float)(1.0 / sd);
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < PS; i += (long long)gridDim.x * blockDim.x) {
        const float a = extract_load(p, i);
        p.out[i] = p.standardize ? (a - mean) * inv : a;
    }
}


// ---------------------------------------------------------------------------------------
// Round 2: all targets of one patch in ONE launch, four x per thread (16-byte loads / stores) when the
// geometry allows, 32-bit index arithmetic (two unsigned divisions per four voxels instead of three 64-bit
// div/mod per voxel).  Per element the arithmetic is the one of blend_accumulate_kernel above (same rounding,
// same order), so the uniform blend stays bit-identical to the reference loop.
// Algorithmic traffic per patch voxel at c_tot = 4 (sheet 1 + normals 3), fp32 predictions:
//   read pred 4*c_tot = 16 B, read-modify-write sum 8*c_tot = 32 B, RMW weight-sum 8 B, weight map 4 B (L2 resident)
//   = 56 B (+4) per patch voxel.
// ---------------------------------------------------------------------------------------
constexpr int BLEND_MAX_TARGETS = 8;
constexpr int BLEND_MAX_C = 8;

struct BlendTarget {
    const float* pred;   // [C][PZ][PY][PX] one patch
    float* sum;          // [C][VZ][VY][VX]
    int C;
    int activation;      // 0 none, 1 sigmoid, 2 softmax over C
};

struct BlendMultiParams {
    BlendTarget t[BLEND_MAX_TARGETS];
    int nt;
    const float* weight;  // [PZ][PY][PX] or null
    float* wsum;          // [VZ][VY][VX] or null
    int PZ, PY, PX;
    int VZ, VY, VX;
    int z0, y0, x0;
};

template <int VEC> struct BlendVec;
template <> struct BlendVec<1> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = *p; }
    static __device__ __forceinline__ void ldro(const float* p, float* v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float* p, const float* v) { *p = v[0]; }
};
template <> struct BlendVec<4> {
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void ldro(const float* p, float* v) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(p));     // predictions are read exactly once
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

template <int VEC>
__global__ void __launch_bounds__(256) blend_accumulate_multi_kernel(const BlendMultiParams p) {
    const unsigned rowVecs = (unsigned)p.PX / VEC;
    const unsigned total = (unsigned)p.PZ * (unsigned)p.PY * rowVecs;
    const size_t PS = (size_t)p.PZ * p.PY * p.PX;
    const size_t VS = (size_t)p.VZ * p.VY * p.VX;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned row = i / rowVecs;
        const unsigned xv = i - row * rowVecs;
        const unsigned pz = row / (unsigned)p.PY;
        const unsigned py = row - pz * (unsigned)p.PY;
        const int z = p.z0 + (int)pz, y = p.y0 + (int)py, x = p.x0 + (int)(xv * VEC);
        if (z < 0 || z >= p.VZ || y < 0 || y >= p.VY) continue;
        if (VEC == 1 && (x < 0 || x >= p.VX)) continue;      // the vector path is only launched for patches inside the row
        const size_t pi = (size_t)row * p.PX + xv * VEC;
        const size_t v = ((size_t)z * p.VY + y) * p.VX + x;
        float w[VEC];
        if (p.weight) BlendVec<VEC>::ld(p.weight + pi, w);
        else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) w[k] = 1.f;
        }
        for (int ti = 0; ti < p.nt; ++ti) {
            const BlendTarget& t = p.t[ti];
            float a[BLEND_MAX_C][VEC];
#pragma unroll
            for (int c = 0; c < BLEND_MAX_C; ++c)
                if (c < t.C) BlendVec<VEC>::ldro(t.pred + c * PS + pi, a[c]);
            if (t.activation == 2) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < BLEND_MAX_C; ++c)
                        if (c < t.C) mx = fmaxf(mx, a[c][k]);
                    float den = 0.f;
#pragma unroll
                    for (int c = 0; c < BLEND_MAX_C; ++c)
                        if (c < t.C) den += expf(a[c][k] - mx);
#pragma unroll
                    for (int c = 0; c < BLEND_MAX_C; ++c)
                        if (c < t.C) a[c][k] = expf(a[c][k] - mx) / den;
                }
            } else if (t.activation == 1) {
#pragma unroll
                for (int c = 0; c < BLEND_MAX_C; ++c)
                    if (c < t.C) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) a[c][k] = 1.f / (1.f + expf(-a[c][k]));
                    }
            }
#pragma unroll
            for (int c = 0; c < BLEND_MAX_C; ++c)
                if (c < t.C) {
                    float* d = t.sum + c * VS + v;
                    float s[VEC];
                    BlendVec<VEC>::ld(d, s);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) s[k] = __fadd_rn(s[k], p.weight ? __fmul_rn(a[c][k], w[k]) : a[c][k]);
                    BlendVec<VEC>::st(d, s);
                }
        }
        if (p.wsum) {
            float s[VEC];
            BlendVec<VEC>::ld(p.wsum + v, s);
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[k] = __fadd_rn(s[k], w[k]);
            BlendVec<VEC>::st(p.wsum + v, s);
        }
    }
}

// Finalise + cast, four voxels per thread, sum channels `cstride` elements apart (a z-range of a slab is finalised in
// place, no gathering copy).  Arithmetic identical to blend_finalize_cast_kernel.
// Algorithmic traffic per output voxel at c_tot = 4: read 4*c_tot + 4 = 20 B, write 1 + 2*3 = 7 B => 27 B.
struct FinalizeCast2Params {
    const float* sum;     // channel c at sum + c * cstride
    const float* wsum;
    void* out;            // [C][V]
    float* favg;          // optional [C][V]
    long long V, cstride;
    int C, kind;
};

template <int VEC>
__global__ void __launch_bounds__(256) blend_finalize_cast2_kernel(const FinalizeCast2Params p) {
    const long long nvec = p.V / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long v = i * VEC;
        float cnt[VEC];
        BlendVec<VEC>::ldro(p.wsum + v, cnt);
        float s[BLEND_MAX_C][VEC];
#pragma unroll
        for (int c = 0; c < BLEND_MAX_C; ++c)
            if (c < p.C) BlendVec<VEC>::ldro(p.sum + c * p.cstride + v, s[c]);
        if (p.kind == 1) {
            if (p.C == 3) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float mag = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(s[0][k], s[0][k]), __fmul_rn(s[1][k], s[1][k])),
                                                                      __fmul_rn(s[2][k], s[2][k]))), 1e-8f);
                    if (cnt[k] > 0.f) {
                        s[0][k] = __fdiv_rn(s[0][k], mag); s[1][k] = __fdiv_rn(s[1][k], mag); s[2][k] = __fdiv_rn(s[2][k], mag);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < BLEND_MAX_C; ++c)
                if (c < p.C) {
                    if (p.favg) BlendVec<VEC>::st(p.favg + c * p.V + v, s[c]);
                    uint16_t q[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        q[k] = (uint16_t)clipf(__fmul_rn(__fdiv_rn(__fadd_rn(s[c][k], 1.0f), 2.0f), 65535.0f), 0.f, 65535.f);
                    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + c * p.V + v;
                    if (VEC == 4) *reinterpret_cast<uint2*>(o) = make_uint2((unsigned)q[0] | ((unsigned)q[VEC > 1 ? 1 : 0] << 16),
                                                                             (unsigned)q[VEC > 2 ? 2 : 0] | ((unsigned)q[VEC > 3 ? 3 : 0] << 16));
                    else o[0] = q[0];
                }
        } else {
#pragma unroll
            for (int c = 0; c < BLEND_MAX_C; ++c)
                if (c < p.C) {
                    uint8_t q[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        if (cnt[k] > 0.f) s[c][k] = __fdiv_rn(s[c][k], cnt[k]);
                        q[k] = (uint8_t)clipf(__fmul_rn(s[c][k], 255.0f), 0.f, 255.f);
                    }
                    if (p.favg) BlendVec<VEC>::st(p.favg + c * p.V + v, s[c]);
                    uint8_t* o = reinterpret_cast<uint8_t*>(p.out) + c * p.V + v;
                    if (VEC == 4) *reinterpret_cast<unsigned*>(o) = (unsigned)q[0] | ((unsigned)q[VEC > 1 ? 1 : 0] << 8) |
                                                                    ((unsigned)q[VEC > 2 ? 2 : 0] << 16) | ((unsigned)q[VEC > 3 ? 3 : 0] << 24);
                    else o[0] = q[0];
                }
        }
    }
}

// Batched patch extraction: all patches of a forward batch in two launches (statistics, write) instead of three
// calls per patch.  Patch origins travel in the parameter block (no device-side index table, no host copy).
constexpr int EXTRACT_MAX_BATCH = 16;
struct ExtractBatchParams {
    const void* vol;
    int is_u16;
    int VZ, VY, VX;
    int PZ, PY, PX;
    int nb;
    int z0[EXTRACT_MAX_BATCH], y0[EXTRACT_MAX_BATCH], x0[EXTRACT_MAX_BATCH];
    double* stats;      // [nb][2]
    float* out;         // [nb][PZ][PY][PX]
    int standardize;
};

__device__ __forceinline__ float extract_load_b(const ExtractBatchParams& p, int b, unsigned i) {
    const unsigned px = i % (unsigned)p.PX;
    const unsigned t = i / (unsigned)p.PX;
    const unsigned py = t % (unsigned)p.PY;
    const unsigned pz = t / (unsigned)p.PY;
    const size_t v = ((size_t)(p.z0[b] + pz) * p.VY + (p.y0[b] + py)) * p.VX + (p.x0[b] + px);
    if (p.is_u16) return __fdiv_rn((float)reinterpret_cast<const uint16_t*>(p.vol)[v], 65535.0f);
    return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(p.vol)[v], 255.0f);
}

__global__ void __launch_bounds__(256) patch_stats_batch_kernel(const ExtractBatchParams p) {
    const int b = blockIdx.y;
    const unsigned PS = (unsigned)p.PZ * p.PY * p.PX;
    double s1 = 0.0, s2 = 0.0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < PS; i += gridDim.x * blockDim.x) {
        const double a = (double)extract_load_b(p, b, i);
        s1 += a;
        s2 += a * a;
    }
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; c += sh[1][w]; }
        atomicAdd(p.stats + 2 * b, a);
        atomicAdd(p.stats + 2 * b + 1, c);
    }
}

__global__ void __launch_bounds__(256) patch_write_batch_kernel(const ExtractBatchParams p) {
    const int b = blockIdx.y;
    const unsigned PS = (unsigned)p.PZ * p.PY * p.PX;
    float mean = 0.f, inv = 1.f;
    if (p.standardize) {
        const double m = p.stats[2 * b] / (double)PS;
        double var = p.stats[2 * b + 1] / (double)PS - m * m;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var);
        if (sd < 1e-10) sd = 1e-10;
        mean = (float)m;
        inv = (float)(1.0 / sd);
    }
    float* out = p.out + (size_t)b * PS;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < PS; i += gridDim.x * blockDim.x) {
        const float a = extract_load_b(p, b, i);
        out[i] = p.standardize ? (a - mean) * inv : a;
    }
}

}  // namespace rb
