// Shared device helpers for the sm_100a kernels of the ResEnc U-Net hot path.
// Raw PTX wrappers (mbarrier, TMA, tcgen05/TMEM) - no CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

namespace rb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------
// Error reporting from device code: kernels with mbarrier waits never spin forever; on a
// timeout they record a code here and fall through so the launch drains.  The host API reads
// it back after synchronising (rb_check_device_error).
// ---------------------------------------------------------------------------------------
static __device__ int g_dev_error = 0;
// cycle counters of CTA 0 for pipeline experiments (RESENC_TC5_DEBUG & 8): see rb_debug_counters
static __device__ unsigned long long g_dbg[16];  // single translation unit (api.cu)

enum DevErr : int {
    DEVERR_NONE = 0,
    DEVERR_WAIT_FULL = 1,
    DEVERR_WAIT_EMPTY = 2,
    DEVERR_WAIT_TMEM_FULL = 3,
    DEVERR_WAIT_TMEM_EMPTY = 4,
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------- mbarrier ----------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait (about 0.5 s at 2 GHz).  Returns false and records `code` on timeout.  The spin loop touches
// nothing but the mbarrier, the SM clock and (rarely) a CTA-local flag in shared memory: a global-memory poll
// here would put an L2 round trip on every producer/consumer hand-off.  Once one wait of a CTA has timed out
// the flag makes that CTA's later waits give up at once, so a broken pipeline drains in about the timeout.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int code, uint32_t err_flag_smem) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            uint32_t flag;
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(flag) : "r"(err_flag_smem));
            if (flag != 0) return false;
            if (clock64() - t0 > 1000000000LL) {
                asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(err_flag_smem), "r"(1u) : "memory");
                atomicCAS(&g_dev_error, 0, code);
                return false;
            }
        }
    }
    return true;
}

// ------------------------------- TMA ---------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ------------------------------- tcgen05 / TMEM ----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 1 fp32 column.
__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t (&v)[1]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1").  Addresses/offsets are in bytes and
// must be multiples of 16.  layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // D format fp32
           | (1u << 7)                     // A format bf16
           | (1u << 10)                    // B format bf16
           | ((uint32_t)a_mn_major << 15)  // A major (0 = K-major)
           | ((uint32_t)b_mn_major << 16)  // B major
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t swizzle_layout_code(int swizzle_bytes) {
    return swizzle_bytes == 128 ? 2u : swizzle_bytes == 64 ? 4u : swizzle_bytes == 32 ? 6u : 0u;
}

// ------------------------------- fast division ----------------------------------------
// Division by a run-time constant as multiply-high + shift (valid for n < 2^31).  Integer division on
// the GPU is a ~30-instruction dependent sequence; the single producer / issuer threads of the
// tcgen05 kernels must not execute one per pipeline stage.
struct FastDiv {
    uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(int d_) {
    FastDiv f;
    f.d = (uint32_t)d_;
    if (d_ <= 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t lg = 0;
    while ((1u << lg) < (uint32_t)d_) ++lg;
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + (uint32_t)d_ - 1) / (uint32_t)d_);
    f.shr = p - 32;
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}
__device__ __forceinline__ void fdivmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
    q = fdiv(n, f);
    r = n - q * f.d;
}

// ------------------------------- misc --------------------------------------------------
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// fp16 with saturation (the pre-norm conv outputs can be stored as fp16: 3 more mantissa bits than bf16, half the bytes of
// fp32; cvt.satfinite clamps to +-65504 instead of producing inf).  out mode convention of the conv epilogues:
// 0 = bf16, 1 = fp32, 2 = fp16.
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool f16) { return f16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
__device__ __forceinline__ unsigned short cvt16(float x, bool f16) {
    if (f16) {
        unsigned short h;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
        return h;
    }
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float f16lo(uint32_t v) {
    float f;
    asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo; }" : "=f"(f) : "r"(v));
    return f;
}
__device__ __forceinline__ float f16hi(uint32_t v) {
    float f;
    asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi; }" : "=f"(f) : "r"(v));
    return f;
}

// fp32 x4 reduction into global memory (16-byte aligned), one L2 operation per lane.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rb
