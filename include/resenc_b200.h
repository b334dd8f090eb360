/* resenc_b200.h — C ABI of the B200-native (sm_100a) ResEnc U-Net hot path.
 *
 * The reference (bruniss/multi-task-3d-resencoder-unet) is pure Python on top of torch.nn; it has
 * no FFI of its own.  Each entry point below therefore names the reference *call site* whose
 * arithmetic it replaces (path:line relative to the reference tree); the Python host mirror in
 * multi-task-3d-resencoder-unet_b200/builders/ binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes, no torch types; all pointers are DEVICE pointers unless noted
 *   - activations: channels-last NDHWC bf16, C % 8 == 0, 16-byte aligned; pre-norm conv outputs may be
 *     NDHWC fp32 instead (y_f32 / outF32 flags)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - every function returns 0 on success or a negative RbStatus; rb_last_error() gives the text
 *   - nothing is allocated persistently; workspaces are passed in by the caller
 *   - re-entrant across threads and streams
 */
#ifndef RESENC_B200_H
#define RESENC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum RbStatus {
    RB_OK = 0,
    RB_ERR_INVALID = -1,      /* bad descriptor / unsupported shape */
    RB_ERR_CUDA = -2,         /* CUDA runtime or driver error */
    RB_ERR_UNSUPPORTED = -3,  /* requested implementation cannot run this shape */
    RB_ERR_DEVICE = -4        /* device-side pipeline timeout recorded by a kernel */
};

enum RbConvImpl {
    RB_IMPL_AUTO = 0,     /* tcgen05 when the shape qualifies, else mma.sync */
    RB_IMPL_MMA_SYNC = 1, /* shape-generic warp-level tensor path (cross-check + odd shapes) */
    RB_IMPL_TCGEN05 = 2,  /* TMA + tcgen05.mma + TMEM; RB_ERR_UNSUPPORTED if the shape does not qualify */
    RB_IMPL_TCGEN05_SPLITK = 3, /* only returned by rb_conv_gather_plan: tcgen05 with taps split over the chip
                                   (needs the workspace: one fp32 slice per tap split; the finish pass sums the slices
                                   and takes the fused statistics) */
    RB_IMPL_TCGEN05_SLAB = 4    /* only returned by rb_conv_gather_plan: the z-marching tcgen05 kernel for 3x3x3
                                   stride-1 convolutions between 32-channel tensors (fused statistics available) */
};

const char* rb_last_error(void);
int rb_version(void);
/* Synchronises `stream`, returns RB_ERR_DEVICE (and clears the flag) if any kernel recorded a
 * pipeline timeout since the last call. */
int rb_device_error(void* stream);
/* Number of kernel launches issued through this library by the calling process (bench.py's
 * gpu_launches). */
long long rb_launch_count(void);
/* Pipeline cycle counters of CTA 0 of the last tcgen05 conv launch (only filled when the environment variable
 * RESENC_TC5_DEBUG has bit 3 set; profiling experiments, see profiles/). out16: host array of 16 u64. */
int rb_debug_counters(unsigned long long* out16);

/* ------------------------------------------------------------------------------------------
 * Gather convolution: one implicit GEMM covers every dense contraction of the network.
 *   out[m, n] = sum_{tap, c} A[src(m, tap), c] * W[tap][n][c]
 *   m over the output class grid (NB, OD, OH, OW); input coordinate i = o*istr + off + k,
 *   out-of-range reads as 0; A = one or two NDHWC tensors concatenated along C (virtual
 *   torch.cat); W packed bf16 [taps][Nout][Ctot].
 * Replaces:
 *   nn.Conv3d k3/k1, stride 1/2, pad (k-1)/2      builders/simple_conv_blocks.py:43-51
 *   its data gradient (flipped taps / 8 parity classes for stride 2)
 *   nn.ConvTranspose3d k == stride                builders/decoder.py:110-113 (mode 1: pixel shuffle)
 *   its data gradient (= k2 s2 p0 conv)
 *   torch.cat((up, skip), 1)                      builders/decoder.py:147 (nsrc == 2)
 * Output: direct (mode 0) to voxel (o*ostr + ooff) of an [NB, FD, FH, FW, outC0 (+outC1)] tensor
 * (two destinations split the channel range: used to route the concat gradient), or pixel
 * shuffle (mode 1): column = parity * psC + channel, voxel = o*ostr + parity offset.
 * ------------------------------------------------------------------------------------------ */
typedef struct RbConvDesc {
    int nsrc, srcC0, srcC1;
    int NB, ID, IH, IW;
    int tapD, tapH, tapW;
    int offD, offH, offW;
    int istrD, istrH, istrW;
    int OD, OH, OW;
    int Nout;
    int mode;
    int ostrD, ostrH, ostrW, ooffD, ooffH, ooffW;
    int FD, FH, FW;
    int outC0, outC1;
    int psC, psD, psH, psW;
    int impl;   /* RbConvImpl */
    int splitK; /* 0 = library heuristic, 1 = never, >1 = forced (mma.sync path only) */
    int outF32; /* 1: destinations are fp32 (pre-norm conv outputs keep the accumulator precision) */
} RbConvDesc;

/* fp32 workspace bytes needed by rb_conv_gather for this descriptor (0 when no split-K). */
size_t rb_conv_gather_workspace(const RbConvDesc* d);
/* stat_sum / stat_sq: optional fp32 [NB][Nout] accumulators (caller zeroes them) receiving the
 * per-(sample, channel) sum and sum of squares of the fp32 accumulators — the InstanceNorm
 * statistics (build_network_from_config.py:172) fused into the producing conv's epilogue.
 * Only the tcgen05 path fills them; returns RB_ERR_UNSUPPORTED if requested on mma.sync. */
int rb_conv_gather(const RbConvDesc* d, const void* src0, const void* src1, const void* w_packed,
                   void* out0, void* out1, float* stat_sum, float* stat_sq,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Which implementation rb_conv_gather will run for this descriptor: RB_IMPL_MMA_SYNC, RB_IMPL_TCGEN05 or
 * RB_IMPL_TCGEN05_SPLITK / RB_IMPL_TCGEN05_SLAB (negative status if a forced implementation cannot run it).
 * Fused statistics are produced by every tcgen05 variant (RB_IMPL_TCGEN05, _SLAB, _SPLITK: the finish pass of the
 * tap-split kernel takes them), not by RB_IMPL_MMA_SYNC. */
int rb_conv_gather_plan(const RbConvDesc* d);
/* 1 if the tcgen05 kernel can run this descriptor. */
int rb_conv_gather_tc5_supported(const RbConvDesc* d);

/* ------------------------------------------------------------------------------------------
 * Weight gradient:  dW[tap][a][b] += sum_m P[m][a] * Q[gather(m, tap)][b]      (fp32 atomics)
 *   conv  wgrad: P = dy on the output grid (a = Cout), Q = x (b = Cin; two sources allowed)
 *   convT wgrad: P = x on the input grid (a = Cin), Q = dy gathered at 2i + p (b = Cout)
 * Replaces the weight half of aten::convolution_backward for the call sites listed above
 * (train.py:224 `scaler.scale(loss).backward()`).  dw ([taps][PC][QC0+QC1] fp32) is fully initialised by the
 * call ("+=" describes the internal split accumulation; the caller need not zero it).
 * ------------------------------------------------------------------------------------------ */
typedef struct RbWgradDesc {
    int PC;
    int nq, QC0, QC1;
    int NB, GD, GH, GW;
    int QD, QH, QW;
    int tapD, tapH, tapW, offD, offH, offW, istrD, istrH, istrW;
    int splits; /* 0 = heuristic */
    int impl;   /* RbConvImpl: auto = tcgen05 when channel counts are multiples of 16 (and 32 in total for Q) */
} RbWgradDesc;
int rb_wgrad_gather(const RbWgradDesc* d, const void* P, const void* Q0, const void* Q1, float* dw, void* stream);

/* ------------------------------------------------------------------------------------------
 * InstanceNorm3d(affine=False|True, eps) + LeakyReLU + residual add + SE gate
 *   builders/simple_conv_blocks.py:58-64, builders/resblocks.py:106-114, DNA SqueezeExcite
 * rb_plane_reduce   per-(n,[w],c) reductions in double: kind 0 (sum y, sum y^2); kind 1
 *                   (sum g, sum g*y) with g = dz * lrelu'(z).  out: [NB][G][C][2], G = 1 or W.
 *                   The sign of z comes either from the stored activation `z` or, when z == NULL and
 *                   sign_scale / sign_shift ([NB][C], the forward's folded scale / shift) are given, from
 *                   fmaf(y, scale, shift) > 0 - layers without residual then never re-read z in backward.
 * rb_in_finalize_*  tiny: statistics -> folded scale/shift (fwd) or backward coefficients.
 * rb_norm_act_fwd   z = act(y * scale + shift + res)              one read of y (+res), one write
 * rb_norm_act_bwd   g = dz*act'(z); dres = g; dy = g*k1 + y*k2 + k3
 * ------------------------------------------------------------------------------------------ */
int rb_plane_reduce(int kind, const void* y, int y_f32, const void* dz, const void* z, const float* sign_scale,
                    const float* sign_shift, double* out, int NB, long long S, int C, int W, int perW, float slope,
                    void* stream);
/* statistics either as double sums[NB][C][2] (rb_plane_reduce) or as the fp32 fsum/fsq[NB][C] the tcgen05
 * conv epilogue accumulated */
int rb_in_finalize_fwd(const double* sums, const float* fsum, const float* fsq, const float* gamma, const float* beta,
                       float* mean, float* rstd, float* scale, float* shift, int NB, int C, double S, double eps,
                       void* stream);
int rb_in_finalize_bwd(const double* red, const float* mean, const float* rstd, const float* gamma,
                       float* k1, float* k2, float* k3, float* dgamma, float* dbeta,
                       int NB, int C, double S, void* stream);
int rb_norm_act_fwd(const void* y, int y_f32, const void* res, void* z, const float* scale, const float* shift,
                    int NB, long long S, int C, int W, int perW, int act, float slope, void* stream);
int rb_norm_act_bwd(const void* dz, const void* z, const float* sign_scale, const float* sign_shift, const void* y,
                    int y_f32, void* dy, void* dres, const float* k1, const float* k2, const float* k3,
                    int NB, long long S, int C, int W, int perW, int act, float slope, void* stream);

/* AvgPool3d(stride, stride) of the ResNet-D skip, builders/resblocks.py:92-95.  D,H,W = full-res dims. */
int rb_avgpool_fwd(const void* in, void* out, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream);
int rb_avgpool_bwd(const void* dout, void* din, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream);

/* Task head Conv3d(C, K<=8, 1, bias=True) (+ eval-mode activation), builders/decoder.py:131,151-152,
 * builders/build_network_from_config.py:6-18,322-323.  x NDHWC bf16 -> out NCDHW fp32.
 * act: 0 none, 1 sigmoid, 2 softmax over K. */
int rb_head_fwd(const void* x, const float* w, const float* b, float* out, int NB, long long S, int C, int K, int act, void* stream);
int rb_head_bwd(const void* x, const float* w, const float* dlogits, void* dx, float* dw, float* db,
                int NB, long long S, int C, int K, void* stream);

/* Inference tail of a task decoder as ONE pass (no autograd): the last conv block's InstanceNorm + LeakyReLU
 * (builders/simple_conv_blocks.py:58-60; + residual, builders/resblocks.py:113-114) followed by the 1x1x1 head
 * (builders/decoder.py:144-152) and the eval-mode activation (builders/build_network_from_config.py:322-323).
 * The decoder's last activation is never stored: y (pre-norm, y_mode 0 bf16 / 1 fp32 / 2 fp16, NDHWC) is read once,
 * out is NCDHW fp32 [NB][K][S].  scale / shift [NB][C] as produced by rb_in_finalize_fwd.  C / 8 must be a power of
 * two <= 32 (RB_ERR_UNSUPPORTED otherwise: the caller runs rb_norm_act_fwd + rb_head_fwd). */
int rb_norm_act_head_fwd(const void* y, int y_mode, const void* res, const float* scale, const float* shift,
                         const float* w, const float* b, float* out, int NB, long long S, int C, int K,
                         int act, float slope, int head_act, void* stream);

/* Stem im2col (builders/encoder.py:81-86, Cin not a multiple of 8): x NCDHW fp32 ->
 * col [NB,D,H,W,Kp] bf16, column = tap*Cin + ci, zero padded to Kp (multiple of 16). */
int rb_stem_im2col(const float* x, void* col, int NB, int Cin, int D, int H, int W, int kd, int kh, int kw, int Kp, void* stream);

/* ------------------------------------------------------------------------------------------
 * Split-precision ("bf16x3") inference tier: the north star's "1e-4 with fp32 accumulation" bound.  A value is
 * carried as hi = bf16(v), lo = bf16(v - hi); a voxel's activation row is the 3C channels [hi | lo | hi] and a
 * weight row [hi_w | hi_w | lo_w], so v*w ~= hi*hi_w + lo*hi_w + hi*lo_w is ONE rb_conv_gather call with
 * Cin' = 3*Cin and fp32 output.  These three entry points produce that layout; they replace the same reference
 * arithmetic as their bf16 twins (builders/simple_conv_blocks.py:58-64, builders/resblocks.py:92-95,106-114,
 * builders/encoder.py:81-86) with the normalise / gate / residual / LeakyReLU step evaluated in fp32.
 *   rb_split_apply        z[hi|lo|hi] = act(y * scale[n,c] + shift[n,c] + (res_hi + res_lo)); y [NB,S,C] fp32,
 *                         res / z [NB,S,3C] bf16, scale / shift [NB][C] fp32 or both NULL (identity)
 *   rb_avgpool_split      window mean of (hi + lo) in fp32, re-split; C = logical channels, rows are 3C wide
 *   rb_stem_im2col_split  col [NB,D,H,W,3*Kp] = [hi | lo | hi] of the rb_stem_im2col columns
 * ------------------------------------------------------------------------------------------ */
int rb_split_apply(const float* y, const void* res, void* z, const float* scale, const float* shift, int NB, long long S,
                   int C, int act, float slope, void* stream);
int rb_avgpool_split(const void* in, void* out, int NB, int D, int H, int W, int C, int sd, int sh, int sw, void* stream);
int rb_stem_im2col_split(const float* x, void* col, int NB, int Cin, int D, int H, int W, int kd, int kh, int kw, int Kp,
                         void* stream);

/* Weight (un)packing between the canonical parameter layout w[Cout][Cin][taps] fp32 (the reference's nn.Conv3d
 * weight, builders/simple_conv_blocks.py:43-51) and the kernels' operand layouts, tiled through shared memory:
 *   out_f[t][Cout][Cin] bf16 (fprop operand), out_d[taps-1-t][Cin][Cout] bf16 (stride-1 data-gradient operand);
 *   grad[A][B][taps] = dwp[taps][A][B] (rb_wgrad_gather result -> canonical gradient). Either output may be NULL. */
int rb_pack_conv_weights(const float* w, void* out_f, void* out_d, int Cout, int Cin, int taps, void* stream);
int rb_unpack_wgrad(const float* dwp, float* grad, int A, int B, int taps, void* stream);
/* Data-gradient operand of a strided conv (builders/simple_conv_blocks.py:43-51 with stride 2; autograd's
 * conv_backward in the reference) for the one-launch pixel-shuffle gather: out[window tap][(parity, ci)][co] bf16,
 * zero blocks where a (parity, tap) pair has no kernel index.  ntaps / stride: int[3]; kidx: signed char [3][2][4],
 * kidx[axis][parity][window tap] = kernel index or -1. */
int rb_pack_conv_dgrad_merged(const float* w, void* out, int Cout, int Cin, int K0, int K1, int K2, const int* ntaps,
                              const int* stride, const signed char* kidx, void* stream);

/* Optimiser step of the training loop: torch.nn.utils.clip_grad_norm_(model.parameters(), 3) followed by
 * torch.optim.AdamW.step() (train.py:79-83,227-228) as two multi-tensor passes over fp32 tensors.
 *   rb_grad_sumsq      *sumsq (device double, zeroed here) = sum over all listed gradients of g^2
 *   rb_adamw_clip_step  g' = g * min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (sumsq NULL: no clipping), then the decoupled-
 *                       weight-decay Adam update of p, m (exp_avg), v (exp_avg_sq) with bias correction for step *step
 *                       (device float, the 1-based count of THIS update); lr is a device float (LR schedules under CUDA
 *                       graphs).  Gradients are read, not modified.
 * `tensors` is a HOST array; p / m / v may be NULL for rb_grad_sumsq. */
typedef struct rb_opt_tensor {
    void* p;
    const void* g;
    void* m;
    void* v;
    long long n;
} rb_opt_tensor;
int rb_grad_sumsq(const rb_opt_tensor* tensors, int count, double* sumsq, void* stream);
int rb_adamw_clip_step(const rb_opt_tensor* tensors, int count, const float* lr, const float* step, const double* sumsq,
                       float max_norm, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* The same update for ONE conv weight w[Cout][Cin][taps] (Cin % 32 == 0, taps <= 27, 16-byte aligned tensors),
 * fused with rb_pack_conv_weights of the updated weight: out_f / out_d receive the bf16 operands of the next step, so
 * the pack kernel does not re-read the parameter.  Bit-identical to rb_adamw_clip_step + rb_pack_conv_weights. */
int rb_adamw_clip_pack_step(float* w, const float* g, float* m, float* v, void* out_f, void* out_d, int Cout, int Cin,
                            int taps, const float* lr, const float* step, const double* sumsq, float max_norm,
                            float beta1, float beta2, float eps, float weight_decay, void* stream);

/* Layout conversion at module boundaries: NCDHW fp32 <-> NDHWC bf16 (C % 8 == 0). */
int rb_ncdhw_to_cl(const float* src, void* dst, int NB, int C, long long S, void* stream);
int rb_cl_to_ncdhw(const void* src, float* dst, int NB, int C, long long S, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sliding-window blend, inference.py:135-157 (accumulate), :166-210 (finalise), :213-263 (cast),
 * inference/helpers.py:8-68 (gaussian weight).  fp32, no FMA contraction: weight == NULL is
 * bit-identical to the reference numpy loop.  One call = one patch; calls on one stream apply
 * in order (the reference's z-major order).
 * activation (fused inference.py:124-133): 0 none, 1 sigmoid, 2 softmax over C.
 * kind: 0 average -> uint8, 1 "normals" -> uint16.
 * ------------------------------------------------------------------------------------------ */
int rb_blend_accumulate(const float* pred, const float* weight, float* sum, float* wsum,
                        int C, int PZ, int PY, int PX, int VZ, int VY, int VX,
                        int z0, int y0, int x0, int activation, void* stream);
int rb_blend_finalize_cast(const float* sum, const float* wsum, void* out, float* favg,
                           long long V, int C, int kind, void* stream);
/* All targets of one patch in one launch (the loop over `infer_output_targets`, inference.py:139-157): per element the
 * same arithmetic as rb_blend_accumulate, 16-byte vector path when PX, VX, x0 are multiples of 4.  `targets` is a HOST
 * array; `wsum` is the one shared count / weight-sum volume (the reference's per-target counts are identical). */
typedef struct RbBlendTarget {
    const float* pred;   /* device [C][PZ][PY][PX] fp32, one patch */
    float* sum;          /* device [C][VZ][VY][VX] fp32 */
    int C;               /* 1..8 */
    int activation;      /* 0 none, 1 sigmoid, 2 softmax over C (inference.py:124-133) */
} RbBlendTarget;
int rb_blend_accumulate_multi(const RbBlendTarget* targets, int ntargets, const float* weight, float* wsum,
                              int PZ, int PY, int PX, int VZ, int VY, int VX, int z0, int y0, int x0, void* stream);
/* rb_blend_finalize_cast with the sum channels `sum_cstride` elements apart (finalise a z-range of a slab in place). */
int rb_blend_finalize_cast2(const float* sum, long long sum_cstride, const float* wsum, void* out, float* favg,
                            long long V, int C, int kind, void* stream);
int rb_blend_add(float* dst, const float* src, long long n, void* stream);

/* Patch extraction + per-patch standardisation, dataloading/inference_dataset.py:62-75.
 * vol: device uint8/uint16 [VZ][VY][VX]; stats: device double[2] scratch; out: fp32 [PZ][PY][PX]. */
int rb_extract_patch(const void* vol, int is_u16, int VZ, int VY, int VX, int z0, int y0, int x0,
                     int PZ, int PY, int PX, int standardize, double* stats, float* out, void* stream);

/* Batched form: `nb` (<= 16) patches per call, origins = HOST int[nb][3] (z0, y0, x0), stats device double[nb][2],
 * out fp32 [nb][PZ][PY][PX] (the batch the network reads, inference_dataset.py:62-75 + the DataLoader collate). */
int rb_extract_patches(const void* vol, int is_u16, int VZ, int VY, int VX, const int* origins, int nb,
                       int PZ, int PY, int PX, int standardize, double* stats, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused task losses on the fp32 NCDHW logits (training/losses/losses.py): BCEDiceLoss :307-318 (label-smoothed
 * BCE-with-logits :217-238 + DiceLoss :105-126 / compute_per_channel_dice :17-43) and MaskedCosineLoss :187-215.
 * *_reduce accumulates into `stats` (device double, zeroed by the caller): BCEDice [C][4] = sum bce, sum p*t,
 * sum p*p, sum t*t; cosine [2] = sum mask*cos, sum mask.  *_grad writes d(loss)/d(x) scaled by the device scalar
 * `grad_out`, using the same stats.
 * ------------------------------------------------------------------------------------------ */
int rb_loss_bce_dice_reduce(const float* logits, const float* target, double* stats, int NB, int C, long long S,
                            float smoothing, void* stream);
int rb_loss_bce_dice_grad(const float* logits, const float* target, const double* stats, const float* grad_out,
                          float* dlogits, int NB, int C, long long S, float alpha, float beta, float eps,
                          float smoothing, void* stream);
int rb_loss_cosine_reduce(const float* pred, const float* target, double* stats, int NB, long long S, void* stream);
int rb_loss_cosine_grad(const float* pred, const float* target, const double* stats, const float* grad_out,
                        float* dpred, int NB, long long S, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RESENC_B200_H */
