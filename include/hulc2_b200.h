/* hulc2_b200 -- C-ABI of the B200-native HULC++ low-level policy step.
 *
 * The reference (mees/hulc2) has no native interface for this path: its hot path is ~1.3 kLoC of
 * PyTorch nn.Modules.  Each entry point below therefore replaces the body of a
 * reference Python function (cited as file:line relative to the reference root); the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions (SURVEY.md section 8b "C-ABI layer"):
 *  - all pointers are DEVICE pointers borrowed for the duration of the call; fp32 unless noted;
 *  - every call is asynchronous on `stream`, never synchronises, never allocates; scratch memory is
 *    a caller-provided workspace;
 *  - returns 0 on success, a negative HULC2_E* code otherwise (hulc2_last_error() gives the text);
 *  - `precision`: 0 = fp32 CUDA-core path (1e-5 parity), 1 = bf16 tcgen05 tensor-core path.
 */
#ifndef HULC2_B200_H
#define HULC2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* hulc2_stream_t; /* == cudaStream_t */

const char* hulc2_last_error(void);
int hulc2_version(void);
/* number of kernels this library has launched in this process (evidence for bench.py's gpu_launches) */
unsigned long long hulc2_launch_count(void);
/* 1 when the running device is sm_100 (B200) and the tcgen05 path is usable. */
int hulc2_device_supports_tcgen05(void);

/* ------------------------------------------------------------------ dense contractions
 * C[m,n] = epilogue( alpha * sum_k A(m,k) * B(n,k) ),  A(m,k) = A[R(m) + k*a_ks], B(n,k) = B[n*b_rs + k*b_ks]
 * R(m) = m*a_rs, or (m / a_inner)*a_rs_outer + (m % a_inner)*a_rs_inner when a_inner > 0 (same for C rows).
 * epilogue: +bias[n]; +add[m,n]; +C (accumulate); relu; zero where mask[m,n] <= 0; dropout keep (u8) * keep_scale.
 * Serves nn.Linear forward / input-grad / weight-grad everywhere on the path, e.g.
 * plan_proposal_net.py:42-47, goal_encoders.py:29-34, logistic_decoder_rnn.py:269-274.
 * A16 / B16 / C16 / ld16 (precision 1 only, all optional): row-major bf16 mirrors of the operands -- element i of A16 is
 * bf16(A[i]) -- so the strides and offsets above address them unchanged; written by hulc2_f32_to_bf16 or by a producing
 * GEMM through its C16.  With both mirrors present and 16-byte-aligned rows the contraction runs on the TMA-fed tcgen05
 * kernel (csrc/gemm_tma_sm100.cu); otherwise the gather kernel converts the fp32 operands on the fly.  A and/or B may be
 * NULL when the mirror is the only copy (then the strides describe the mirror and a layout the TMA path cannot take is
 * an error, not a fallback).  C16 (row stride ld16) additionally receives the epilogue result as bf16: the next layer's
 * operand.
 * rowsum (precision 1, TMA path only, optional): float[M] that receives sum_k A(m,k) of the bf16 operand -- in a weight
 * gradient dW = g^T x this is the bias gradient sum_rows g (the reference's autograd of nn.Linear's bias), computed by one
 * more narrow MMA against a constant ones operand instead of a separate column-sum pass.  Requesting it on a problem the
 * TMA path cannot take is an error. */
typedef struct {
  int M, N, K;
  const float* A; long long a_rs, a_ks; int a_inner; long long a_rs_outer, a_rs_inner;
  const float* B; long long b_rs, b_ks;
  float* C; long long ldc; int c_inner; long long c_rs_outer, c_rs_inner;
  const float* bias;
  const float* add; long long ld_add;
  const float* mask; long long ld_mask;
  const unsigned char* keep; long long ld_keep; float keep_scale;
  int relu, accumulate;
  float alpha;
  int precision;
  void* workspace; long long workspace_bytes;
  const void* A16; const void* B16;
  void* C16; long long ld16;
  float* rowsum;
} hulc2_gemm_args;
int hulc2_gemm(const hulc2_gemm_args* a, hulc2_stream_t stream);
/* number of contractions served by the TMA-fed kernel so far (tests assert the fast path was taken) */
unsigned long long hulc2_tma_gemm_count(void);
/* dst[i] = bf16(src[i]) (round to nearest even), the operand mirrors of hulc2_gemm_args */
int hulc2_f32_to_bf16(const float* src, void* dst, long long n, hulc2_stream_t stream);
/* strided fp32 rows -> compact bf16 rows: dst[r*ld_dst + c] = bf16(src[r*ld_src + c]) for c < cols */
int hulc2_f32_to_bf16_2d(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols,
                         hulc2_stream_t stream);

/* ------------------------------------------------------------------ convolutions (implicit GEMM)
 * vision_network.py:38-48 and vision_network_gripper.py:11-26 (valid padding, square stride).
 * Activations produced by this library are NHWC; the first conv reads the caller's NCHW frames.
 *  fwd  : y[F,OH,OW,Cout] = relu?(conv(x,w)+bias);  x NCHW with w OIHW (in_nhwc=0) or x NHWC with w OHWI (in_nhwc=1)
 *  wgrad: dw (same layout as w) (+)= sum_pixels dy^T im2col(x);   needs workspace for split-K
 *  dgrad: dx[F,H,W,C] (NHWC) = gather(dy, w_hwoi[KH,KW,Cout,C]), zeroed where xmask <= 0 (xmask may be null) */
typedef struct {
  int F, C, H, W, Cout, KH, KW, stride, in_nhwc;
  const float* x; const float* w; const float* bias; float* y; int relu;
  const float* dy; float* dw; float* dx; const float* xmask; int accumulate;
  int precision;
  void* workspace; long long workspace_bytes;
} hulc2_conv_args;
int hulc2_conv2d_fwd(const hulc2_conv_args* a, hulc2_stream_t stream);
int hulc2_conv2d_wgrad(const hulc2_conv_args* a, hulc2_stream_t stream);
int hulc2_conv2d_dgrad(const hulc2_conv_args* a, hulc2_stream_t stream);
/* weight layout shuffles between the state_dict layout (OIHW) and the kernel layouts.
 * dir 0: OIHW -> OHWI, 1: OHWI -> OIHW (accumulate optional), 2: OIHW -> HWOI */
int hulc2_permute_conv_weight(const float* src, float* dst, int O, int I, int KH, int KW, int dir, int accumulate,
                              hulc2_stream_t stream);

/* ------------------------------------------------------------------ bf16 conv trunk (sm_100a, precision 1)
 * Persistent warp-specialised tcgen05 implicit GEMMs over bf16 NHWC activations (csrc/conv_sm100.cu); replaces the
 * conv stacks of vision_network.py:38-48 / vision_network_gripper.py:11-26 and their autograd.
 *  pack_frames : fp32 NCHW frames [F,C,H,W] -> bf16 [F, H/4, W/4, 16C], channel (ci,a,b) = x[f,ci,4I+a,4J+b]; the
 *                k8/s4 first conv becomes a k2/s1 conv over 16C channels (describe it that way in hulc2_convb_args).
 *  pack_weight : fp32 OIHW -> bf16.  mode 0: [Cout][(kh,kw,ci)];  mode 1: first conv over packed frames, dims are the
 *                ORIGINAL [Cout,Cin,8,8]: [Cout][(dI,dJ,ci,a,b)];  mode 2: input-gradient operand, one [Cin][(a,b,co)]
 *                matrix per stride-parity class (ph,pw), classes concatenated in (ph,pw) order.
 *  fwd   : y[F,OH,OW,Cout] = relu?(conv(x, w) + bias), x/y bf16 NHWC, w from pack_weight mode 0/1.
 *  dgrad : dx[F,H,W,C] = conv^T(dy, w) zeroed where xmask <= 0 (xmask = forward activation x, may be null); w mode 2.
 *  wgrad : dw (fp32, OIHW; dw_layout 1 = ORIGINAL [Cout, C/16, 4KH, 4KW] of a packed-frames conv) and db[Cout]
 *          from x and dy; workspace >= 148 * (KH*KW*C + 128) * 64 * 4 bytes.
 * hulc2_convb_supported() says whether a layer shape is served (C % 8 == 0, KH*KW*C % 64 == 0, Cout in {32,64}, stride <= 2).
 * Kernel selection is internal (csrc/conv_halo_sm100.cu: one TMA box per tile + shifted UMMA descriptors for stride-1 / the
 * stride-2 layers of this trunk; csrc/conv_sm100.cu: cp.async gather otherwise; HULC2_CONV_HALO=0 forces the gather kernels).
 * The halo path reads pixels narrower than 64 channels as 64-element rows: `x` (fwd, C < 64) must be followed by >= 128
 * readable bytes (any content); fwd passes `workspace` (>= Cout*KH*KW*128 bytes) for the re-tiled weights of that case. */
typedef struct {
  int F, C, H, W, Cout, KH, KW, stride;
  const void* x; const void* w; const float* bias; void* y; int relu;
  const void* dy; void* dx; const void* xmask;
  float* dw; float* db; int dw_layout;
  void* workspace; long long workspace_bytes;
  /* ReLU sign bits of an activation tensor, 1 bit per element in NHWC element order (bit i of the tensor = bit i % 8 of byte
   * i / 8; 2-byte aligned, >= ceil(elements / 16) * 2 bytes).  fwd (relu != 0): optional OUTPUT -- the halo kernels also write
   * (y > 0) here and set mask_bits_written = 1 (the gather kernel leaves it 0).  dgrad: optional INPUT, used INSTEAD of reading the
   * bf16 activations `xmask` (16x fewer mask bytes; the halo kernels only -- keep passing xmask for the gather kernel). */
  void* mask_bits; int mask_bits_written;
} hulc2_convb_args;
int hulc2_convb_supported(int C, int Cout, int KH, int KW, int stride);
int hulc2_pack_frames_bf16(const float* x, void* xs, int F, int C, int H, int W, hulc2_stream_t stream);
int hulc2_convb_pack_weight(const float* w_oihw, void* wp, int Cout, int Cin, int KH, int KW, int mode, int stride,
                            hulc2_stream_t stream);
int hulc2_convb_fwd(const hulc2_convb_args* a, hulc2_stream_t stream);
int hulc2_convb_dgrad(const hulc2_convb_args* a, hulc2_stream_t stream);
int hulc2_convb_wgrad(const hulc2_convb_args* a, hulc2_stream_t stream);
/* SpatialSoftmax / nn.Flatten boundaries of the bf16 trunk (x, dx bf16 NHWC; everything else fp32) */
int hulc2_spatial_softmax_fwd_bf16(const void* x, const float* x_map, const float* y_map, const float* temperature,
                                   float* out, int F, int HW, int C, hulc2_stream_t stream);
int hulc2_spatial_softmax_bwd_bf16(const void* x, const float* x_map, const float* y_map, const float* temperature,
                                   const float* dout, void* dx, float* dtemperature, int F, int HW, int C, int relu_mask,
                                   hulc2_stream_t stream);
/* Same, with the per-(frame, channel) softmax statistics (max, 1/sum; fp32 [F, 2C]) saved by the forward and consumed -- together
 * with the forward output `out` -- by the backward, which then makes ONE pass over the frame instead of three.  Served by the
 * shared-memory bf16 kernel only: check hulc2_spatial_softmax_stats_supported(HW, C) first (else HULC2_ENOTIMPL). */
int hulc2_spatial_softmax_stats_supported(int HW, int C);
int hulc2_spatial_softmax_fwd_bf16_stats(const void* x, const float* x_map, const float* y_map, const float* temperature,
                                         float* out, float* stats, int F, int HW, int C, hulc2_stream_t stream);
int hulc2_spatial_softmax_bwd_bf16_stats(const void* x, const float* x_map, const float* y_map, const float* temperature,
                                         const float* out, const float* stats, const float* dout, void* dx,
                                         float* dtemperature, int F, int HW, int C, int relu_mask, hulc2_stream_t stream);
int hulc2_nhwc_bf16_to_nchw(const void* src, float* dst, int F, int HW, int C, hulc2_stream_t stream);
int hulc2_nchw_to_nhwc_bf16(const float* src, void* dst, int F, int HW, int C, const void* mask, hulc2_stream_t stream);

/* ------------------------------------------------------------------ datamodule: uint8 frames -> trunk operand (csrc/frames.cu)
 * Replaces, per camera, what the reference's CPU dataloader workers do for every window:
 *   window slice + pad_with_repetition   hulc2/datasets/base_dataset.py:121-163, npz_dataset.py:117-143
 *   HWC uint8 -> CHW                     hulc2/datasets/utils/episode_utils.py:61-86 (process_rgb)
 *   RandomShiftsAug(pad)                 hulc2/utils/transforms.py:85-106  (integer crop of the replicate-padded frame)
 *   ScaleImageTensor + Normalize(.5,.5)  hulc2/utils/transforms.py:8-19, conf/datamodule/transforms/rand_shift.yaml:2-10
 * store  : uint8 [N,H,W,C] frames (an episode store resident in HBM, or the B*S frames of one batch);
 * win_start[B] (int64, first frame of window b) / win_len[B] (int32 valid steps, null = S): output frame f=(b,t) reads
 *          store[win_start[b] + min(t, win_len[b]-1)]; win_start null = identity (frame f reads store[f]);
 * shift  : int32 [F,2] = (dx,dy) per output frame, each in [-pad,pad] (the reference's randint draw minus pad), null = none:
 *          out[y,x] = in[clamp(y+dy,0,H-1), clamp(x+dx,0,W-1)];  value = ((u8/255) - 0.5)/0.5 in fp32, reference op order.
 * pack_bf16 writes the packed-frames layout of hulc2_pack_frames_bf16 directly; to_f32 writes fp32 NCHW [F,C,H,W].
 * Both pack functions need 128 bytes of slack behind the last packed pixel of `xs` and ZERO them (conv1's halo path reads
 * 48-channel pixels as 64-element rows: the last row ends in the slack, where only finite values are harmless).
 * window_gather_f32: out[b,t,:] = store[win_start[b]+t, :] (fp32 [N,D]) for t < win_len[b]; padded steps by mode:
 *          0 repeat the last valid row, 1 zeros, 2 zeros except the last component which repeats (relative actions). */
int hulc2_frames_u8_pack_bf16(const void* store, const long long* win_start, const int* win_len, const int* shift, void* xs,
                              int F, int S, int C, int H, int W, hulc2_stream_t stream);
int hulc2_frames_u8_to_f32(const void* store, const long long* win_start, const int* win_len, const int* shift, float* out,
                           int F, int S, int C, int H, int W, hulc2_stream_t stream);
int hulc2_window_gather_f32(const float* store, const long long* win_start, const int* win_len, float* out, int B, int S, int D,
                            int mode, hulc2_stream_t stream);

/* ------------------------------------------------------------------ small data movement
 * copy2d: dst[r*ldd + c] (+)= src[r*lds + c];  colsum: out[c] (+)= sum_r x[r*ld + c] (bias gradients);
 * nhwc<->nchw per-frame transposes (nn.Flatten order of nature_cnn, vision_network_gripper.py:22-23). */
int hulc2_copy2d(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols, int accumulate,
                 hulc2_stream_t stream);
/* dst[d1, d0, :D2] (+)= src[d0*src_s0 + d1*src_s1 + :D2]: batch-major [B,S,*] <-> time-major [S,B,*] row shuffle */
int hulc2_transpose01(const float* src, long long src_s0, long long src_s1, float* dst, long long dst_ld, int D0, int D1,
                      int D2, int accumulate, hulc2_stream_t stream);
int hulc2_fill(float* dst, long long n, float value, hulc2_stream_t stream);
int hulc2_axpy(const float* x, float* y, long long n, float a, hulc2_stream_t stream); /* y += a*x */
/* Scalar loss combine (hulc2/models/hulc2.py:243,426-430: total = action + kl_beta*kl + clip_beta*clip and the logged
 * means): out[0] = sum_i w[i] * xs[i][0] for n <= 8 device scalars (xs, w: HOST arrays read at call time), and its
 * gradient fan-out out[i] = w[i] * g[0]. */
int hulc2_weighted_sum(const float* const* xs, const float* w, int n, float* out, hulc2_stream_t stream);
int hulc2_weighted_fanout(const float* g, const float* w, int n, float* out, hulc2_stream_t stream);
int hulc2_colsum(const float* x, long long ld, long long rows, int cols, float* out, int accumulate, void* workspace,
                 long long workspace_bytes, hulc2_stream_t stream);
int hulc2_relu_mask(const float* dy, const float* y, float* dz, long long n, hulc2_stream_t stream);
int hulc2_nhwc_to_nchw(const float* src, float* dst, int F, int HW, int C, hulc2_stream_t stream);
int hulc2_nchw_to_nhwc(const float* src, float* dst, int F, int HW, int C, const float* mask, hulc2_stream_t stream);

/* ------------------------------------------------------------------ SpatialSoftmax  (vision_network.py:100-108)
 * x NHWC [F,HW,C]; out[f, 2c] = sum_i x_map[i] p_i, out[f, 2c+1] = sum_i y_map[i] p_i, p = softmax(x/temperature)
 * bwd: dx[f,i,c] = p_i/T * (gx (x_map_i - Ex) + gy (y_map_i - Ey)), zeroed where x <= 0 when relu_mask != 0. */
int hulc2_spatial_softmax_fwd(const float* x, const float* x_map, const float* y_map, const float* temperature,
                              float* out, int F, int HW, int C, hulc2_stream_t stream);
int hulc2_spatial_softmax_bwd(const float* x, const float* x_map, const float* y_map, const float* temperature,
                              const float* out, const float* dout, float* dx, float* dtemperature, int F, int HW, int C,
                              int relu_mask, hulc2_stream_t stream);

/* ------------------------------------------------------------------ LayerNorm (+ residual + dropout)
 * t = x + keep*scale*res (res/keep optional); y = (t-mean)/sqrt(var+eps)*gamma + beta   (eps 1e-5 everywhere)
 * bwd: dx (= grad of t), dres = dx*keep*scale (optional), dgamma/dbeta accumulated atomically. */
int hulc2_layernorm_fwd(const float* x, long long ldx, const float* res, long long ldr, const unsigned char* keep,
                        float keep_scale, const float* gamma, const float* beta, float* y, long long ldy, float* tsum,
                        float* mean, float* rstd, long long rows, int D, float eps, hulc2_stream_t stream);
int hulc2_layernorm_bwd(const float* dy, long long ldy, const float* t, long long ldt, const float* gamma,
                        const float* mean, const float* rstd, float* dx, long long lddx, float* dres,
                        const unsigned char* keep, float keep_scale, float* dgamma, float* dbeta, long long rows, int D,
                        hulc2_stream_t stream);
/* "_m" variants (here and below): the same operation that ALSO writes the bf16 operand mirror of its result (row stride ld16
 * elements, null = none) for the tcgen05 contraction that consumes it -- the mirror comes out of the producer's registers
 * instead of a separate fp32 -> bf16 pass over HBM (round-to-nearest-even, bit-identical to hulc2_f32_to_bf16). */
int hulc2_layernorm_fwd_m(const float* x, long long ldx, const float* res, long long ldr, const unsigned char* keep,
                          float keep_scale, const float* gamma, const float* beta, float* y, long long ldy, float* tsum,
                          float* mean, float* rstd, long long rows, int D, float eps, void* y16, long long ld16,
                          hulc2_stream_t stream);
int hulc2_layernorm_bwd_m(const float* dy, long long ldy, const float* t, long long ldt, const float* gamma,
                          const float* mean, const float* rstd, float* dx, long long lddx, float* dres,
                          const unsigned char* keep, float keep_scale, float* dgamma, float* dbeta, long long rows, int D,
                          void* dx16, void* dres16, long long ld16, hulc2_stream_t stream);

/* ------------------------------------------------------------------ plan-recognition transformer pieces
 * plan_recognition_net.py:125-148 + torch nn.TransformerEncoderLayer (post-LN, ReLU, 8 heads x 16, S <= 32).
 * add_pos: x[b,s,:] = (emb[b,s,:] + pos[s,:]) * keep*scale.   attention: qkv [B*S, 3E] packed [q|k|v],
 * probabilities p [B,H,S,S] are saved (pre-dropout) for the backward.   mean_seq: out[b,:] = mean_s x[b,s,:]. */
int hulc2_add_pos_fwd(const float* emb, const float* pos, const unsigned char* keep, float keep_scale, float* out, int B,
                      int S, int E, hulc2_stream_t stream);
int hulc2_add_pos_bwd(const float* dout, const unsigned char* keep, float keep_scale, float* demb, float* dpos, int B,
                      int S, int E, hulc2_stream_t stream);
int hulc2_attention_fwd(const float* qkv, const unsigned char* keep, float keep_scale, float* out, float* probs, int B,
                        int S, int H, int Dh, hulc2_stream_t stream);
int hulc2_attention_bwd(const float* qkv, const float* probs, const unsigned char* keep, float keep_scale,
                        const float* dout, float* dqkv, int B, int S, int H, int Dh, hulc2_stream_t stream);
int hulc2_add_pos_fwd_m(const float* emb, const float* pos, const unsigned char* keep, float keep_scale, float* out,
                        void* out16, int B, int S, int E, hulc2_stream_t stream);
int hulc2_attention_fwd_m(const float* qkv, const unsigned char* keep, float keep_scale, float* out, float* probs,
                          void* out16, long long ld16, int B, int S, int H, int Dh, hulc2_stream_t stream);
int hulc2_attention_bwd_m(const float* qkv, const float* probs, const unsigned char* keep, float keep_scale,
                          const float* dout, float* dqkv, void* dqkv16, long long ld16, int B, int S, int H, int Dh,
                          hulc2_stream_t stream);
int hulc2_mean_seq_fwd(const float* x, float* out, int B, int S, int E, hulc2_stream_t stream);
int hulc2_mean_seq_bwd(const float* dout, float* dx, int B, int S, int E, hulc2_stream_t stream);

/* ------------------------------------------------------------------ latent plan  (distributions.py:15-60, hulc2.py:444-466)
 * kl: loss = beta*(alpha*KL(sg(pr)||pp) + (1-alpha)*KL(pr||sg(pp))), categorical over `classes`, summed over
 * categories, mean over B.  Gradients are written scaled by *gout (device scalar, may be null = 1).
 * onehot: plan[b, cat*classes + idx[b,cat]] = 1.  st_bwd: straight-through gradient of rsample() w.r.t. logits.
 * sample: idx[b,cat] = inverse-CDF draw from softmax(logits) with supplied uniform u[b,cat]. */
int hulc2_kl_fwd(const float* pp_logits, const float* pr_logits, float* loss, int B, int cats, int classes, float alpha,
                 float beta, hulc2_stream_t stream);
int hulc2_kl_bwd(const float* pp_logits, const float* pr_logits, const float* gout, float* dpp, float* dpr, int B,
                 int cats, int classes, float alpha, float beta, hulc2_stream_t stream);
int hulc2_onehot_fwd(const long long* idx, float* plan, int B, int cats, int classes, hulc2_stream_t stream);
int hulc2_st_onehot_bwd(const float* logits, const float* dplan, float* dlogits, int B, int cats, int classes,
                        hulc2_stream_t stream);
int hulc2_categorical_sample(const float* logits, const float* u, long long* idx, int B, int cats, int classes,
                             hulc2_stream_t stream);

/* ------------------------------------------------------------------ logistic-mixture decoder head
 * heads [rows, ld] = [logit_probs(A*M) | means(A*M) | log_scales(A*M, unclamped) | gripper logits(2)] per row,
 * rows are TIME-MAJOR (row = s*B + b); actions [B,S,A+1] batch-major.
 * loss (logistic_decoder_rnn.py:133-152,181-228): out[0] = total, out[1] = logistic NLL, out[2] = gripper CE.
 * bwd writes d(heads) scaled by *gout/(B*S).   sample (:231-255): uniforms u1 [B,S,A,M], u2 [B,S,A] -> act [B,S,A+1]. */
int hulc2_logistic_loss_fwd(const float* heads, long long ld, const float* actions, const float* act_min,
                            const float* act_max, float* out, int B, int S, int A, int M, int num_classes,
                            float log_scale_min, float gripper_alpha, int time_major, void* workspace,
                            long long workspace_bytes, hulc2_stream_t stream);
int hulc2_logistic_loss_bwd(const float* heads, long long ld, const float* actions, const float* act_min,
                            const float* act_max, const float* gout, float* dheads, int B, int S, int A, int M,
                            int num_classes, float log_scale_min, float gripper_alpha, int time_major,
                            hulc2_stream_t stream);
/* Segment variants: the B windows are columns [b0, b0+B) of a wider buffer holding B_total windows (several modalities
 * decoded in one recurrence call); `heads`/`dheads` point at the segment's first row (time-major: + b0*ld, batch-major:
 * + b0*S*ld), `actions` is the segment's own [B,S,A+1]; the mean is over the segment's B*S rows (hulc2.py:386-400 keeps
 * one action loss per modality). */
int hulc2_logistic_loss_seg_fwd(const float* heads, long long ld, const float* actions, const float* act_min,
                                const float* act_max, float* out, int B, int S, int A, int M, int num_classes,
                                float log_scale_min, float gripper_alpha, int time_major, int B_total, void* workspace,
                                long long workspace_bytes, hulc2_stream_t stream);
int hulc2_logistic_loss_seg_bwd(const float* heads, long long ld, const float* actions, const float* act_min,
                                const float* act_max, const float* gout, float* dheads, int B, int S, int A, int M,
                                int num_classes, float log_scale_min, float gripper_alpha, int time_major, int B_total,
                                hulc2_stream_t stream);
int hulc2_logistic_sample(const float* heads, long long ld, const float* u1, const float* u2,
                          const float* gripper_bounds, float* act, int B, int S, int A, int M, float log_scale_min,
                          int time_major, hulc2_stream_t stream);
/* splits the fused heads buffer into the reference's forward() outputs (logistic_decoder_rnn.py:275-284) */
/* Validation metrics of lmp_val / validation_step (hulc2.py:292-302, 559-575) for one modality's windows: pred = sampled
 * actions [B,S,A+1], actions = ground truth; out[4] = {mean |err| over the A continuous dims, over dims 0-2 (position), over
 * dims 3-5 (orientation), success rate of the discretised gripper action (pred > 0 ? 1 : -1) == gt}. */
int hulc2_val_metrics(const float* pred, const float* actions, float* out, int B, int S, int A, hulc2_stream_t stream);
int hulc2_heads_unpack(const float* heads, long long ld, float* logit_probs, float* log_scales, float* means,
                       float* gripper, int B, int S, int A, int M, float log_scale_min, int time_major,
                       hulc2_stream_t stream);

/* ------------------------------------------------------------------ tcp <-> world frames (gripper_control.py:16-63) */
int hulc2_world_to_tcp(const float* action, const float* robot_obs, int robot_dim, float* out, long long rows,
                       hulc2_stream_t stream);
int hulc2_tcp_to_world(const float* action, const float* robot_obs, int robot_dim, float* out, long long rows,
                       hulc2_stream_t stream);

/* ------------------------------------------------------------------ InfoNCE (hulc2.py:472-508)
 * img,txt [B,D] projected features (un-normalised); use [B] u8 row mask (null = all); logit_scale device scalar.
 * fwd: loss[0]; saves nothing (bwd recomputes).  workspace >= (2*B*(D+1) + 2*B*B + 4*B) floats (only
 * used when that exceeds 200 KB of shared memory). */
int hulc2_infonce_fwd(const float* img, const float* txt, const unsigned char* use, const float* logit_scale,
                      float* loss, int B, int D, void* workspace, long long workspace_bytes, hulc2_stream_t stream);
int hulc2_infonce_bwd(const float* img, const float* txt, const unsigned char* use, const float* logit_scale,
                      const float* gout, float* dimg, float* dtxt, float* dlogit_scale, int B, int D, void* workspace,
                      long long workspace_bytes, hulc2_stream_t stream);

/* ------------------------------------------------------------------ decoder recurrence (decoders/utils/rnn.py:5-14)
 * Elman ReLU RNN layer over time-major buffers: h[t] = relu(pre[t] + h[t-1] W_hh^T), pre [S,B,H] already holds
 * W_ih x_t + b_ih + b_hh.  bwd: dz[t] = (dh_out[t] + dz[t+1] W_hh) * (h[t] > 0), in place over dh (becomes dz). */
int hulc2_rnn_relu_fwd(const float* pre, const float* w_hh, const float* h0, float* h, int S, int B, int H,
                       int precision, void* workspace, long long workspace_bytes, hulc2_stream_t stream);
int hulc2_rnn_relu_bwd(float* dh_inout, const float* w_hh, const float* h, float* dh0, int S, int B, int H,
                       int precision, void* workspace, long long workspace_bytes, hulc2_stream_t stream);
/* + bf16 mirror of the states: h16 / dz16 = bf16 [S+1, B, H], 16-byte aligned; on return slot t + 1 holds bf16(h[t]) /
 * bf16(dz[t]) (slot 0 is scratch: the initial state).  Kernel (a') uses the buffer as its own step-to-step operand store (its
 * workspace need drops to 1024 bytes), so the mirror the caller's input / weight-gradient contractions read costs nothing. */
int hulc2_rnn_relu_fwd_m(const float* pre, const float* w_hh, const float* h0, float* h, void* h16, int S, int B, int H,
                         int precision, void* workspace, long long workspace_bytes, hulc2_stream_t stream);
int hulc2_rnn_relu_bwd_m(float* dh_inout, const float* w_hh, const float* h, float* dh0, void* dz16, int S, int B, int H,
                         int precision, void* workspace, long long workspace_bytes, hulc2_stream_t stream);
/* precision 1 runs all S steps in ONE persistent tcgen05 kernel when the shape fits, tried in this order:
 *  (a') TMA-fed cluster split-K kernel (rnn_cluster2_sm100.cu): as (a) with workspace >= 2*(S+1)*B*H + 4096 bytes; the state
 *       slice of a step arrives as TMA boxes, CTAs publish per warp (HULC2_RNN_V1=1 in the environment skips it);
 *  (a) cluster split-K kernel (rnn_cluster_sm100.cu): B <= 128, H % 512 == 0, H <= 2048, workspace >= 2*S*B*H + 1024
 *      bytes, H/16 CTAs in clusters of 8 or 4 all co-resident -- W_hh block resident in shared memory, partial sums reduced
 *      through distributed shared memory, per-K-slice release/acquire flags between steps;
 *  (b) 1-D persistent kernel (rnn_persistent_sm100.cu): B <= 128, H % 64 == 0, H/16 <= #SMs, workspace >= 2*S*B*H + 256;
 *  otherwise one GEMM per step. */
/* test / benchmark hook: 0 = automatic (default), -1 = skip (a'), 1 = skip (a') and (a), 2 = skip (a'), (a) and (b).
 * Returns the previous value. */
int hulc2_rnn_select_kernel(int which);
/* number of cluster_size-CTA (8 or 4) clusters of kernel (a) that can be co-resident on the current device; (a) runs with
 * clusters of 8 when H/128 of them fit, else with clusters of 4 when H/64 fit (a 148-SM B200 reports 15 clusters of 8) */
int hulc2_rnn_cluster_capacity(int cluster_size);
/* The persistent kernels (a)/(b) are launched cooperatively (every CTA / cluster co-resident or the launch is refused and the
 * next kernel in the list runs).  Should a step-flag wait still give up, the kernel flags the error on the device and finishes
 * (no trap, the context survives).  This reads the flags -- SYNCHRONISES the device: bit 0 = kernel (a), bit 1 = kernel (b),
 * bit 2 = kernel (a');
 * clear != 0 resets them; -1 on a CUDA error. */
int hulc2_rnn_device_error(int clear);
/* Which kernel served the last hulc2_rnn_relu_{fwd,bwd} call with precision 1 (bits 0-7: 1 = (a'), 2 = (a), 3 = (b), 4 = one GEMM per
 * step) and, in bits 8-15, why (a') last declined (0 = it ran; 1 shape, 2 workspace, 3 alignment, 4 no driver entry point,
 * 5 clusters not co-resident, 6 tensor map, 7 launch refused).  Tests use it to make sure no silent fallback is being measured. */
int hulc2_rnn_last_path(void);
/* Diagnostic: device buffer of (S+1)*16 u64 that receives clock64 stamps of CTA 0 of kernel (a') at its phase boundaries
 * (slot = 16*step + phase: 0 step start, 1 flag seen, 2 last box issued, 3 / 4 first / last box landed, 5 last MMA committed,
 * 6 accumulator complete, 7 TMEM read, 8 strips exchanged, 9 state stored, 10 flag released); null switches it off. */
int hulc2_rnn_set_trace(void* device_buffer);

/* ------------------------------------------------------------------ gated recurrence cells (decoders/utils/rnn.py:17-36)
 * One step of nn.GRU (gate order r,z,n) / nn.LSTM (i,f,g,o); the contractions are hulc2_gemm calls, these are the
 * fused elementwise cells.  gi [B,G*H] (row stride ldgi) = W_ih x_t + b_ih, gh [B,G*H] dense = W_hh h_{t-1} + b_hh.
 * GRU fwd: h = (1-z) n + z h_prev, save [B,4H] = r|z|n|gh_n.  GRU bwd: dh = dh_a + dh_b (either may be null) ->
 *   dgi [B,3H] (row stride lddgi), dgh [B,3H], dh_prev [B,H] = dh*z (the recurrent GEMM then accumulates dgh W_hh).
 * LSTM fwd: c = f c_prev + i g, h = o tanh(c), save [B,4H] = activated i|f|g|o.  LSTM bwd: dc (null = 0) is the
 *   gradient into c_t from step t+1; writes dgates [B,4H] (row stride lddg; same for gi and gh) and dc_prev (may alias dc).
 * h_prev / c_prev null = zeros (h_0 is None).  H and the row strides must be multiples of 4. */
int hulc2_gru_cell_fwd(const float* gi, long long ldgi, const float* gh, const float* h_prev, float* h, float* save,
                       int B, int H, hulc2_stream_t stream);
int hulc2_gru_cell_bwd(const float* dh_a, const float* dh_b, const float* save, const float* h_prev, float* dgi,
                       long long lddgi, float* dgh, float* dh_prev, int B, int H, hulc2_stream_t stream);
int hulc2_lstm_cell_fwd(const float* gi, long long ldgi, const float* gh, const float* c_prev, float* h, float* c,
                        float* save, int B, int H, hulc2_stream_t stream);
int hulc2_lstm_cell_bwd(const float* dh_a, const float* dh_b, const float* dc, const float* save, const float* c,
                        const float* c_prev, float* dgates, long long lddg, float* dc_prev, int B, int H,
                        hulc2_stream_t stream);

/* ------------------------------------------------------------------ continuous latent plan (distributions.py:28-29,55-59)
 * state: x [B,2P] -> mean = x[:, :P], std = softplus(x[:, P:]) + 1e-4 (and its gradient);  rsample: mean + std*eps with
 * caller-supplied standard-normal eps;  kl: beta*(alpha*KL(sg(pr)||pp) + (1-alpha)*KL(pr||sg(pp))) for diagonal normals,
 * summed over P, mean over B (hulc2.py:444-466; rsample's mean may be null = 0); gradients scaled by *gout (device scalar, null = 1), null outputs skipped. */
int hulc2_gauss_state_fwd(const float* x, float* mean, float* std, int B, int P, hulc2_stream_t stream);
int hulc2_gauss_state_bwd(const float* x, const float* dmean, const float* dstd, float* dx, int B, int P,
                          hulc2_stream_t stream);
int hulc2_gauss_rsample(const float* mean, const float* std, const float* eps, float* plan, long long n,
                        hulc2_stream_t stream);
/* eps = sqrt(-2 ln u1) cos(2 pi u2): standard normals from two Philox uniform streams (the library's default draw) */
int hulc2_box_muller(const float* u1, const float* u2, float* eps, long long n, hulc2_stream_t stream);
int hulc2_gauss_kl_fwd(const float* pp_mean, const float* pp_std, const float* pr_mean, const float* pr_std, float* loss,
                       int B, int P, float alpha, float beta, hulc2_stream_t stream);
int hulc2_gauss_kl_bwd(const float* pp_mean, const float* pp_std, const float* pr_mean, const float* pr_std,
                       const float* gout, float* dpp_mean, float* dpp_std, float* dpr_mean, float* dpr_std, int B, int P,
                       float alpha, float beta, hulc2_stream_t stream);

/* ------------------------------------------------------------------ optimizer + noise
 * Adam (torch.optim.Adam semantics, conf/model/optimizer/adam.yaml): one launch over a flat arena. */
int hulc2_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, hulc2_stream_t stream);
/* Philox4x32-10 counter-based noise: uniforms in [0,1) / dropout keep masks with P(keep) = 1-p. */
int hulc2_philox_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset,
                         hulc2_stream_t stream);
int hulc2_dropout_mask(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                       hulc2_stream_t stream);

/* CUDA-graph friendly variants: a whole train step (forward, backward, all-reduce, Adam) is captured once and replayed,
 * so per-step scalars live in device counters.  epoch: noise counter offset += *epoch << 40 (null = 0);
 * step_counter: Adam bias corrections use step = *step_counter + step_bias.  counter_add: *counter += inc. */
int hulc2_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, const unsigned long long* step_counter, int step_bias,
                        float grad_scale, hulc2_stream_t stream);
/* As adam_step_dev with the learning rate read from device memory (*lr_dev) at run time, so that an LR scheduler
 * (hulc2.py:185-198, conf/model/lr_scheduler) keeps steering a captured train step. */
int hulc2_adam_step_graph(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float beta1, float beta2,
                          float eps, float weight_decay, const unsigned long long* step_counter, int step_bias,
                          float grad_scale, hulc2_stream_t stream);
int hulc2_philox_uniform_ep(float* out, long long n, unsigned long long seed, unsigned long long offset,
                            const unsigned long long* epoch, hulc2_stream_t stream);
int hulc2_dropout_mask_ep(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                          const unsigned long long* epoch, hulc2_stream_t stream);
int hulc2_counter_add(unsigned long long* counter, unsigned long long inc, hulc2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HULC2_B200_H */
