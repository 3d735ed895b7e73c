// Shared operand addressing / epilogue definitions of the dense-contraction kernels (fp32 CUDA-core and
// bf16 tcgen05 backends).  Operands are addressed as off(r,k) = R(r) + Kf(k) (separable): row-major,
// transposed, two-level row strides, both im2col layouts; the dgrad gather adds a validity predicate.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace hulc2 {

enum { OP_DENSE = 0, OP_IM2COL = 1, OP_IM2COL_T = 2, OP_DGRAD = 3, OP_DGRAD_W = 4 };

struct ConvGeom {
  int C, H, W, KH, KW, OH, OW, stride, nhwc, Cout;
  // input-gradient parity class (ih % stride == ph, iw % stride == pw): class extent CH x CW pixels, KA x KB taps
  int ph, pw, CH, CW, KA, KB;
};

struct Operand {
  const float* p;
  long long rs, ks;            // dense: off = r*rs + k*ks
  int r_inner;                 // dense: optional 2-level rows: R(r) = (r / r_inner)*rs_outer + (r % r_inner)*rs_inner
  long long rs_outer, rs_inner;
  ConvGeom g;
};

struct Epilogue {
  float* C;
  long long ldc;
  int c_inner;
  long long cs_outer, cs_inner;
  const float* bias;           // [N] or null
  const float* add;            // [M, ld_add] or null
  long long ld_add;
  const float* mask;           // keep where mask[m,n] > 0 (ReLU backward), or null
  long long ld_mask;
  const unsigned char* keep;   // dropout keep mask (u8) or null; value *= keep ? keep_scale : 0
  long long ld_keep;
  float keep_scale;
  int relu, accumulate;
  float alpha;
  // optional output-row remap for the strided input-gradient classes: m = (f, i, j) -> (f*H + i*s+ph)*W + j*s+pw
  int rm_on, rm_chw, rm_cw, rm_s, rm_ph, rm_pw, rm_H, rm_W;
};

struct GemmParams {
  int M, N, K;
  Operand A, B;
  Epilogue E;
  int splits, kchunk;          // split-K: blockIdx.z handles [z*kchunk, (z+1)*kchunk)
  float* partial;              // [splits, M, N] when splits > 1
};

__device__ __forceinline__ long long im2col_pixel_off(const ConvGeom& g, int r) {
  int ohw = g.OH * g.OW;
  int f = r / ohw, rem = r - f * ohw;
  int oh = rem / g.OW, ow = rem - oh * g.OW;
  if (g.nhwc) return (((long long)f * g.H + oh * g.stride) * g.W + ow * g.stride) * g.C;
  return (long long)f * g.C * g.H * g.W + (long long)(oh * g.stride) * g.W + ow * g.stride;
}
__device__ __forceinline__ long long im2col_k_off(const ConvGeom& g, int k) {
  if (g.nhwc) {  // k = (kh, kw, ci)
    int kwc = g.KW * g.C;
    int kh = k / kwc, rem = k - kh * kwc;
    int kw = rem / g.C, ci = rem - kw * g.C;
    return ((long long)kh * g.W + kw) * g.C + ci;
  }
  int khw = g.KH * g.KW;  // k = (ci, kh, kw)
  int ci = k / khw, rem = k - ci * khw;
  int kh = rem / g.KW, kw = rem - kh * g.KW;
  return (long long)ci * g.H * g.W + (long long)kh * g.W + kw;
}

template <int MODE>
__device__ __forceinline__ long long row_off(const Operand& o, int r) {
  if (MODE == OP_DENSE) {
    if (o.r_inner > 0) return (long long)(r / o.r_inner) * o.rs_outer + (long long)(r % o.r_inner) * o.rs_inner;
    return (long long)r * o.rs;
  } else if (MODE == OP_IM2COL) {
    return im2col_pixel_off(o.g, r);
  } else if (MODE == OP_IM2COL_T) {
    return im2col_k_off(o.g, r);
  } else if (MODE == OP_DGRAD_W) {
    return r;
  }
  return 0;
}
template <int MODE>
__device__ __forceinline__ long long col_off(const Operand& o, int k) {
  if (MODE == OP_DENSE) return (long long)k * o.ks;
  if (MODE == OP_IM2COL) return im2col_k_off(o.g, k);
  if (MODE == OP_IM2COL_T) return im2col_pixel_off(o.g, k);
  if (MODE == OP_DGRAD_W) {  // k = (a, b, co) of the parity class -> w_hwoi[kh = ph + a s][kw = pw + b s][co][:]
    const ConvGeom& g = o.g;
    int bc = g.KB * g.Cout;
    int a = k / bc, rem = k - a * bc;
    int b = rem / g.Cout, co = rem - b * g.Cout;
    return ((long long)((g.ph + a * g.stride) * g.KW + (g.pw + b * g.stride)) * g.Cout + co) * g.C;
  }
  return 0;
}

// dgrad gather within one parity class: row = class pixel (f, i, j) [input pixel ih = i s + ph, iw = j s + pw],
// k = (a, b, co) [tap kh = ph + a s, kw = pw + b s]  ->  dZ[f, i - a, j - b, co] when in range, else 0.
// Only taps that can reach the class are enumerated, so a stride-2 4x4 conv needs K = 2*2*Cout instead of 4*4*Cout.
__device__ __forceinline__ bool dgrad_src(const ConvGeom& g, int f, int i, int j, int a, int b, long long& off) {
  int oh = i - a, ow = j - b;
  if (oh < 0 || ow < 0 || oh >= g.OH || ow >= g.OW) return false;
  off = (((long long)f * g.OH + oh) * g.OW + ow) * g.Cout;
  return true;
}
__device__ __forceinline__ void dgrad_row_decode(const ConvGeom& g, int r, int& f, int& i, int& j) {
  int chw = g.CH * g.CW;
  f = r / chw;
  int rem = r - f * chw;
  i = rem / g.CW;
  j = rem - i * g.CW;
}
__device__ __forceinline__ void dgrad_k_decode(const ConvGeom& g, int k, int& a, int& b, int& co) {
  int bc = g.KB * g.Cout;
  a = k / bc;
  int rem = k - a * bc;
  b = rem / g.Cout;
  co = rem - b * g.Cout;
}


// epilogue applied to one accumulator value
__device__ __forceinline__ float apply_epilogue(const Epilogue& E, float acc, int m, int n, long long crow) {
  float v = E.alpha * acc;
  if (E.bias) v += E.bias[n];
  if (E.add) v += E.add[(long long)m * E.ld_add + n];
  if (E.accumulate) v += E.C[crow + n];
  if (E.relu) v = fmaxf(v, 0.f);
  if (E.mask) v = (E.mask[(long long)m * E.ld_mask + n] > 0.f) ? v : 0.f;
  if (E.keep) v = E.keep[(long long)m * E.ld_keep + n] ? v * E.keep_scale : 0.f;
  return v;
}
// logical output row -> physical row (identity unless the strided input-gradient remap is on)
__device__ __forceinline__ int phys_row(const Epilogue& E, int m) {
  if (!E.rm_on) return m;
  int f = m / E.rm_chw, rem = m - f * E.rm_chw;
  int i = rem / E.rm_cw, j = rem - i * E.rm_cw;
  return (f * E.rm_H + i * E.rm_s + E.rm_ph) * E.rm_W + j * E.rm_s + E.rm_pw;
}
__device__ __forceinline__ long long c_row_off(const Epilogue& E, int m) {
  return E.c_inner > 0 ? (long long)(m / E.c_inner) * E.cs_outer + (long long)(m % E.c_inner) * E.cs_inner : (long long)m * E.ldc;
}

// split-K second stage: C = epilogue(sum_z partial[z]) -- the FULL epilogue (alpha, bias, add, accumulate, ReLU, mask, dropout keep) and
// the optional bf16 mirror of the result, so that skinny problems with a fused epilogue (the M = 64..128 layers of the plan
// networks and goal encoders: one M-tile x 16 N-tiles = 16 CTAs streaming 8 MB of weights) can be split along K as well.
static __global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, Epilogue E,
                                            __nv_bfloat16* __restrict__ C16, long long ld16,
                                            const float* __restrict__ rowsum_part = nullptr, float* __restrict__ rowsum = nullptr) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)M * N;
  if (idx >= total) return;
  if (rowsum && idx < M) {               // row sums of A (gemm_tma_sm100.cu), one partial per split
    float rs = 0.f;
    for (int z = 0; z < splits; ++z) rs += rowsum_part[(long long)z * M + idx];
    rowsum[idx] = rs;
  }
  int m = (int)(idx / N), n = (int)(idx - (long long)m * N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long long)z * total + idx];
  long long crow = c_row_off(E, m);
  const float v = apply_epilogue(E, s, m, n, crow);
  E.C[crow + n] = v;
  if (C16) C16[(long long)m * ld16 + n] = __float2bfloat16_rn(v);
}

// 16-byte variant: one thread = 4 adjacent columns of one row (N % 4 == 0, every pointer involved 16-byte aligned, dense C rows).
// All partial loads of a thread are issued before the first add (49 of these run per train step: at one element per thread they
// were the largest item of the launch tail, 6.2 us each).
static __global__ void __launch_bounds__(256) splitk_reduce4_kernel(const float* __restrict__ part, int splits, int M, int N, Epilogue E,
                                                                   __nv_bfloat16* __restrict__ C16, long long ld16,
                                                                   const float* __restrict__ rowsum_part, float* __restrict__ rowsum) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)M * N;
  if (rowsum && t < M) {
    float rs = 0.f;
    for (int z = 0; z < splits; ++z) rs += rowsum_part[(long long)z * M + t];
    rowsum[t] = rs;
  }
  const long long idx = t * 4;
  if (idx >= total) return;
  const int m = (int)(idx / N), n = (int)(idx - (long long)m * N);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int z = 0;
  for (; z + 4 <= splits; z += 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(part + (long long)z * total + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(part + (long long)(z + 1) * total + idx));
    const float4 c = __ldg(reinterpret_cast<const float4*>(part + (long long)(z + 2) * total + idx));
    const float4 d = __ldg(reinterpret_cast<const float4*>(part + (long long)(z + 3) * total + idx));
    // same left-to-right order as the scalar kernel: ((s + a) + b) + c) + d
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
    s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
    s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
  }
  for (; z < splits; ++z) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(part + (long long)z * total + idx));
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  const long long crow = (long long)m * E.ldc;
  float v[4] = {E.alpha * s.x, E.alpha * s.y, E.alpha * s.z, E.alpha * s.w};
  if (E.bias) { const float4 q = __ldg(reinterpret_cast<const float4*>(E.bias + n)); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
  if (E.add) { const float4 q = *reinterpret_cast<const float4*>(E.add + (long long)m * E.ld_add + n); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
  if (E.accumulate) { const float4 q = *reinterpret_cast<const float4*>(E.C + crow + n); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
  if (E.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
  if (E.mask) {
    const float4 q = *reinterpret_cast<const float4*>(E.mask + (long long)m * E.ld_mask + n);
    v[0] = q.x > 0.f ? v[0] : 0.f; v[1] = q.y > 0.f ? v[1] : 0.f; v[2] = q.z > 0.f ? v[2] : 0.f; v[3] = q.w > 0.f ? v[3] : 0.f;
  }
  if (E.keep) {
    const uchar4 q = *reinterpret_cast<const uchar4*>(E.keep + (long long)m * E.ld_keep + n);
    v[0] = q.x ? v[0] * E.keep_scale : 0.f; v[1] = q.y ? v[1] * E.keep_scale : 0.f;
    v[2] = q.z ? v[2] * E.keep_scale : 0.f; v[3] = q.w ? v[3] * E.keep_scale : 0.f;
  }
  *reinterpret_cast<float4*>(E.C + crow + n) = make_float4(v[0], v[1], v[2], v[3]);
  if (C16) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(C16 + (long long)m * ld16 + n) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}

// picks the 16-byte kernel when the layout allows it
static inline void launch_splitk_reduce(const float* part, int splits, int M, int N, const Epilogue& E, __nv_bfloat16* C16, long long ld16,
                                        const float* rowsum_part, float* rowsum, cudaStream_t st) {
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  bool vec = N % 4 == 0 && E.c_inner == 0 && !E.rm_on && E.ldc % 4 == 0 && al16(part) && al16(E.C);
  if (E.bias) vec = vec && al16(E.bias);
  if (E.add) vec = vec && al16(E.add) && E.ld_add % 4 == 0;
  if (E.mask) vec = vec && al16(E.mask) && E.ld_mask % 4 == 0;
  if (E.keep) vec = vec && (((uintptr_t)E.keep & 3) == 0) && E.ld_keep % 4 == 0;
  if (C16) vec = vec && (((uintptr_t)C16 & 7) == 0) && ld16 % 4 == 0;
  const long long total = (long long)M * N;
  if (vec) {
    long long threads = total / 4 > M ? total / 4 : M;
    splitk_reduce4_kernel<<<(unsigned)hulc2_cdiv(threads, 256), 256, 0, st>>>(part, splits, M, N, E, C16, ld16, rowsum_part, rowsum);
  } else {
    splitk_reduce_kernel<<<(unsigned)hulc2_cdiv(total, 256), 256, 0, st>>>(part, splits, M, N, E, C16, ld16, rowsum_part, rowsum);
  }
}

// ------------------------------------------------------------------ host-side parameter builders
inline void fill_epilogue(Epilogue& E, float* C, long long ldc) {
  E = Epilogue{};
  E.C = C; E.ldc = ldc; E.alpha = 1.f; E.keep_scale = 1.f;
}
inline ConvGeom geom_of(const hulc2_conv_args* a) {
  ConvGeom g;
  g.C = a->C; g.H = a->H; g.W = a->W; g.KH = a->KH; g.KW = a->KW; g.stride = a->stride;
  g.OH = (a->H - a->KH) / a->stride + 1; g.OW = (a->W - a->KW) / a->stride + 1;
  g.nhwc = a->in_nhwc; g.Cout = a->Cout;
  return g;
}
// split the contraction over `want_ctas` CTAs when the output tile grid is small; kgran = k-tile size
inline void plan_splitk(GemmParams& p, long long out_ctas, int want_ctas, int min_k_per_split, int kgran, void* ws, long long ws_bytes) {
  p.splits = 1; p.kchunk = ((p.K + kgran - 1) / kgran) * kgran;
  if (p.kchunk == 0) p.kchunk = kgran;
  p.partial = nullptr;
  if (!ws || p.K < 2 * min_k_per_split) return;
  int want = (int)((want_ctas + out_ctas - 1) / out_ctas);
  int maxs = p.K / min_k_per_split; if (maxs < 1) maxs = 1;
  int s = want < maxs ? want : maxs;
  long long per = (long long)p.M * p.N * sizeof(float);
  while (s > 1 && (long long)s * per > ws_bytes) s /= 2;
  if (s <= 1) return;
  p.kchunk = ((hulc2_cdiv(p.K, s) + kgran - 1) / kgran) * kgran;
  p.splits = hulc2_cdiv(p.K, p.kchunk);
  p.partial = (float*)ws;
}
inline bool dense_params(const hulc2_gemm_args* a, GemmParams& p) {
  if (!a || a->M < 0 || a->N < 0 || a->K < 0 || !a->A || !a->B || !a->C) { hulc2_set_error("gemm: bad args"); return false; }
  p = GemmParams{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.A.p = a->A; p.A.rs = a->a_rs; p.A.ks = a->a_ks; p.A.r_inner = a->a_inner; p.A.rs_outer = a->a_rs_outer; p.A.rs_inner = a->a_rs_inner;
  p.B.p = a->B; p.B.rs = a->b_rs; p.B.ks = a->b_ks;
  fill_epilogue(p.E, a->C, a->ldc);
  p.E.c_inner = a->c_inner; p.E.cs_outer = a->c_rs_outer; p.E.cs_inner = a->c_rs_inner;
  p.E.bias = a->bias; p.E.add = a->add; p.E.ld_add = a->ld_add; p.E.mask = a->mask; p.E.ld_mask = a->ld_mask;
  p.E.keep = a->keep; p.E.ld_keep = a->ld_keep; p.E.keep_scale = a->keep_scale;
  p.E.relu = a->relu; p.E.accumulate = a->accumulate; p.E.alpha = a->alpha;
  return true;
}
inline bool simple_epilogue(const hulc2_gemm_args* a) { return !a->add && !a->mask && !a->keep && !a->relu; }
inline void conv_fwd_params(const hulc2_conv_args* a, GemmParams& p) {
  ConvGeom g = geom_of(a);
  p = GemmParams{};
  p.M = a->F * g.OH * g.OW; p.N = a->Cout; p.K = a->C * a->KH * a->KW;
  p.A.p = a->x; p.A.g = g;
  p.B.p = a->w; p.B.rs = p.K; p.B.ks = 1;
  fill_epilogue(p.E, a->y, a->Cout);
  p.E.bias = a->bias; p.E.relu = a->relu;
}
inline void conv_wgrad_params(const hulc2_conv_args* a, GemmParams& p) {
  ConvGeom g = geom_of(a);
  p = GemmParams{};
  p.M = a->Cout; p.N = a->C * a->KH * a->KW; p.K = a->F * g.OH * g.OW;
  p.A.p = a->dy; p.A.rs = 1; p.A.ks = a->Cout;       // A(m=co, k=pixel) = dZ[pixel*Cout + co]
  p.B.p = a->x; p.B.g = g;                            // B(n=kidx, k=pixel) = x[pix_off(pixel) + k_off(kidx)]
  fill_epilogue(p.E, a->dw, p.N);
  p.E.accumulate = a->accumulate;
}
// one parity class (ph, pw) of the input gradient; returns false when the class is empty
inline bool conv_dgrad_params(const hulc2_conv_args* a, int ph, int pw, GemmParams& p) {
  ConvGeom g = geom_of(a);
  const int s = a->stride;
  g.ph = ph; g.pw = pw;
  g.CH = (a->H - ph + s - 1) / s; g.CW = (a->W - pw + s - 1) / s;
  g.KA = (a->KH - ph + s - 1) / s; g.KB = (a->KW - pw + s - 1) / s;
  p = GemmParams{};
  if (ph >= a->H || pw >= a->W || g.CH <= 0 || g.CW <= 0) return false;
  p.M = a->F * g.CH * g.CW; p.N = a->C; p.K = (g.KA > 0 && g.KB > 0) ? g.KA * g.KB * a->Cout : 0;
  p.A.p = a->dy; p.A.g = g;
  p.B.p = a->w; p.B.rs = 1; p.B.ks = a->C; p.B.g = g; // B(n=ci, k=(a,b,co)) = w_hwoi[kh][kw][co][ci]
  fill_epilogue(p.E, a->dx, a->C);
  p.E.mask = a->xmask; p.E.ld_mask = a->C;
  p.E.rm_on = 1; p.E.rm_chw = g.CH * g.CW; p.E.rm_cw = g.CW; p.E.rm_s = s; p.E.rm_ph = ph; p.E.rm_pw = pw;
  p.E.rm_H = a->H; p.E.rm_W = a->W;
  return p.M > 0;
}

}  // namespace hulc2
