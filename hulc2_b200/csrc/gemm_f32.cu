// fp32 CUDA-core GEMM with implicit-convolution operand loaders (the exact-parity path).
//
// One register-tiled kernel serves every dense contraction of the HULC++ policy step in fp32:
//   Linear fwd / dgrad / wgrad          (goal encoders, plan proposal, transformer, RNN, heads)
//   conv fwd   = implicit GEMM, A = im2col(x)            (vision_network.py:38-48, vision_network_gripper.py:11-26)
//   conv wgrad = implicit GEMM, B = im2col(x)^T, split-K over output pixels
//   conv dgrad = implicit GEMM, A = strided gather of dZ
// Operands are addressed as  off(r,k) = R(r) + Kf(k)  (separable), which covers row-major, transposed,
// two-level row strides and both im2col layouts; the dgrad gather adds a validity predicate.
// The bf16 tensor-core path (gemm_bf16_sm100.cu) shares the same argument structs.
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

enum { OP_DENSE = 0, OP_IM2COL = 1, OP_IM2COL_T = 2, OP_DGRAD = 3 };

struct ConvGeom {
  int C, H, W, KH, KW, OH, OW, stride, nhwc, Cout;
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
  }
  return 0;
}
template <int MODE>
__device__ __forceinline__ long long col_off(const Operand& o, int k) {
  if (MODE == OP_DENSE) return (long long)k * o.ks;
  if (MODE == OP_IM2COL) return im2col_k_off(o.g, k);
  if (MODE == OP_IM2COL_T) return im2col_pixel_off(o.g, k);
  return 0;
}

// dgrad gather: row r = input pixel (f, ih, iw) of the NHWC input-gradient, k = (kh, kw, co).
// value = dZ[f, (ih-kh)/s, (iw-kw)/s, co] when the division is exact and in range, else 0.
__device__ __forceinline__ float dgrad_load(const Operand& o, int f, int ih, int iw, int kh, int kw, int co) {
  const ConvGeom& g = o.g;
  int th = ih - kh, tw = iw - kw;
  if (th < 0 || tw < 0) return 0.f;
  int oh = th / g.stride, ow = tw / g.stride;
  if (oh * g.stride != th || ow * g.stride != tw || oh >= g.OH || ow >= g.OW) return 0.f;
  return o.p[(((long long)f * g.OH + oh) * g.OW + ow) * g.Cout + co];
}

// Loads a [ROWS x BK] operand tile into registers. KFAST: thread owns one k-lane and ROWS/16 rows;
// else thread owns one row (or row group) and several k-lanes.
template <int ROWS, int MODE, bool KFAST>
struct TileLoader {
  static constexpr int PER = ROWS * BK / NTHREADS;   // elements per thread
  static constexpr int RSTEP = KFAST ? 16 : 0;
  static constexpr int KGROUPS = KFAST ? 1 : (NTHREADS / ROWS > 0 ? NTHREADS / ROWS : 1);
  float v[PER];
  long long roff[KFAST ? PER : 1];
  int rdec[KFAST ? PER : 1][3];
  bool rvalid[KFAST ? PER : 1];
  int r0, klane;

  __device__ __forceinline__ void init(const Operand& o, int row_base, int nrows_total) {
    int tid = threadIdx.x;
    if (KFAST) {
      klane = tid & 15;
      r0 = tid >> 4;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        int r = row_base + r0 + 16 * i;
        rvalid[i] = r < nrows_total;
        int rr = rvalid[i] ? r : 0;
        if (MODE == OP_DGRAD) {
          int hw = o.g.H * o.g.W;
          int f = rr / hw, rem = rr - f * hw;
          rdec[i][0] = f; rdec[i][1] = rem / o.g.W; rdec[i][2] = rem - (rem / o.g.W) * o.g.W;
          roff[i] = 0;
        } else {
          roff[i] = row_off<MODE>(o, rr);
        }
      }
    } else {
      r0 = tid % ROWS;
      klane = tid / ROWS;
      int r = row_base + r0;
      rvalid[0] = r < nrows_total;
      roff[0] = row_off<MODE>(o, rvalid[0] ? r : 0);
    }
  }

  __device__ __forceinline__ void load(const Operand& o, int k0, int kend) {
    if (KFAST) {
      int kg = k0 + klane;
      bool kv = kg < kend;
      if (MODE == OP_DGRAD) {
        int kwc = o.g.KW * o.g.Cout;
        int kk = kv ? kg : 0;
        int kh = kk / kwc, rem = kk - kh * kwc;
        int kw = rem / o.g.Cout, co = rem - kw * o.g.Cout;
#pragma unroll
        for (int i = 0; i < PER; ++i)
          v[i] = (kv && rvalid[i]) ? dgrad_load(o, rdec[i][0], rdec[i][1], rdec[i][2], kh, kw, co) : 0.f;
      } else {
        long long koff = col_off<MODE>(o, kv ? kg : 0);
#pragma unroll
        for (int i = 0; i < PER; ++i) v[i] = (kv && rvalid[i]) ? __ldg(o.p + roff[i] + koff) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        int kg = k0 + klane + KGROUPS * i;
        bool kv = kg < kend && rvalid[0];
        v[i] = kv ? __ldg(o.p + roff[0] + col_off<MODE>(o, kg)) : 0.f;
      }
    }
  }

  template <int LD>
  __device__ __forceinline__ void store(float (*s)[LD]) const {
    if (KFAST) {
#pragma unroll
      for (int i = 0; i < PER; ++i) s[klane][r0 + 16 * i] = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < PER; ++i) s[klane + KGROUPS * i][r0] = v[i];
    }
  }
};

// thread-tile index -> tile row/col. For 8-wide thread tiles the 8 values are split in two groups
// of 4 half a tile apart, so consecutive threads read consecutive float4 from shared memory.
template <int T, int BT>
__device__ __forceinline__ int tile_idx(int t, int i) {
  if (T == 8) return (i < 4) ? (t * 4 + i) : (BT / 2 + t * 4 + (i - 4));
  return t * T + i;
}

template <int BM, int BN, int TM, int TN, int AMODE, bool AKF, int BMODE, bool BKF>
__global__ void __launch_bounds__(NTHREADS) gemm_f32_kernel(const GemmParams p) {
  static_assert((BM / TM) * (BN / TN) == NTHREADS, "thread tiling must cover the block tile");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);

  TileLoader<BM, AMODE, AKF> la;
  TileLoader<BN, BMODE, BKF> lb;
  la.init(p.A, m0, p.M);
  lb.init(p.B, n0, p.N);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ntiles = (kend - kbeg + BK - 1) / BK;
  if (ntiles > 0) {
    la.load(p.A, kbeg, kend);
    lb.load(p.B, kbeg, kend);
    la.template store<BM + 4>(As[0]);
    lb.template store<BN + 4>(Bs[0]);
  }
  __syncthreads();

  for (int t = 0; t < ntiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < ntiles) {
      la.load(p.A, kbeg + (t + 1) * BK, kend);
      lb.load(p.B, kbeg + (t + 1) * BK, kend);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      if (TM % 4 == 0) {
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          float4 t4 = *reinterpret_cast<const float4*>(&As[cur][kk][tile_idx<TM, BM>(ty, i)]);
          a[i] = t4.x; a[i + 1] = t4.y; a[i + 2] = t4.z; a[i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[cur][kk][tile_idx<TM, BM>(ty, i)];
      }
      if (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          float4 t4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tile_idx<TN, BN>(tx, j)]);
          b[j] = t4.x; b[j + 1] = t4.y; b[j + 2] = t4.z; b[j + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[cur][kk][tile_idx<TN, BN>(tx, j)];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      la.template store<BM + 4>(As[cur ^ 1]);
      lb.template store<BN + 4>(Bs[cur ^ 1]);
    }
    __syncthreads();
  }

  const Epilogue& E = p.E;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + tile_idx<TM, BM>(ty, i);
    if (m >= p.M) continue;
    if (p.splits > 1) {
      float* prow = p.partial + ((long long)blockIdx.z * p.M + m) * p.N;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int n = n0 + tile_idx<TN, BN>(tx, j);
        if (n < p.N) prow[n] = acc[i][j];
      }
      continue;
    }
    long long crow = E.c_inner > 0 ? (long long)(m / E.c_inner) * E.cs_outer + (long long)(m % E.c_inner) * E.cs_inner
                                   : (long long)m * E.ldc;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tile_idx<TN, BN>(tx, j);
      if (n >= p.N) continue;
      float v = E.alpha * acc[i][j];
      if (E.bias) v += E.bias[n];
      if (E.add) v += E.add[(long long)m * E.ld_add + n];
      if (E.accumulate) v += E.C[crow + n];
      if (E.relu) v = fmaxf(v, 0.f);
      if (E.mask) v = (E.mask[(long long)m * E.ld_mask + n] > 0.f) ? v : 0.f;
      if (E.keep) v = E.keep[(long long)m * E.ld_keep + n] ? v * E.keep_scale : 0.f;
      E.C[crow + n] = v;
    }
  }
}

// split-K second stage: C = (accumulate ? C : 0) + alpha * sum_z partial[z] (+ bias)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, Epilogue E) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)M * N;
  if (idx >= total) return;
  int m = (int)(idx / N), n = (int)(idx - (long long)m * N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(long long)z * total + idx];
  long long crow = E.c_inner > 0 ? (long long)(m / E.c_inner) * E.cs_outer + (long long)(m % E.c_inner) * E.cs_inner
                                 : (long long)m * E.ldc;
  float v = E.alpha * s;
  if (E.bias) v += E.bias[n];
  if (E.accumulate) v += E.C[crow + n];
  E.C[crow + n] = v;
}

template <int BM, int BN, int TM, int TN, int AMODE, bool AKF, int BMODE, bool BKF>
int launch_cfg(const GemmParams& p, cudaStream_t st) {
  dim3 grid(hulc2_cdiv(p.M, BM), hulc2_cdiv(p.N, BN), p.splits);
  gemm_f32_kernel<BM, BN, TM, TN, AMODE, AKF, BMODE, BKF><<<grid, NTHREADS, 0, st>>>(p);
  HULC2_CHECK_LAUNCH();
  if (p.splits > 1) {
    long long total = (long long)p.M * p.N;
    splitk_reduce_kernel<<<hulc2_cdiv(total, 256), 256, 0, st>>>(p.partial, p.splits, p.M, p.N, p.E);
    HULC2_CHECK_LAUNCH();
  }
  return HULC2_OK;
}

// tile selection: the largest tile that still yields >= ~1 wave of CTAs, else the smallest.
template <int AMODE, bool AKF, int BMODE, bool BKF>
int launch_modes(const GemmParams& p, cudaStream_t st) {
  auto ctas = [&](int bm, int bn) { return (long long)hulc2_cdiv(p.M, bm) * hulc2_cdiv(p.N, bn) * p.splits; };
  if (p.N <= 32 && p.M >= 256) return launch_cfg<256, 32, 8, 4, AMODE, AKF, BMODE, BKF>(p, st);
  if (p.N > 64 && p.M > 64 && ctas(128, 128) >= 120) return launch_cfg<128, 128, 8, 8, AMODE, AKF, BMODE, BKF>(p, st);
  if (p.M > 64 && ctas(128, 64) >= 120) return launch_cfg<128, 64, 8, 4, AMODE, AKF, BMODE, BKF>(p, st);
  if (ctas(64, 64) >= 120) return launch_cfg<64, 64, 4, 4, AMODE, AKF, BMODE, BKF>(p, st);
  return launch_cfg<32, 32, 2, 2, AMODE, AKF, BMODE, BKF>(p, st);
}

int dispatch(const GemmParams& p, int amode, int bmode, cudaStream_t st) {
  bool akf = (amode == OP_DENSE) ? (p.A.ks == 1) : true;
  bool bkf = (bmode == OP_DENSE) ? (p.B.ks == 1) : false;
  if (amode == OP_DENSE && bmode == OP_DENSE) {
    if (akf && bkf) return launch_modes<OP_DENSE, true, OP_DENSE, true>(p, st);
    if (akf && !bkf) return launch_modes<OP_DENSE, true, OP_DENSE, false>(p, st);
    if (!akf && !bkf) return launch_modes<OP_DENSE, false, OP_DENSE, false>(p, st);
    return launch_modes<OP_DENSE, false, OP_DENSE, true>(p, st);
  }
  if (amode == OP_IM2COL && bmode == OP_DENSE && bkf) return launch_modes<OP_IM2COL, true, OP_DENSE, true>(p, st);
  if (amode == OP_DENSE && !akf && bmode == OP_IM2COL_T) return launch_modes<OP_DENSE, false, OP_IM2COL_T, false>(p, st);
  if (amode == OP_DGRAD && bmode == OP_DENSE && !bkf) return launch_modes<OP_DGRAD, true, OP_DENSE, false>(p, st);
  hulc2_set_error("gemm_f32: unsupported operand mode combination");
  return HULC2_EINVAL;
}

void fill_epilogue(Epilogue& E, float* C, long long ldc) {
  E = Epilogue{};
  E.C = C; E.ldc = ldc; E.alpha = 1.f; E.keep_scale = 1.f;
}

ConvGeom geom_of(const hulc2_conv_args* a) {
  ConvGeom g;
  g.C = a->C; g.H = a->H; g.W = a->W; g.KH = a->KH; g.KW = a->KW; g.stride = a->stride;
  g.OH = (a->H - a->KH) / a->stride + 1; g.OW = (a->W - a->KW) / a->stride + 1;
  g.nhwc = a->in_nhwc; g.Cout = a->Cout;
  return g;
}

}  // namespace

int hulc2_gemm_f32_impl(const hulc2_gemm_args* a, cudaStream_t st) {
  if (!a || a->M < 0 || a->N < 0 || a->K < 0 || !a->A || !a->B || !a->C) { hulc2_set_error("gemm: bad args"); return HULC2_EINVAL; }
  if (a->M == 0 || a->N == 0) return HULC2_OK;
  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.A.p = a->A; p.A.rs = a->a_rs; p.A.ks = a->a_ks; p.A.r_inner = a->a_inner; p.A.rs_outer = a->a_rs_outer; p.A.rs_inner = a->a_rs_inner;
  p.B.p = a->B; p.B.rs = a->b_rs; p.B.ks = a->b_ks;
  fill_epilogue(p.E, a->C, a->ldc);
  p.E.c_inner = a->c_inner; p.E.cs_outer = a->c_rs_outer; p.E.cs_inner = a->c_rs_inner;
  p.E.bias = a->bias; p.E.add = a->add; p.E.ld_add = a->ld_add; p.E.mask = a->mask; p.E.ld_mask = a->ld_mask;
  p.E.keep = a->keep; p.E.ld_keep = a->ld_keep; p.E.keep_scale = a->keep_scale;
  p.E.relu = a->relu; p.E.accumulate = a->accumulate; p.E.alpha = a->alpha;
  p.splits = 1; p.kchunk = ((a->K + BK - 1) / BK) * BK;
  if (p.kchunk == 0) p.kchunk = BK;
  // split-K for reductions over many rows into a small output (weight gradients)
  long long out_ctas = (long long)hulc2_cdiv(a->M, 64) * hulc2_cdiv(a->N, 64);
  if (a->workspace && a->K >= 2048 && out_ctas < 64 && !a->add && !a->mask && !a->keep && !a->relu) {
    int want = (int)((296 + out_ctas - 1) / out_ctas);
    int maxs = a->K / 256; if (maxs < 1) maxs = 1;
    int s = want < maxs ? want : maxs;
    long long need = (long long)s * a->M * a->N * sizeof(float);
    if (s > 1 && need <= a->workspace_bytes) {
      p.splits = s;
      p.kchunk = ((hulc2_cdiv(a->K, s) + BK - 1) / BK) * BK;
      p.splits = hulc2_cdiv(a->K, p.kchunk);
      p.partial = (float*)a->workspace;
    }
  }
  return dispatch(p, OP_DENSE, OP_DENSE, st);
}

// y[F,OH,OW,Cout] (NHWC) = act(conv(x, w) + bias).  x is NCHW (w OIHW, k=(ci,kh,kw)) or NHWC (w OHWI, k=(kh,kw,ci)).
int hulc2_conv2d_fwd_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  ConvGeom g = geom_of(a);
  GemmParams p{};
  p.M = a->F * g.OH * g.OW; p.N = a->Cout; p.K = a->C * a->KH * a->KW;
  if (p.M == 0) return HULC2_OK;
  p.A.p = a->x; p.A.g = g;
  p.B.p = a->w; p.B.rs = p.K; p.B.ks = 1;
  fill_epilogue(p.E, a->y, a->Cout);
  p.E.bias = a->bias; p.E.relu = a->relu;
  p.splits = 1; p.kchunk = ((p.K + BK - 1) / BK) * BK;
  return dispatch(p, OP_IM2COL, OP_DENSE, st);
}

// dw[Cout, K] (+)= dZ^T [Cout, pixels] . im2col(x) [pixels, K]   (same k-order / layout as the forward weight)
int hulc2_conv2d_wgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  ConvGeom g = geom_of(a);
  GemmParams p{};
  int pixels = a->F * g.OH * g.OW;
  p.M = a->Cout; p.N = a->C * a->KH * a->KW; p.K = pixels;
  if (pixels == 0) return HULC2_OK;
  p.A.p = a->dy; p.A.rs = 1; p.A.ks = a->Cout;       // A(m=co, k=pixel) = dZ[pixel*Cout + co]
  p.B.p = a->x; p.B.g = g;                            // B(n=kidx, k=pixel) = x[pix_off(pixel) + k_off(kidx)]
  fill_epilogue(p.E, a->dw, p.N);
  p.E.accumulate = a->accumulate;
  p.splits = 1; p.kchunk = ((p.K + BK - 1) / BK) * BK;
  long long out_ctas = (long long)hulc2_cdiv(p.M, 64) * hulc2_cdiv(p.N, 64);
  if (a->workspace && p.K >= 2048) {
    int want = (int)((592 + out_ctas - 1) / out_ctas);
    int maxs = p.K / 512; if (maxs < 1) maxs = 1;
    int s = want < maxs ? want : maxs;
    long long need = (long long)s * p.M * p.N * sizeof(float);
    while (s > 1 && need > a->workspace_bytes) { s /= 2; need = (long long)s * p.M * p.N * sizeof(float); }
    if (s > 1) {
      p.kchunk = ((hulc2_cdiv(p.K, s) + BK - 1) / BK) * BK;
      p.splits = hulc2_cdiv(p.K, p.kchunk);
      p.partial = (float*)a->workspace;
    }
  }
  return dispatch(p, OP_DENSE, OP_IM2COL_T, st);
}

// dx[F,H,W,C] (NHWC) = gather-conv(dZ[F,OH,OW,Cout], w_hwoi[KH,KW,Cout,C]) masked by (xmask > 0) (ReLU of the producer)
int hulc2_conv2d_dgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  ConvGeom g = geom_of(a);
  GemmParams p{};
  p.M = a->F * a->H * a->W; p.N = a->C; p.K = a->KH * a->KW * a->Cout;
  if (p.M == 0) return HULC2_OK;
  p.A.p = a->dy; p.A.g = g;
  p.B.p = a->w; p.B.rs = 1; p.B.ks = a->C;            // B(n=ci, k=(kh,kw,co)) = w_hwoi[k*C + ci]
  fill_epilogue(p.E, a->dx, a->C);
  p.E.mask = a->xmask; p.E.ld_mask = a->C;
  p.splits = 1; p.kchunk = ((p.K + BK - 1) / BK) * BK;
  return dispatch(p, OP_DGRAD, OP_DENSE, st);
}
