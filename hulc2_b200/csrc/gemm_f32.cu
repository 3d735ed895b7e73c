// fp32 CUDA-core GEMM with implicit-convolution operand loaders (the exact-parity path).
//
// One register-tiled kernel serves every dense contraction of the HULC++ policy step in fp32:
//   Linear fwd / dgrad / wgrad          (goal encoders, plan proposal, transformer, RNN, heads)
//   conv fwd   = implicit GEMM, A = im2col(x)            (vision_network.py:38-48, vision_network_gripper.py:11-26)
//   conv wgrad = implicit GEMM, B = im2col(x)^T, split-K over output pixels
//   conv dgrad = implicit GEMM, A = strided gather of dZ
// Operand addressing and the epilogue are shared with the bf16 tcgen05 backend (gemm_common.cuh).
#include "gemm_common.cuh"

using namespace hulc2;

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

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
          dgrad_row_decode(o.g, rr, rdec[i][0], rdec[i][1], rdec[i][2]);
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
        int ta, tb, co;
        dgrad_k_decode(o.g, kv ? kg : 0, ta, tb, co);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          long long off;
          v[i] = (kv && rvalid[i] && dgrad_src(o.g, rdec[i][0], rdec[i][1], rdec[i][2], ta, tb, off)) ? __ldg(o.p + off + co) : 0.f;
        }
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
    const int mp = phys_row(E, m);
    long long crow = c_row_off(E, mp);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tile_idx<TN, BN>(tx, j);
      if (n >= p.N) continue;
      E.C[crow + n] = apply_epilogue(E, acc[i][j], mp, n, crow);
    }
  }
}

template <int BM, int BN, int TM, int TN, int AMODE, bool AKF, int BMODE, bool BKF>
int launch_cfg(const GemmParams& p, cudaStream_t st) {
  dim3 grid(hulc2_cdiv(p.M, BM), hulc2_cdiv(p.N, BN), p.splits);
  gemm_f32_kernel<BM, BN, TM, TN, AMODE, AKF, BMODE, BKF><<<grid, NTHREADS, 0, st>>>(p);
  HULC2_CHECK_LAUNCH();
  if (p.splits > 1) {
    long long total = (long long)p.M * p.N;
    launch_splitk_reduce(p.partial, p.splits, p.M, p.N, p.E, nullptr, 0, nullptr, nullptr, st);
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
  if (amode == OP_DGRAD && bmode == OP_DGRAD_W) return launch_modes<OP_DGRAD, true, OP_DGRAD_W, false>(p, st);
  hulc2_set_error("gemm_f32: unsupported operand mode combination");
  return HULC2_EINVAL;
}

}  // namespace

int hulc2_gemm_f32_impl(const hulc2_gemm_args* a, cudaStream_t st) {
  GemmParams p;
  if (!dense_params(a, p)) return HULC2_EINVAL;
  if (a->M == 0 || a->N == 0) return HULC2_OK;
  long long out_ctas = (long long)hulc2_cdiv(a->M, 64) * hulc2_cdiv(a->N, 64);
  // split-K for reductions over many rows into a small output (weight gradients)
  // (the reduce kernel applies the full epilogue, so layers with bias / ReLU / mask split as well: e.g. the fp32 contrastive
  // projection 64 x 128 x 4096 would otherwise run on 8 CTAs)
  if (out_ctas < 64) plan_splitk(p, out_ctas, 296, 256, BK, a->workspace, a->workspace_bytes);
  else plan_splitk(p, 1, 1, 1 << 30, BK, nullptr, 0);
  return dispatch(p, OP_DENSE, OP_DENSE, st);
}

// y[F,OH,OW,Cout] (NHWC) = act(conv(x, w) + bias).  x is NCHW (w OIHW, k=(ci,kh,kw)) or NHWC (w OHWI, k=(kh,kw,ci)).
int hulc2_conv2d_fwd_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  GemmParams p;
  conv_fwd_params(a, p);
  if (p.M == 0) return HULC2_OK;
  plan_splitk(p, 1, 1, 1 << 30, BK, nullptr, 0);
  return dispatch(p, OP_IM2COL, OP_DENSE, st);
}

// dw[Cout, K] (+)= dZ^T [Cout, pixels] . im2col(x) [pixels, K]   (same k-order / layout as the forward weight)
int hulc2_conv2d_wgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  GemmParams p;
  conv_wgrad_params(a, p);
  if (p.K == 0) return HULC2_OK;
  long long out_ctas = (long long)hulc2_cdiv(p.M, 64) * hulc2_cdiv(p.N, 64);
  plan_splitk(p, out_ctas, 592, 512, BK, a->workspace, a->workspace_bytes);
  return dispatch(p, OP_DENSE, OP_IM2COL_T, st);
}

// dx[F,H,W,C] (NHWC) = gather-conv(dZ[F,OH,OW,Cout], w_hwoi[KH,KW,Cout,C]) masked by (xmask > 0) (ReLU of the producer)
int hulc2_conv2d_dgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st) {
  for (int ph = 0; ph < a->stride; ++ph)
    for (int pw = 0; pw < a->stride; ++pw) {
      GemmParams p;
      if (!conv_dgrad_params(a, ph, pw, p)) continue;
      plan_splitk(p, 1, 1, 1 << 30, BK, nullptr, 0);
      if (int e = dispatch(p, OP_DGRAD, OP_DGRAD_W, st)) return e;
    }
  return HULC2_OK;
}
