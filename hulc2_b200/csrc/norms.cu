// LayerNorm (+residual +dropout) and SpatialSoftmax, forward and backward.  Memory-bound: one pass
// over HBM per tensor, warp-shuffle reductions, coalesced along the feature/channel axis.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

constexpr int LN_VALS = 8;  // per-lane register slots: D <= 256

// --------------------------------------------------------------------------- LayerNorm
constexpr int LN_MAX_BLOCKS = 148 * 2;
__device__ float g_ln_part[2][LN_MAX_BLOCKS][LN_VALS * 32];     // per-block partial sums of dgamma / dbeta (backward)
__device__ unsigned int g_ln_count = 0;
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ res, long long ldr,
                                     const unsigned char* __restrict__ keep, float keep_scale,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                     long long ldy, float* __restrict__ tsum, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, long long rows, int D, float eps,
                                     __nv_bfloat16* __restrict__ y16, long long ld16) {
  int lane = threadIdx.x & 31;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float v[LN_VALS];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_VALS; ++j) {
      int c = lane + 32 * j;
      float t = 0.f;
      if (c < D) {
        t = x[r * ldx + c];
        if (res) {
          float rv = res[r * ldr + c];
          if (keep) rv = keep[r * D + c] ? rv * keep_scale : 0.f;
          t += rv;
        }
        if (tsum) tsum[r * D + c] = t;
      }
      v[j] = t;
      s += t;
    }
    float mu = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_VALS; ++j) {
      int c = lane + 32 * j;
      float d = (c < D) ? v[j] - mu : 0.f;
      q += d * d;
    }
    float rs = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int j = 0; j < LN_VALS; ++j) {
      int c = lane + 32 * j;
      if (c < D) {
        const float o = (v[j] - mu) * rs * gamma[c] + beta[c];
        y[r * ldy + c] = o;
        if (y16) y16[r * ld16 + c] = __float2bfloat16_rn(o);      // operand mirror for the contraction that consumes y
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[r] = mu;
      if (rstd_out) rstd_out[r] = rs;
    }
  }
}

__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, long long ldy, const float* __restrict__ t, long long ldt,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, float* __restrict__ dx, long long lddx,
                                     float* __restrict__ dres, const unsigned char* __restrict__ keep, float keep_scale,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int D,
                                     __nv_bfloat16* __restrict__ dx16, __nv_bfloat16* __restrict__ dres16, long long ld16) {
  __shared__ float sg[8][LN_VALS * 32];
  __shared__ float sb[8][LN_VALS * 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float ag[LN_VALS], ab[LN_VALS];
#pragma unroll
  for (int j = 0; j < LN_VALS; ++j) ag[j] = ab[j] = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    float mu = mean[r], rs = rstd[r];
    float xh[LN_VALS], g[LN_VALS];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_VALS; ++j) {
      int c = lane + 32 * j;
      xh[j] = g[j] = 0.f;
      if (c < D) {
        float d = dy[r * ldy + c];
        xh[j] = (t[r * ldt + c] - mu) * rs;
        g[j] = d * gamma[c];
        ag[j] += d * xh[j];
        ab[j] += d;
        c1 += g[j];
        c2 += g[j] * xh[j];
      }
    }
    c1 = warp_sum(c1) / (float)D;
    c2 = warp_sum(c2) / (float)D;
#pragma unroll
    for (int j = 0; j < LN_VALS; ++j) {
      int c = lane + 32 * j;
      if (c < D) {
        float v = rs * (g[j] - c1 - xh[j] * c2);
        dx[r * lddx + c] = v;
        if (dx16) dx16[r * ld16 + c] = __float2bfloat16_rn(v);
        if (dres) {
          const float dr = keep ? (keep[r * D + c] ? v * keep_scale : 0.f) : v;
          dres[r * D + c] = dr;
          if (dres16) dres16[r * ld16 + c] = __float2bfloat16_rn(dr);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < LN_VALS; ++j) { sg[w][lane + 32 * j] = ag[j]; sb[w][lane + 32 * j] = ab[j]; }
  __syncthreads();
  // Deterministic cross-block reduction: every block parks its partial sums, the block that arrives LAST adds them up in block
  // order.  (fp32 atomics here used to be the one source of run-to-run noise of a train step: the order of up to 296 adds per
  // element differed between two runs, and after a few Adam steps in bf16 the rounding differences grow chaotically -- the
  // graph-vs-eager parity tests had to tolerate that.)  The scratch is per device; LayerNorm backward kernels of one process run
  // on one stream, so they never overlap.
  __shared__ bool last_block;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += sg[k][c]; b += sb[k][c]; }
    g_ln_part[0][blockIdx.x][c] = a;
    g_ln_part[1][blockIdx.x][c] = b;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_block = atomicAdd(&g_ln_count, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) { a += __ldcg(&g_ln_part[0][k][c]); b += __ldcg(&g_ln_part[1][k][c]); }
    if (dgamma) dgamma[c] += a;
    if (dbeta) dbeta[c] += b;
  }
  if (threadIdx.x == 0) g_ln_count = 0;
}

// --------------------------------------------------------------------------- SpatialSoftmax (NHWC)
// block = one frame; thread (c, g): channel c, position group g of G = blockDim.x / C.
__device__ __forceinline__ float ld_act(const float* p) { return *p; }
__device__ __forceinline__ float ld_act(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_act(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// T = float (fp32 path) or __nv_bfloat16 (bf16 conv trunk: activations and their gradients live in HBM as bf16)
template <bool BWD, typename T>
__global__ void spatial_softmax_kernel(const T* __restrict__ x, const float* __restrict__ x_map,
                                       const float* __restrict__ y_map, const float* __restrict__ temperature,
                                       float* __restrict__ out, const float* __restrict__ dout, T* __restrict__ dx,
                                       float* __restrict__ dtemp, int HW, int C, int relu_mask) {
  extern __shared__ float sm[];  // [G][C] x 3
  const int G = blockDim.x / C;
  const int c = threadIdx.x % C, g = threadIdx.x / C;
  const bool active = g < G;
  const float invT = 1.f / temperature[0];
  const T* xf = x + (long long)blockIdx.x * HW * C;
  float* s0 = sm; float* s1 = sm + G * C; float* s2 = sm + 2 * G * C;

  float mx = -INFINITY;
  if (active) for (int i = g; i < HW; i += G) mx = fmaxf(mx, ld_act(xf + (long long)i * C + c) * invT);
  if (active) s0[g * C + c] = mx;
  __syncthreads();
  if (active) { mx = s0[c]; for (int k = 1; k < G; ++k) mx = fmaxf(mx, s0[k * C + c]); }
  __syncthreads();

  float se = 0.f, sx = 0.f, sy = 0.f;
  if (active)
    for (int i = g; i < HW; i += G) {
      float e = expf(ld_act(xf + (long long)i * C + c) * invT - mx);
      se += e; sx += e * x_map[i]; sy += e * y_map[i];
    }
  if (active) { s0[g * C + c] = se; s1[g * C + c] = sx; s2[g * C + c] = sy; }
  __syncthreads();
  if (active) {
    se = sx = sy = 0.f;
    for (int k = 0; k < G; ++k) { se += s0[k * C + c]; sx += s1[k * C + c]; sy += s2[k * C + c]; }
  }
  const float inv = active ? 1.f / se : 0.f;
  const float ex = sx * inv, ey = sy * inv;
  if (!BWD) {
    if (active && g == 0) {
      out[(long long)blockIdx.x * 2 * C + 2 * c] = ex;
      out[(long long)blockIdx.x * 2 * C + 2 * c + 1] = ey;
    }
    return;
  }
  if (!active) return;
  const float gx = dout[(long long)blockIdx.x * 2 * C + 2 * c], gy = dout[(long long)blockIdx.x * 2 * C + 2 * c + 1];
  T* dxf = dx + (long long)blockIdx.x * HW * C;
  float dt = 0.f;
  for (int i = g; i < HW; i += G) {
    float xv = ld_act(xf + (long long)i * C + c);
    float p = expf(xv * invT - mx) * inv;
    float dl = p * (gx * (x_map[i] - ex) + gy * (y_map[i] - ey));  // d/d(logit_i), logit = x/T
    dt += dl * xv;
    float v = dl * invT;
    if (relu_mask && !(xv > 0.f)) v = 0.f;
    st_act(dxf + (long long)i * C + c, v);
  }
  if (dtemp) atomicAdd(dtemp, -dt * invT * invT);
}

// bf16 trunk variant (exponentials with ex2.approx: the inputs are bf16 activations and the kernel was instruction-bound on the
// accurate expf -- 2 x 441 x 64 of them per frame and pass; the fp32 parity path keeps expf): the whole frame [HW][C] is staged in shared memory with 16-byte loads (one HBM read), every pass
// (max, sums, gradient) runs from shared memory, and the gradient is written back in place and stored with 16-byte
// stores (one HBM write).  Thread (cp, g): channel pair cp = tid % (C/2), position group g of G = blockDim / (C/2).
constexpr int SSM_NT = 256;   // threads per frame (512 with 3 frames per SM measured slower: 0.28 -> 0.32 ms backward)
template <bool BWD>
__global__ void __launch_bounds__(SSM_NT) ssm_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ x_map,
                                                       const float* __restrict__ y_map, const float* __restrict__ temperature,
                                                       float* __restrict__ out, const float* __restrict__ dout,
                                                       __nv_bfloat16* __restrict__ dx, float* __restrict__ dtemp, int HW, int C,
                                                       int relu_mask, float* __restrict__ stats) {
  // stats [F][C/2][4] = (max0, 1/sum0, max1, 1/sum1) per channel pair: written by the forward pass when given; a backward pass
  // that gets them (and the forward output in `out`) skips its two statistics passes over the frame and only runs the gradient pass
  extern __shared__ __align__(16) unsigned char ssm_raw[];
  __nv_bfloat162* tile = reinterpret_cast<__nv_bfloat162*>(ssm_raw);               // [HW][C/2]
  const int C2 = C >> 1, G = blockDim.x / C2;
  float* red = reinterpret_cast<float*>(ssm_raw + (((size_t)HW * C * 2 + 15) & ~(size_t)15));   // [6][G][C2]
  float* maps = red + 6 * G * C2;                                                   // x_map | y_map
  const int n16 = HW * C / 8;
  const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)blockIdx.x * HW * C);
  for (int q = threadIdx.x; q < n16; q += blockDim.x) reinterpret_cast<uint4*>(ssm_raw)[q] = __ldg(src + q);
  for (int q = threadIdx.x; q < HW; q += blockDim.x) { maps[q] = x_map[q]; maps[HW + q] = y_map[q]; }
  const float invT = 1.f / temperature[0];
  __syncthreads();
  const int cp = threadIdx.x % C2, g = threadIdx.x / C2;
  const bool active = g < G;
  float m0 = -INFINITY, m1 = -INFINITY;
  float inv0, inv1, ex0, ey0, ex1, ey1;
  const bool have_stats = BWD && stats != nullptr;
  if (have_stats) {
    const float4 st = *reinterpret_cast<const float4*>(stats + (size_t)blockIdx.x * 2 * C + 4 * cp);
    const float4 eo = *reinterpret_cast<const float4*>(out + (size_t)blockIdx.x * 2 * C + 4 * cp);
    m0 = st.x; inv0 = st.y; m1 = st.z; inv1 = st.w; ex0 = eo.x; ey0 = eo.y; ex1 = eo.z; ey1 = eo.w;
  } else {
  if (active)
    for (int i = g; i < HW; i += G) {
      const float2 v = __bfloat1622float2(tile[i * C2 + cp]);
      m0 = fmaxf(m0, v.x * invT); m1 = fmaxf(m1, v.y * invT);
    }
  if (active) { red[g * C2 + cp] = m0; red[(G + g) * C2 + cp] = m1; }
  __syncthreads();
  if (active)
    for (int k = 0; k < G; ++k) { m0 = fmaxf(m0, red[k * C2 + cp]); m1 = fmaxf(m1, red[(G + k) * C2 + cp]); }
  __syncthreads();
  float se0 = 0.f, sx0 = 0.f, sy0 = 0.f, se1 = 0.f, sx1 = 0.f, sy1 = 0.f;
  if (active)
    for (int i = g; i < HW; i += G) {
      const float2 v = __bfloat1622float2(tile[i * C2 + cp]);
      const float xm = maps[i], ym = maps[HW + i];
      const float e0 = __expf(v.x * invT - m0), e1 = __expf(v.y * invT - m1);
      se0 += e0; sx0 += e0 * xm; sy0 += e0 * ym;
      se1 += e1; sx1 += e1 * xm; sy1 += e1 * ym;
    }
  if (active) {
    float* r = red + g * C2 + cp;
    r[0] = se0; r[G * C2] = sx0; r[2 * G * C2] = sy0; r[3 * G * C2] = se1; r[4 * G * C2] = sx1; r[5 * G * C2] = sy1;
  }
  __syncthreads();
  if (active) {
    se0 = sx0 = sy0 = se1 = sx1 = sy1 = 0.f;
    for (int k = 0; k < G; ++k) {
      const float* r = red + k * C2 + cp;
      se0 += r[0]; sx0 += r[G * C2]; sy0 += r[2 * G * C2]; se1 += r[3 * G * C2]; sx1 += r[4 * G * C2]; sy1 += r[5 * G * C2];
    }
  }
  inv0 = active ? 1.f / se0 : 0.f; inv1 = active ? 1.f / se1 : 0.f;
  ex0 = sx0 * inv0; ey0 = sy0 * inv0; ex1 = sx1 * inv1; ey1 = sy1 * inv1;
  }   // !have_stats
  if (!BWD) {
    if (active && g == 0) {
      *reinterpret_cast<float4*>(out + (size_t)blockIdx.x * 2 * C + 4 * cp) = make_float4(ex0, ey0, ex1, ey1);
      if (stats) *reinterpret_cast<float4*>(stats + (size_t)blockIdx.x * 2 * C + 4 * cp) = make_float4(m0, inv0, m1, inv1);
    }
    return;
  }
  float dt = 0.f;
  if (active) {
    const float4 gq = *reinterpret_cast<const float4*>(dout + (size_t)blockIdx.x * 2 * C + 4 * cp);   // (gx0, gy0, gx1, gy1)
    for (int i = g; i < HW; i += G) {
      const float2 v = __bfloat1622float2(tile[i * C2 + cp]);
      const float xm = maps[i], ym = maps[HW + i];
      const float p0 = __expf(v.x * invT - m0) * inv0, p1 = __expf(v.y * invT - m1) * inv1;
      const float dl0 = p0 * (gq.x * (xm - ex0) + gq.y * (ym - ey0)), dl1 = p1 * (gq.z * (xm - ex1) + gq.w * (ym - ey1));
      dt += dl0 * v.x + dl1 * v.y;
      float o0 = dl0 * invT, o1 = dl1 * invT;
      if (relu_mask) { if (!(v.x > 0.f)) o0 = 0.f; if (!(v.y > 0.f)) o1 = 0.f; }
      tile[i * C2 + cp] = __floats2bfloat162_rn(o0, o1);        // only this thread ever touches element (i, cp)
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(dx + (size_t)blockIdx.x * HW * C);
  for (int q = threadIdx.x; q < n16; q += blockDim.x) dst[q] = reinterpret_cast<const uint4*>(ssm_raw)[q];
  if (dtemp) {
    dt = block_sum(dt, red);
    if (threadIdx.x == 0) atomicAdd(dtemp, -dt * invT * invT);
  }
}

// Streaming SpatialSoftmax for the 64-channel trunk (third generation).  ncu r02 showed the shared-memory kernel above
// issue-bound (71 % of issue slots, 28-30 % of HBM): a frame was staged with LDG.128 + STS.128, then walked two or three times
// with LDS per channel pair.  Here thread (lane = channel pair, warp = position group g of 8) loads ITS OWN bf16 pairs straight
// from HBM -- a warp reads one 128-byte position row per load, fully coalesced -- and every element is touched once;
// exponentials are one FFMA + one ex2.approx (1/T and log2 e folded into the scale).  No frame staging, no barrier before the
// arithmetic, ~3x fewer instructions per element.  The maps are read through L1 (every lane of a warp reads the same
// address).  stats keep the format of ssm_bf16_kernel: (max of x/T, 1/sum) per channel.
constexpr float SSM_LOG2E = 1.4426950408889634f, SSM_LN2 = 0.6931471805599453f;
__device__ __forceinline__ float ssm_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// forward: ONE streaming pass with an online maximum.  Positions are taken in batches of 8 per thread (8 independent 128-byte
// row loads in flight per warp); per batch the running maximum moves at most once, so rescaling the three running sums costs one
// ex2 + three multiplies per 8 positions.  ~45 registers: 5 blocks per SM.
__global__ void __launch_bounds__(256) ssm_reg_fwd_kernel(const uint32_t* __restrict__ x, const float* __restrict__ x_map,
                                                          const float* __restrict__ y_map, const float* __restrict__ temperature,
                                                          float* __restrict__ out, float* __restrict__ stats, int HW) {
  __shared__ float red[8][8][32];        // [m0, se0, sx0, sy0, m1, se1, sx1, sy1][warp][lane]
  const int cp = threadIdx.x & 31, g = threadIdx.x >> 5;
  const uint32_t* xf = x + (size_t)blockIdx.x * HW * 32 + cp;
  const float a = SSM_LOG2E / temperature[0];
  float m0 = -INFINITY, m1 = -INFINITY;
  float se0 = 0.f, sx0 = 0.f, sy0 = 0.f, se1 = 0.f, sx1 = 0.f, sy1 = 0.f;
  for (int i0 = g; i0 < HW; i0 += 64) {
    uint32_t u[8];
    float xm[8], ym[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + 8 * k;
      const bool ok = i < HW;
      u[k] = ok ? __ldg(xf + (size_t)i * 32) : 0u;
      xm[k] = ok ? __ldg(x_map + i) : 0.f;
      ym[k] = ok ? __ldg(y_map + i) : 0.f;
    }
    float b0 = m0, b1 = m1;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (i0 + 8 * k < HW) { b0 = fmaxf(b0, bf16lo(u[k]) * a); b1 = fmaxf(b1, bf16hi(u[k]) * a); }
    // (first batch: m = -inf, sums = 0: ex2(-inf) = 0 keeps them 0)
    const float r0 = ssm_ex2(m0 - b0), r1 = ssm_ex2(m1 - b1);
    se0 *= r0; sx0 *= r0; sy0 *= r0; se1 *= r1; sx1 *= r1; sy1 *= r1;
    m0 = b0; m1 = b1;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (i0 + 8 * k < HW) {
        const float e0 = ssm_ex2(fmaf(bf16lo(u[k]), a, -m0)), e1 = ssm_ex2(fmaf(bf16hi(u[k]), a, -m1));
        se0 += e0; sx0 = fmaf(e0, xm[k], sx0); sy0 = fmaf(e0, ym[k], sy0);
        se1 += e1; sx1 = fmaf(e1, xm[k], sx1); sy1 = fmaf(e1, ym[k], sy1);
      }
  }
  red[0][g][cp] = m0; red[1][g][cp] = se0; red[2][g][cp] = sx0; red[3][g][cp] = sy0;
  red[4][g][cp] = m1; red[5][g][cp] = se1; red[6][g][cp] = sx1; red[7][g][cp] = sy1;
  __syncthreads();
  if (g == 0) {
    float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) { M0 = fmaxf(M0, red[0][k][cp]); M1 = fmaxf(M1, red[4][k][cp]); }
    se0 = sx0 = sy0 = se1 = sx1 = sy1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {        // a warp that saw no position has m = -inf, sums 0: weight ex2(-inf) = 0
      const float w0 = ssm_ex2(red[0][k][cp] - M0), w1 = ssm_ex2(red[4][k][cp] - M1);
      se0 = fmaf(w0, red[1][k][cp], se0); sx0 = fmaf(w0, red[2][k][cp], sx0); sy0 = fmaf(w0, red[3][k][cp], sy0);
      se1 = fmaf(w1, red[5][k][cp], se1); sx1 = fmaf(w1, red[6][k][cp], sx1); sy1 = fmaf(w1, red[7][k][cp], sy1);
    }
    const float inv0 = 1.f / se0, inv1 = 1.f / se1;
    *reinterpret_cast<float4*>(out + (size_t)blockIdx.x * 128 + 4 * cp) = make_float4(sx0 * inv0, sy0 * inv0, sx1 * inv1, sy1 * inv1);
    if (stats) *reinterpret_cast<float4*>(stats + (size_t)blockIdx.x * 128 + 4 * cp) = make_float4(M0 * SSM_LN2, inv0, M1 * SSM_LN2, inv1);
  }
}

// backward from the saved statistics: ONE streaming pass, x in -> dz out (ReLU mask of the producing conv fused), nothing staged;
// loads are issued 8 positions ahead of the arithmetic / stores (the compiler does not move loads across the stores by itself)
__global__ void __launch_bounds__(256) ssm_reg_bwd_kernel(const uint32_t* __restrict__ x, const float* __restrict__ x_map,
                                                          const float* __restrict__ y_map, const float* __restrict__ temperature,
                                                          const float* __restrict__ out, const float* __restrict__ stats,
                                                          const float* __restrict__ dout, uint32_t* __restrict__ dx,
                                                          float* __restrict__ dtemp, int HW, int relu_mask) {
  __shared__ float red[32];
  const int cp = threadIdx.x & 31, g = threadIdx.x >> 5;
  const size_t fo = (size_t)blockIdx.x * 128 + 4 * cp;
  const float4 st = *reinterpret_cast<const float4*>(stats + fo);      // (max0, 1/sum0, max1, 1/sum1), natural-log units
  const float4 eo = *reinterpret_cast<const float4*>(out + fo);        // (ex0, ey0, ex1, ey1)
  const float4 gq = *reinterpret_cast<const float4*>(dout + fo);       // (gx0, gy0, gx1, gy1)
  const float invT = 1.f / temperature[0], a = SSM_LOG2E * invT;
  const float mb0 = st.x * SSM_LOG2E, mb1 = st.z * SSM_LOG2E;
  // dl = p (gx (xm - ex) + gy (ym - ey)) = p (gx xm + gy ym - c) with c = gx ex + gy ey
  const float c0 = gq.x * eo.x + gq.y * eo.y, c1 = gq.z * eo.z + gq.w * eo.w;
  const uint32_t* xf = x + (size_t)blockIdx.x * HW * 32 + cp;
  uint32_t* df = dx + (size_t)blockIdx.x * HW * 32 + cp;
  float dt = 0.f;
  for (int i0 = g; i0 < HW; i0 += 64) {
    uint32_t u[8];
    float xm[8], ym[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + 8 * k;
      const bool ok = i < HW;
      u[k] = ok ? __ldg(xf + (size_t)i * 32) : 0u;
      xm[k] = ok ? __ldg(x_map + i) : 0.f;
      ym[k] = ok ? __ldg(y_map + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + 8 * k;
      if (i < HW) {
        const float v0 = bf16lo(u[k]), v1 = bf16hi(u[k]);
        const float p0 = ssm_ex2(fmaf(v0, a, -mb0)) * st.y, p1 = ssm_ex2(fmaf(v1, a, -mb1)) * st.w;
        const float dl0 = p0 * (fmaf(gq.x, xm[k], gq.y * ym[k]) - c0), dl1 = p1 * (fmaf(gq.z, xm[k], gq.w * ym[k]) - c1);
        dt = fmaf(dl0, v0, fmaf(dl1, v1, dt));
        float o0 = dl0 * invT, o1 = dl1 * invT;
        if (relu_mask) { if (!(v0 > 0.f)) o0 = 0.f; if (!(v1 > 0.f)) o1 = 0.f; }
        const __nv_bfloat162 o = __floats2bfloat162_rn(o0, o1);
        df[(size_t)i * 32] = *reinterpret_cast<const uint32_t*>(&o);
      }
    }
  }
  if (dtemp) {
    dt = block_sum(dt, red);
    if (threadIdx.x == 0) atomicAdd(dtemp, -dt * invT * invT);
  }
}

static int ssm_reg_np(int HW, int C) {           // 0: shape not served by the streaming kernels (one warp = one 64-channel position row)
  return (C == 64 && HW > 0) ? 1 : 0;
}
static bool ssm_reg_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HULC2_SSM_REG"); v = (e && e[0] == '0') ? 0 : 1; }   // A/B switch, read once
  return v == 1;
}

static size_t ssm_bf16_smem(int HW, int C) {
  const int C2 = C / 2, G = SSM_NT / C2;
  return (((size_t)HW * C * 2 + 15) & ~(size_t)15) + (size_t)(6 * G * C2 + 2 * HW) * sizeof(float);
}
// served when channel pairs tile a 256-thread block, rows are 16-byte multiples and 2 frames fit one SM's shared memory
static bool ssm_bf16_ok(int HW, int C) {
  return C >= 8 && C % 8 == 0 && C <= 512 && SSM_NT % (C / 2) == 0 && ssm_bf16_smem(HW, C) <= 100 * 1024;
}

}  // namespace

extern "C" {

int hulc2_layernorm_fwd_m(const float* x, long long ldx, const float* res, long long ldr, const unsigned char* keep,
                          float keep_scale, const float* gamma, const float* beta, float* y, long long ldy, float* tsum,
                          float* mean, float* rstd, long long rows, int D, float eps, void* y16, long long ld16, cudaStream_t st) {
  if (rows <= 0) return HULC2_OK;
  if (D > 32 * LN_VALS || D <= 0) { hulc2_set_error("layernorm: D must be in (0,256]"); return HULC2_EINVAL; }
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  layernorm_fwd_kernel<<<(int)blocks, 256, 0, st>>>(x, ldx, res, ldr, keep, keep_scale, gamma, beta, y, ldy, tsum, mean, rstd, rows, D, eps,
                                                    (__nv_bfloat16*)y16, ld16);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_layernorm_fwd(const float* x, long long ldx, const float* res, long long ldr, const unsigned char* keep,
                        float keep_scale, const float* gamma, const float* beta, float* y, long long ldy, float* tsum,
                        float* mean, float* rstd, long long rows, int D, float eps, cudaStream_t st) {
  return hulc2_layernorm_fwd_m(x, ldx, res, ldr, keep, keep_scale, gamma, beta, y, ldy, tsum, mean, rstd, rows, D, eps, nullptr, 0, st);
}

int hulc2_layernorm_bwd_m(const float* dy, long long ldy, const float* t, long long ldt, const float* gamma, const float* mean,
                          const float* rstd, float* dx, long long lddx, float* dres, const unsigned char* keep, float keep_scale,
                          float* dgamma, float* dbeta, long long rows, int D, void* dx16, void* dres16, long long ld16,
                          cudaStream_t st) {
  if (rows <= 0) return HULC2_OK;
  if (D > 32 * LN_VALS || D <= 0) { hulc2_set_error("layernorm: D must be in (0,256]"); return HULC2_EINVAL; }
  long long blocks = (rows + 7) / 8;
  if (blocks > LN_MAX_BLOCKS) blocks = LN_MAX_BLOCKS;
  layernorm_bwd_kernel<<<(int)blocks, 256, 0, st>>>(dy, ldy, t, ldt, gamma, mean, rstd, dx, lddx, dres, keep, keep_scale, dgamma, dbeta, rows, D,
                                                    (__nv_bfloat16*)dx16, (__nv_bfloat16*)dres16, ld16);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_layernorm_bwd(const float* dy, long long ldy, const float* t, long long ldt, const float* gamma, const float* mean,
                        const float* rstd, float* dx, long long lddx, float* dres, const unsigned char* keep, float keep_scale,
                        float* dgamma, float* dbeta, long long rows, int D, cudaStream_t st) {
  return hulc2_layernorm_bwd_m(dy, ldy, t, ldt, gamma, mean, rstd, dx, lddx, dres, keep, keep_scale, dgamma, dbeta, rows, D, nullptr,
                               nullptr, 0, st);
}

static int ssm_threads(int C) {
  if (C > 256) return 0;
  int G = 256 / C;
  return G * C > 0 ? 256 : 0;
}

int hulc2_spatial_softmax_fwd(const float* x, const float* x_map, const float* y_map, const float* temperature, float* out,
                              int F, int HW, int C, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (!ssm_threads(C)) { hulc2_set_error("spatial_softmax: C must be <= 256"); return HULC2_EINVAL; }
  int G = 256 / C;
  spatial_softmax_kernel<false, float><<<F, 256, 3 * G * C * sizeof(float), st>>>(x, x_map, y_map, temperature, out, nullptr, nullptr, nullptr, HW, C, 0);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_spatial_softmax_bwd(const float* x, const float* x_map, const float* y_map, const float* temperature,
                              const float* out, const float* dout, float* dx, float* dtemperature, int F, int HW, int C,
                              int relu_mask, cudaStream_t st) {
  (void)out;
  if (F <= 0) return HULC2_OK;
  if (!ssm_threads(C)) { hulc2_set_error("spatial_softmax: C must be <= 256"); return HULC2_EINVAL; }
  int G = 256 / C;
  spatial_softmax_kernel<true, float><<<F, 256, 3 * G * C * sizeof(float), st>>>(x, x_map, y_map, temperature, nullptr, dout, dx, dtemperature, HW, C, relu_mask);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}


int hulc2_spatial_softmax_fwd_bf16(const void* x, const float* x_map, const float* y_map, const float* temperature, float* out,
                                   int F, int HW, int C, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (!ssm_threads(C)) { hulc2_set_error("spatial_softmax: C must be <= 256"); return HULC2_EINVAL; }
  if (const int np = ssm_reg_enabled() ? ssm_reg_np(HW, C) : 0) {
    const uint32_t* xu = (const uint32_t*)x;
    (void)np;
    ssm_reg_fwd_kernel<<<F, 256, 0, st>>>(xu, x_map, y_map, temperature, out, nullptr, HW);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  if (ssm_bf16_ok(HW, C)) {
    const size_t smem = ssm_bf16_smem(HW, C);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(ssm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
    ssm_bf16_kernel<false><<<F, SSM_NT, smem, st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, out, nullptr, nullptr, nullptr, HW, C, 0, nullptr);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  int G = 256 / C;
  spatial_softmax_kernel<false, __nv_bfloat16><<<F, 256, 3 * G * C * sizeof(float), st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, out,
                                                                                          nullptr, nullptr, nullptr, HW, C, 0);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_spatial_softmax_bwd_bf16(const void* x, const float* x_map, const float* y_map, const float* temperature,
                                   const float* dout, void* dx, float* dtemperature, int F, int HW, int C, int relu_mask,
                                   cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (!ssm_threads(C)) { hulc2_set_error("spatial_softmax: C must be <= 256"); return HULC2_EINVAL; }
  if (ssm_bf16_ok(HW, C)) {
    const size_t smem = ssm_bf16_smem(HW, C);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(ssm_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
    ssm_bf16_kernel<true><<<F, SSM_NT, smem, st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, nullptr, dout, (__nv_bfloat16*)dx, dtemperature,
                                                HW, C, relu_mask, nullptr);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  int G = 256 / C;
  spatial_softmax_kernel<true, __nv_bfloat16><<<F, 256, 3 * G * C * sizeof(float), st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, nullptr,
                                                                                         dout, (__nv_bfloat16*)dx, dtemperature, HW, C, relu_mask);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

// Forward that also saves the per-(frame, channel) softmax statistics, and the backward that consumes them together with the
// forward output: one pass over the frame instead of three (max, sums, gradient).  Only the bf16 shared-memory kernel serves
// these (hulc2_spatial_softmax_stats_supported); stats is fp32 [F, 2 C].
int hulc2_spatial_softmax_stats_supported(int HW, int C) { return ssm_bf16_ok(HW, C) ? 1 : 0; }
int hulc2_spatial_softmax_fwd_bf16_stats(const void* x, const float* x_map, const float* y_map, const float* temperature, float* out,
                                         float* stats, int F, int HW, int C, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (!ssm_bf16_ok(HW, C) || !stats) { hulc2_set_error("spatial_softmax_fwd_bf16_stats: unsupported shape"); return HULC2_ENOTIMPL; }
  if (const int np = ssm_reg_enabled() ? ssm_reg_np(HW, C) : 0) {
    const uint32_t* xu = (const uint32_t*)x;
    (void)np;
    ssm_reg_fwd_kernel<<<F, 256, 0, st>>>(xu, x_map, y_map, temperature, out, stats, HW);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  const size_t smem = ssm_bf16_smem(HW, C);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ssm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
  ssm_bf16_kernel<false><<<F, SSM_NT, smem, st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, out, nullptr, nullptr, nullptr, HW, C, 0, stats);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_spatial_softmax_bwd_bf16_stats(const void* x, const float* x_map, const float* y_map, const float* temperature, const float* out,
                                         const float* stats, const float* dout, void* dx, float* dtemperature, int F, int HW, int C,
                                         int relu_mask, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (!ssm_bf16_ok(HW, C) || !stats || !out) { hulc2_set_error("spatial_softmax_bwd_bf16_stats: unsupported shape"); return HULC2_ENOTIMPL; }
  if (ssm_reg_enabled() && ssm_reg_np(HW, C)) {
    ssm_reg_bwd_kernel<<<F, 256, 0, st>>>((const uint32_t*)x, x_map, y_map, temperature, out, stats, dout, (uint32_t*)dx, dtemperature, HW, relu_mask);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  const size_t smem = ssm_bf16_smem(HW, C);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ssm_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; }
  ssm_bf16_kernel<true><<<F, SSM_NT, smem, st>>>((const __nv_bfloat16*)x, x_map, y_map, temperature, const_cast<float*>(out), dout, (__nv_bfloat16*)dx,
                                              dtemperature, HW, C, relu_mask, const_cast<float*>(stats));
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
