// Plan-recognition transformer pieces that are not GEMMs (plan_recognition_net.py:125-148 and
// torch nn.TransformerEncoderLayer): positional add + dropout, 32x32-per-head attention held entirely
// in registers/shared memory (one warp per (window, head)), sequence mean.
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

__global__ void add_pos_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ pos,
                                   const unsigned char* __restrict__ keep, float keep_scale, float* __restrict__ out,
                                   __nv_bfloat16* __restrict__ out16, long long total, int SE) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = emb[i] + pos[i % SE];
    if (keep) v = keep[i] ? v * keep_scale : 0.f;
    out[i] = v;
    if (out16) out16[i] = __float2bfloat16(v);      // operand mirror of the QKV contraction (same rounding as f32_to_bf16)
  }
}

// thread per (s,e): demb[b,s,e] = dout*keep*scale for all b; dpos[s,e] += sum_b demb
// block = 32 columns x 8 window lanes: every thread takes every 8th window of its column (loads issued 8 ahead of the stores),
// the 8 partial sums are added in lane order through shared memory (deterministic; one thread per column used to walk all B
// windows serially: 33 us for 2 MB)
__global__ void __launch_bounds__(256) add_pos_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ keep, float keep_scale,
                                                          float* __restrict__ demb, float* __restrict__ dpos, int B, int SE) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < SE) {
    for (int b0 = threadIdx.y; b0 < B; b0 += 64) {
      float v[8];
      unsigned char k[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + 8 * u;
        const long long o = (long long)b * SE + i;
        v[u] = b < B ? dout[o] : 0.f;
        k[u] = (keep && b < B) ? keep[o] : (unsigned char)1;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + 8 * u;
        if (b < B) {
          const float x = keep ? (k[u] ? v[u] * keep_scale : 0.f) : v[u];
          if (demb) demb[(long long)b * SE + i] = x;
          s += x;
        }
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < SE && dpos) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    dpos[i] += t;
  }
}

// Row helpers: a lane owns one 32-element row of the probability / keep matrices of its (window, head).  All loads of a row
// are issued back to back BEFORE the arithmetic that consumes them (r02: the keep byte and the saved probability used to be
// loaded inside the key loop, right in front of their use -- 32 exposed L2 round trips per lane; the two kernels were pure
// load latency: 23 / 47 us for 13 / 25 MB).  VEC: S == 32 and 16-byte aligned bases -> 16-byte accesses.
template <bool VEC>
__device__ __forceinline__ void load_keep_row(const unsigned char* __restrict__ keep, long long pbase, int S, unsigned (&kw)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) kw[q] = 0x01010101u;
  if (!keep) return;
  if (VEC) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(keep + pbase)), b = __ldg(reinterpret_cast<const uint4*>(keep + pbase) + 1);
    kw[0] = a.x; kw[1] = a.y; kw[2] = a.z; kw[3] = a.w; kw[4] = b.x; kw[5] = b.y; kw[6] = b.z; kw[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const unsigned v = (j < S) ? (unsigned)__ldg(keep + pbase + j) : 0u;
      if (j % 4 == 0) kw[j / 4] = 0u;
      kw[j / 4] |= (v ? 1u : 0u) << (8 * (j % 4));
    }
  }
}
__device__ __forceinline__ bool kept(const unsigned (&kw)[8], int j) { return ((kw[j >> 2] >> (8 * (j & 3))) & 0xFFu) != 0u; }

template <int DH>
__device__ __forceinline__ void load_head_row(const float* __restrict__ row, float (&v)[DH]) {
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(row + d));
    v[d] = x.x; v[d + 1] = x.y; v[d + 2] = x.z; v[d + 3] = x.w;
  }
}

// DH consecutive bf16 values (DH % 8 == 0, 16-byte aligned destination when ld16 % 8 == 0) as 16-byte stores
template <int DH>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const float (&v)[DH], float mul, bool aligned) {
  if (aligned) {
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[d] * mul, v[d + 1] * mul), h1 = __floats2bfloat162_rn(v[d + 2] * mul, v[d + 3] * mul);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[d + 4] * mul, v[d + 5] * mul), h3 = __floats2bfloat162_rn(v[d + 6] * mul, v[d + 7] * mul);
      uint4 o;
      o.x = *reinterpret_cast<unsigned*>(&h0); o.y = *reinterpret_cast<unsigned*>(&h1);
      o.z = *reinterpret_cast<unsigned*>(&h2); o.w = *reinterpret_cast<unsigned*>(&h3);
      *reinterpret_cast<uint4*>(dst + d) = o;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DH; ++d) dst[d] = __float2bfloat16_rn(v[d] * mul);
  }
}

template <int DH, int WARPS, bool VEC>
__global__ void attention_fwd_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ keep, float keep_scale,
                                     float* __restrict__ out, float* __restrict__ probs, __nv_bfloat16* __restrict__ out16,
                                     long long ld16, int B, int S, int H) {
  __shared__ float Ks[WARPS][32][DH + 1];
  __shared__ float Vs[WARPS][32][DH + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x * WARPS + w;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * DH;
  const float scale = rsqrtf((float)DH);
  const long long pbase = ((long long)bh * S + lane) * S;
  float q[DH];
  unsigned kw[8];
  if (lane < S) {
    const float* row = qkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
    float kr[DH], vr[DH];
    load_head_row<DH>(row, q);
    load_head_row<DH>(row + E, kr);
    load_head_row<DH>(row + 2 * E, vr);
    load_keep_row<VEC>(keep, pbase, S, kw);
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      q[d] *= scale;
      Ks[w][lane][d] = kr[d];
      Vs[w][lane][d] = vr[d];
    }
  }
  __syncwarp();
  if (lane >= S) return;
  float s[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float a = -INFINITY;
    if (j < S) {
      a = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) a = fmaf(q[d], Ks[w][j][d], a);
    }
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s[j] = (j < S) ? expf(s[j] - mx) : 0.f;
    sum += s[j];
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] *= inv;                 // (0 for j >= S)
  if (probs) {
    if (VEC) {
#pragma unroll
      for (int qd = 0; qd < 8; ++qd)
        reinterpret_cast<float4*>(probs + pbase)[qd] = make_float4(s[4 * qd], s[4 * qd + 1], s[4 * qd + 2], s[4 * qd + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < S) probs[pbase + j] = s[j];
    }
  }
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < S) {
      float p = s[j];
      if (keep) p = kept(kw, j) ? p * keep_scale : 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, Vs[w][j][d], o[d]);
    }
  }
  float* orow = out + ((long long)(b * S + lane)) * E + h * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 4) *reinterpret_cast<float4*>(orow + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
  if (out16) {      // operand mirror of the output projection
    __nv_bfloat16* o16 = out16 + ((long long)(b * S + lane)) * ld16 + h * DH;
    store_bf16_row<DH>(o16, o, 1.f, ((ld16 & 7) | (((uintptr_t)out16) & 15)) == 0);
  }
}

template <int DH, int WARPS, bool VEC>
__global__ void attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                     const unsigned char* __restrict__ keep, float keep_scale,
                                     const float* __restrict__ dout, float* __restrict__ dqkv, __nv_bfloat16* __restrict__ dqkv16,
                                     long long ld16, int B, int S, int H) {
  __shared__ float Qs[WARPS][32][DH + 1];
  __shared__ float Ks[WARPS][32][DH + 1];
  __shared__ float Vs[WARPS][32][DH + 1];
  __shared__ float Os[WARPS][32][DH + 1];   // dOut rows
  __shared__ float dS[WARPS][32][33];
  __shared__ float Pd[WARPS][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x * WARPS + w;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * DH;
  const float scale = rsqrtf((float)DH);
  const long long pbase = ((long long)bh * S + lane) * S;
  float pr[32];
  unsigned kw[8];
  if (lane < S) {
    const float* row = qkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
    const float* drow = dout + ((long long)(b * S + lane)) * E + h * DH;
    float qr[DH], kr[DH], vr[DH], orr[DH];
    load_head_row<DH>(row, qr);
    load_head_row<DH>(row + E, kr);
    load_head_row<DH>(row + 2 * E, vr);
    load_head_row<DH>(drow, orr);
    // this lane's row of saved probabilities and keep bytes: in flight while the tiles are staged
    if (VEC) {
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(probs + pbase) + qd);
        pr[4 * qd] = x.x; pr[4 * qd + 1] = x.y; pr[4 * qd + 2] = x.z; pr[4 * qd + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) pr[j] = (j < S) ? __ldg(probs + pbase + j) : 0.f;
    }
    load_keep_row<VEC>(keep, pbase, S, kw);
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      Qs[w][lane][d] = qr[d];
      Ks[w][lane][d] = kr[d];
      Vs[w][lane][d] = vr[d];
      Os[w][lane][d] = orr[d];
    }
  }
  __syncwarp();
  if (lane < S) {
    float dp[32];
    float delta = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float g = 0.f;
      if (j < S) {
        const float p = pr[j];
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(Os[w][lane][d], Vs[w][j][d], a);
        float ks = keep ? (kept(kw, j) ? keep_scale : 0.f) : 1.f;
        g = a * ks;                 // dL/dp_ij
        Pd[w][lane][j] = p * ks;    // dropped probabilities
        delta = fmaf(p, g, delta);
      }
      dp[j] = g;
    }
    float dq[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < S) {
        float ds = pr[j] * (dp[j] - delta);
        dS[w][lane][j] = ds;
#pragma unroll
        for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, Ks[w][j][d], dq[d]);
      }
    }
    float* dqrow = dqkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4)
      *reinterpret_cast<float4*>(dqrow + d) = make_float4(dq[d] * scale, dq[d + 1] * scale, dq[d + 2] * scale, dq[d + 3] * scale);
    if (dqkv16) {   // operand mirror of the in-projection's backward contractions
      __nv_bfloat16* q16 = dqkv16 + ((long long)(b * S + lane)) * ld16 + h * DH;
      store_bf16_row<DH>(q16, dq, scale, ((ld16 & 7) | (((uintptr_t)dqkv16) & 15)) == 0);
    }
  }
  __syncwarp();
  if (lane < S) {
    // lane now plays key/value row j
    float dk[DH], dv[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) dk[d] = dv[d] = 0.f;
    for (int i = 0; i < S; ++i) {
      float ds = dS[w][i][lane], pd = Pd[w][i][lane];
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        dk[d] = fmaf(ds, Qs[w][i][d], dk[d]);
        dv[d] = fmaf(pd, Os[w][i][d], dv[d]);
      }
    }
    float* drow = dqkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      *reinterpret_cast<float4*>(drow + E + d) = make_float4(dk[d] * scale, dk[d + 1] * scale, dk[d + 2] * scale, dk[d + 3] * scale);
      *reinterpret_cast<float4*>(drow + 2 * E + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
    }
    if (dqkv16) {
      __nv_bfloat16* r16 = dqkv16 + ((long long)(b * S + lane)) * ld16 + h * DH;
      const bool al = ((ld16 & 7) | (((uintptr_t)dqkv16) & 15)) == 0;
      store_bf16_row<DH>(r16 + E, dk, scale, al);
      store_bf16_row<DH>(r16 + 2 * E, dv, 1.f, al);
    }
  }
}

__global__ void mean_seq_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S, int E) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  int b = i / E, e = i % E;
  float s = 0.f;
  for (int t = 0; t < S; ++t) s += x[((long long)b * S + t) * E + e];
  out[i] = s / (float)S;
}
__global__ void mean_seq_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, long long total, int S, int E) {
  float inv = 1.f / (float)S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / ((long long)S * E);
    int e = (int)(i % E);
    dx[i] = dout[b * E + e] * inv;
  }
}

inline int grid_for(long long n, int block) {
  long long want = (n + block - 1) / block;
  long long cap = 148LL * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

extern "C" {

int hulc2_add_pos_fwd_m(const float* emb, const float* pos, const unsigned char* keep, float keep_scale, float* out, void* out16,
                        int B, int S, int E, cudaStream_t st) {
  long long total = (long long)B * S * E;
  if (total <= 0) return HULC2_OK;
  add_pos_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(emb, pos, keep, keep_scale, out, (__nv_bfloat16*)out16, total, S * E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_add_pos_fwd(const float* emb, const float* pos, const unsigned char* keep, float keep_scale, float* out, int B,
                      int S, int E, cudaStream_t st) {
  return hulc2_add_pos_fwd_m(emb, pos, keep, keep_scale, out, nullptr, B, S, E, st);
}
int hulc2_add_pos_bwd(const float* dout, const unsigned char* keep, float keep_scale, float* demb, float* dpos, int B, int S,
                      int E, cudaStream_t st) {
  if ((long long)B * S * E <= 0) return HULC2_OK;
  add_pos_bwd_kernel<<<hulc2_cdiv(S * E, 32), dim3(32, 8), 0, st>>>(dout, keep, keep_scale, demb, dpos, B, S * E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_attention_fwd_m(const float* qkv, const unsigned char* keep, float keep_scale, float* out, float* probs, void* out16_,
                          long long ld16, int B, int S, int H, int Dh, cudaStream_t st) {
  __nv_bfloat16* out16 = (__nv_bfloat16*)out16_;
  if (B <= 0) return HULC2_OK;
  if (S > 32 || S <= 0) { hulc2_set_error("attention: window length must be in [1,32]"); return HULC2_EINVAL; }
  if (((uintptr_t)qkv | (uintptr_t)out) & 15) { hulc2_set_error("attention: qkv / out must be 16-byte aligned"); return HULC2_EINVAL; }
  // head_dim = latent / num_heads: 16 (RGB static + gripper, 128/8), 24 (+ depth_static, 192/8), 32 (RGBD_both, 256/8)
  const bool vec = S == 32 && (((uintptr_t)probs | (uintptr_t)keep) & 15) == 0;
  int blocks = hulc2_cdiv(B * H, 4);
#define HULC2_ATT_FWD(DH)                                                                                                        \
  do {                                                                                                                           \
    if (vec) attention_fwd_kernel<DH, 4, true><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, out16, ld16, B, S, H);  \
    else attention_fwd_kernel<DH, 4, false><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, out16, ld16, B, S, H);     \
  } while (0)
  if (Dh == 16) HULC2_ATT_FWD(16);
  else if (Dh == 8) HULC2_ATT_FWD(8);
  else if (Dh == 24) HULC2_ATT_FWD(24);
  else if (Dh == 32) HULC2_ATT_FWD(32);
  else { hulc2_set_error("attention: head_dim must be 8, 16, 24 or 32"); return HULC2_EINVAL; }
#undef HULC2_ATT_FWD
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_attention_fwd(const float* qkv, const unsigned char* keep, float keep_scale, float* out, float* probs, int B, int S,
                        int H, int Dh, cudaStream_t st) {
  return hulc2_attention_fwd_m(qkv, keep, keep_scale, out, probs, nullptr, 0, B, S, H, Dh, st);
}
int hulc2_attention_bwd_m(const float* qkv, const float* probs, const unsigned char* keep, float keep_scale, const float* dout,
                          float* dqkv, void* dqkv16_, long long ld16, int B, int S, int H, int Dh, cudaStream_t st) {
  __nv_bfloat16* dqkv16 = (__nv_bfloat16*)dqkv16_;
  if (B <= 0) return HULC2_OK;
  if (S > 32 || S <= 0) { hulc2_set_error("attention: window length must be in [1,32]"); return HULC2_EINVAL; }
  if (((uintptr_t)qkv | (uintptr_t)dout | (uintptr_t)dqkv) & 15) { hulc2_set_error("attention: qkv / dout / dqkv must be 16-byte aligned"); return HULC2_EINVAL; }
  const bool vec = S == 32 && (((uintptr_t)probs | (uintptr_t)keep) & 15) == 0;
#define HULC2_ATT_BWD(DH, W)                                                                                                                  \
  do {                                                                                                                                        \
    const int blocks = hulc2_cdiv(B * H, W);                                                                                                  \
    if (vec) attention_bwd_kernel<DH, W, true><<<blocks, 32 * W, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, dqkv16, ld16, B, S, H);    \
    else attention_bwd_kernel<DH, W, false><<<blocks, 32 * W, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, dqkv16, ld16, B, S, H);       \
  } while (0)
  if (Dh == 16) HULC2_ATT_BWD(16, 2);
  else if (Dh == 8) HULC2_ATT_BWD(8, 2);
  else if (Dh == 24) HULC2_ATT_BWD(24, 2);
  else if (Dh == 32) HULC2_ATT_BWD(32, 1);      // 1 warp: 48 KB static shared-memory limit
  else { hulc2_set_error("attention: head_dim must be 8, 16, 24 or 32"); return HULC2_EINVAL; }
#undef HULC2_ATT_BWD
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_attention_bwd(const float* qkv, const float* probs, const unsigned char* keep, float keep_scale, const float* dout,
                        float* dqkv, int B, int S, int H, int Dh, cudaStream_t st) {
  return hulc2_attention_bwd_m(qkv, probs, keep, keep_scale, dout, dqkv, nullptr, 0, B, S, H, Dh, st);
}
int hulc2_mean_seq_fwd(const float* x, float* out, int B, int S, int E, cudaStream_t st) {
  if ((long long)B * E <= 0) return HULC2_OK;
  mean_seq_fwd_kernel<<<hulc2_cdiv((long long)B * E, 256), 256, 0, st>>>(x, out, B, S, E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_mean_seq_bwd(const float* dout, float* dx, int B, int S, int E, cudaStream_t st) {
  long long total = (long long)B * S * E;
  if (total <= 0) return HULC2_OK;
  mean_seq_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(dout, dx, total, S, E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
