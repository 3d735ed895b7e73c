// Plan-recognition transformer pieces that are not GEMMs (plan_recognition_net.py:125-148 and
// torch nn.TransformerEncoderLayer): positional add + dropout, 32x32-per-head attention held entirely
// in registers/shared memory (one warp per (window, head)), sequence mean.
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

__global__ void add_pos_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ pos,
                                   const unsigned char* __restrict__ keep, float keep_scale, float* __restrict__ out,
                                   long long total, int SE) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = emb[i] + pos[i % SE];
    if (keep) v = keep[i] ? v * keep_scale : 0.f;
    out[i] = v;
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

template <int DH, int WARPS>
__global__ void attention_fwd_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ keep, float keep_scale,
                                     float* __restrict__ out, float* __restrict__ probs, int B, int S, int H) {
  __shared__ float Ks[WARPS][32][DH + 1];
  __shared__ float Vs[WARPS][32][DH + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x * WARPS + w;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  const int E = H * DH;
  const float scale = rsqrtf((float)DH);
  float q[DH];
  if (lane < S) {
    const float* row = qkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      q[d] = row[d] * scale;
      Ks[w][lane][d] = row[E + d];
      Vs[w][lane][d] = row[2 * E + d];
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
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  const long long pbase = ((long long)bh * S + lane) * S;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < S) {
      float p = s[j] * inv;
      if (probs) probs[pbase + j] = p;
      if (keep) p = keep[pbase + j] ? p * keep_scale : 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, Vs[w][j][d], o[d]);
    }
  }
  float* orow = out + ((long long)(b * S + lane)) * E + h * DH;
#pragma unroll
  for (int d = 0; d < DH; ++d) orow[d] = o[d];
}

template <int DH, int WARPS>
__global__ void attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                     const unsigned char* __restrict__ keep, float keep_scale,
                                     const float* __restrict__ dout, float* __restrict__ dqkv, int B, int S, int H) {
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
  if (lane < S) {
    const float* row = qkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
    const float* drow = dout + ((long long)(b * S + lane)) * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      Qs[w][lane][d] = row[d];
      Ks[w][lane][d] = row[E + d];
      Vs[w][lane][d] = row[2 * E + d];
      Os[w][lane][d] = drow[d];
    }
  }
  __syncwarp();
  if (lane < S) {
    const long long pbase = ((long long)bh * S + lane) * S;
    float dp[32];
    float delta = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float p = 0.f, g = 0.f;
      if (j < S) {
        p = probs[pbase + j];
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) a = fmaf(Os[w][lane][d], Vs[w][j][d], a);
        float ks = keep ? (keep[pbase + j] ? keep_scale : 0.f) : 1.f;
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
        float p = probs[pbase + j];
        float ds = p * (dp[j] - delta);
        dS[w][lane][j] = ds;
#pragma unroll
        for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, Ks[w][j][d], dq[d]);
      }
    }
    float* dqrow = dqkv + ((long long)(b * S + lane)) * 3 * E + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) dqrow[d] = dq[d] * scale;
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
    for (int d = 0; d < DH; ++d) {
      drow[E + d] = dk[d] * scale;
      drow[2 * E + d] = dv[d];
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

int hulc2_add_pos_fwd(const float* emb, const float* pos, const unsigned char* keep, float keep_scale, float* out, int B,
                      int S, int E, cudaStream_t st) {
  long long total = (long long)B * S * E;
  if (total <= 0) return HULC2_OK;
  add_pos_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(emb, pos, keep, keep_scale, out, total, S * E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_add_pos_bwd(const float* dout, const unsigned char* keep, float keep_scale, float* demb, float* dpos, int B, int S,
                      int E, cudaStream_t st) {
  if ((long long)B * S * E <= 0) return HULC2_OK;
  add_pos_bwd_kernel<<<hulc2_cdiv(S * E, 32), dim3(32, 8), 0, st>>>(dout, keep, keep_scale, demb, dpos, B, S * E);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_attention_fwd(const float* qkv, const unsigned char* keep, float keep_scale, float* out, float* probs, int B, int S,
                        int H, int Dh, cudaStream_t st) {
  if (B <= 0) return HULC2_OK;
  if (S > 32 || S <= 0) { hulc2_set_error("attention: window length must be in [1,32]"); return HULC2_EINVAL; }
  // head_dim = latent / num_heads: 16 (RGB static + gripper, 128/8), 24 (+ depth_static, 192/8), 32 (RGBD_both, 256/8)
  int blocks = hulc2_cdiv(B * H, 4);
  if (Dh == 16) attention_fwd_kernel<16, 4><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, B, S, H);
  else if (Dh == 8) attention_fwd_kernel<8, 4><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, B, S, H);
  else if (Dh == 24) attention_fwd_kernel<24, 4><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, B, S, H);
  else if (Dh == 32) attention_fwd_kernel<32, 4><<<blocks, 128, 0, st>>>(qkv, keep, keep_scale, out, probs, B, S, H);
  else { hulc2_set_error("attention: head_dim must be 8, 16, 24 or 32"); return HULC2_EINVAL; }
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_attention_bwd(const float* qkv, const float* probs, const unsigned char* keep, float keep_scale, const float* dout,
                        float* dqkv, int B, int S, int H, int Dh, cudaStream_t st) {
  if (B <= 0) return HULC2_OK;
  if (S > 32 || S <= 0) { hulc2_set_error("attention: window length must be in [1,32]"); return HULC2_EINVAL; }
  int blocks = hulc2_cdiv(B * H, 2);
  if (Dh == 16) attention_bwd_kernel<16, 2><<<blocks, 64, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, B, S, H);
  else if (Dh == 8) attention_bwd_kernel<8, 2><<<blocks, 64, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, B, S, H);
  else if (Dh == 24) attention_bwd_kernel<24, 2><<<blocks, 64, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, B, S, H);
  else if (Dh == 32) attention_bwd_kernel<32, 1><<<B * H, 32, 0, st>>>(qkv, probs, keep, keep_scale, dout, dqkv, B, S, H);  // 1 warp: 48 KB static smem limit
  else { hulc2_set_error("attention: head_dim must be 8, 16, 24 or 32"); return HULC2_EINVAL; }
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
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
