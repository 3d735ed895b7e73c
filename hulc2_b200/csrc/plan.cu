// Latent-plan distribution kernels (hulc2/utils/distributions.py:15-60, hulc2/models/hulc2.py:444-466):
// balanced categorical KL forward/backward, one-hot scatter, straight-through gradient, inverse-CDF draw.
// One warp per (window, category); lane = class (classes <= 32).
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

struct LogSm { float lp, p; };
__device__ __forceinline__ LogSm log_softmax_lane(float z, bool valid) {
  float m = warp_max(valid ? z : -INFINITY);
  float e = valid ? expf(z - m) : 0.f;
  float s = warp_sum(e);
  LogSm r;
  r.lp = valid ? (z - m) - logf(s) : 0.f;
  r.p = valid ? e / s : 0.f;
  return r;
}
__device__ __forceinline__ float kl_term(float p, float lp, float q, float lq, bool valid) {
  // torch.distributions.kl._kl_categorical_categorical: t = p*(lp-lq); q==0 -> inf; p==0 -> 0
  if (!valid) return 0.f;
  float t = p * (lp - lq);
  if (q == 0.f) t = INFINITY;
  if (p == 0.f) t = 0.f;
  return t;
}

// One thread-block CLUSTER of 8 CTAs (8 SMs: the 2 x 32 expf + 32 x (expf, div, log) per row make this MUFU-throughput
// bound, 66 us on one SM), one THREAD per (b, cat) row (classes <= 32 values in registers).  Deterministic and without
// global scratch: every CTA writes its block sum into CTA 0's shared memory (distributed shared memory), CTA 0 adds the
// 8 partials in rank order.
constexpr int KL_CLUSTER = 8;
// sum over the rows of a grid-stride range of KL(softmax(pr_row) || softmax(pp_row)), one THREAD per row
__device__ __forceinline__ float kl_rows_sum(const float* __restrict__ pp, const float* __restrict__ pr, long long rows, int classes) {
  float acc = 0.f;
  const bool vec = (classes & 3) == 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float zp[32], zq[32];
    const float* a = pr + (long long)r * classes;
    const float* b = pp + (long long)r * classes;
    if (vec) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (4 * k < classes) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(a) + k), v = __ldg(reinterpret_cast<const float4*>(b) + k);
          zp[4 * k] = u.x; zp[4 * k + 1] = u.y; zp[4 * k + 2] = u.z; zp[4 * k + 3] = u.w;
          zq[4 * k] = v.x; zq[4 * k + 1] = v.y; zq[4 * k + 2] = v.z; zq[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < classes) { zp[k] = a[k]; zq[k] = b[k]; }
    }
    float mp = -INFINITY, mq = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < classes) { mp = fmaxf(mp, zp[k]); mq = fmaxf(mq, zq[k]); }
    float sp = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < classes) { zp[k] -= mp; zq[k] -= mq; sp += expf(zp[k]); sq += expf(zq[k]); }
    const float lsp = logf(sp), lsq = logf(sq);
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < classes) {
        const float lp = zp[k] - lsp, lq = zq[k] - lsq;
        t += kl_term(expf(zp[k]) / sp, lp, expf(zq[k]) / sq, lq, true);
      }
    acc += t;
  }
  return acc;
}

// Large problems (rows beyond what the 8-CTA cluster covers in a few iterations; the size sweep of profiles/): a plain grid
// over all SMs; every block adds its share of the (linear) loss with one atomic into the zeroed output.
__global__ void __launch_bounds__(256) kl_fwd_grid_kernel(const float* __restrict__ pp, const float* __restrict__ pr, float* __restrict__ loss,
                                                         long long rows, int classes, int B, float alpha, float beta) {
  __shared__ float red[32];
  const float tot = block_sum(kl_rows_sum(pp, pr, rows, classes), red);
  if (threadIdx.x == 0) {
    const float kl = tot / (float)B;
    atomicAdd(loss, (alpha * kl + (1.f - alpha) * kl) * beta);
  }
}

__global__ void __cluster_dims__(KL_CLUSTER, 1, 1) __launch_bounds__(512)
    kl_fwd_kernel(const float* __restrict__ pp, const float* __restrict__ pr, float* __restrict__ loss, int rows, int classes, int B,
                  float alpha, float beta) {
  __shared__ float red[32];
  __shared__ float part[KL_CLUSTER];
  const float acc = kl_rows_sum(pp, pr, rows, classes);
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    // part[rank] of CTA 0, written through the cluster shared-memory window
    uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[blockIdx.x]), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(tot) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < KL_CLUSTER; ++c) s += part[c];
    const float kl = s / (float)B;
    loss[0] = (alpha * kl + (1.f - alpha) * kl) * beta;
  }
}

__global__ void kl_bwd_kernel(const float* __restrict__ pp, const float* __restrict__ pr, const float* __restrict__ gout,
                              float* __restrict__ dpp, float* __restrict__ dpr, int rows, int classes, int B, float alpha,
                              float beta) {
  int lane = threadIdx.x & 31;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  bool valid = lane < classes;
  float g = (gout ? gout[0] : 1.f) * beta / (float)B;
  for (long long r = warp; r < rows; r += nwarps) {
    LogSm P = log_softmax_lane(valid ? pr[r * classes + lane] : 0.f, valid);
    LogSm Q = log_softmax_lane(valid ? pp[r * classes + lane] : 0.f, valid);
    float kl = warp_sum(kl_term(P.p, P.lp, Q.p, Q.lp, valid));
    if (valid) {
      if (dpp) dpp[r * classes + lane] = g * alpha * (Q.p - P.p);
      if (dpr) dpr[r * classes + lane] = g * (1.f - alpha) * P.p * ((P.lp - Q.lp) - kl);
    }
  }
}

__global__ void onehot_kernel(const long long* __restrict__ idx, float* __restrict__ plan, long long rows, int classes) {
  long long total = rows * classes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / classes;
    int c = (int)(i - r * classes);
    plan[i] = (idx[r] == c) ? 1.f : 0.f;
  }
}

// rsample() = onehot + (p - sg(p)):  dlogits_k = p_k (g_k - sum_j p_j g_j)
__global__ void st_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dplan, float* __restrict__ dlogits,
                              long long rows, int classes) {
  int lane = threadIdx.x & 31;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  bool valid = lane < classes;
  for (long long r = warp; r < rows; r += nwarps) {
    LogSm P = log_softmax_lane(valid ? logits[r * classes + lane] : 0.f, valid);
    float g = valid ? dplan[r * classes + lane] : 0.f;
    float dot = warp_sum(P.p * g);
    if (valid) dlogits[r * classes + lane] = P.p * (g - dot);
  }
}

__global__ void categorical_sample_kernel(const float* __restrict__ logits, const float* __restrict__ u,
                                          long long* __restrict__ idx, long long rows, int classes) {
  int lane = threadIdx.x & 31;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  bool valid = lane < classes;
  for (long long r = warp; r < rows; r += nwarps) {
    LogSm P = log_softmax_lane(valid ? logits[r * classes + lane] : 0.f, valid);
    float c = P.p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, c, o);
      if (lane >= o) c += t;
    }
    unsigned ball = __ballot_sync(0xffffffffu, valid && c > u[r]);
    int k = ball ? (__ffs(ball) - 1) : (classes - 1);
    if (lane == 0) idx[r] = k;
  }
}

inline int warp_grid(long long rows) {
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int hulc2_kl_fwd(const float* pp, const float* pr, float* loss, int B, int cats, int classes, float alpha, float beta,
                 cudaStream_t st) {
  if (classes > 32 || classes <= 0) { hulc2_set_error("kl: class_size must be in [1,32]"); return HULC2_EINVAL; }
  const long long rows = (long long)B * cats;
  if (rows > 4LL * KL_CLUSTER * 512) {          // beyond 16 k rows the 8-SM cluster is the bottleneck: all SMs + one atomic per block
    if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) { hulc2_set_error("kl_fwd: memset failed"); return HULC2_ELAUNCH; }
    long long blocks = (rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    kl_fwd_grid_kernel<<<(int)blocks, 256, 0, st>>>(pp, pr, loss, rows, classes, B, alpha, beta);
  } else {
    kl_fwd_kernel<<<KL_CLUSTER, 512, 0, st>>>(pp, pr, loss, (int)rows, classes, B, alpha, beta);
  }
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_kl_bwd(const float* pp, const float* pr, const float* gout, float* dpp, float* dpr, int B, int cats, int classes,
                 float alpha, float beta, cudaStream_t st) {
  if (classes > 32 || classes <= 0) { hulc2_set_error("kl: class_size must be in [1,32]"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  kl_bwd_kernel<<<warp_grid((long long)B * cats), 256, 0, st>>>(pp, pr, gout, dpp, dpr, B * cats, classes, B, alpha, beta);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_onehot_fwd(const long long* idx, float* plan, int B, int cats, int classes, cudaStream_t st) {
  long long rows = (long long)B * cats;
  if (rows <= 0) return HULC2_OK;
  long long blocks = (rows * classes + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  onehot_kernel<<<(int)blocks, 256, 0, st>>>(idx, plan, rows, classes);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_st_onehot_bwd(const float* logits, const float* dplan, float* dlogits, int B, int cats, int classes, cudaStream_t st) {
  if (classes > 32 || classes <= 0) { hulc2_set_error("st_onehot: class_size must be in [1,32]"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  st_bwd_kernel<<<warp_grid((long long)B * cats), 256, 0, st>>>(logits, dplan, dlogits, (long long)B * cats, classes);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_categorical_sample(const float* logits, const float* u, long long* idx, int B, int cats, int classes, cudaStream_t st) {
  if (classes > 32 || classes <= 0) { hulc2_set_error("categorical_sample: class_size must be in [1,32]"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  categorical_sample_kernel<<<warp_grid((long long)B * cats), 256, 0, st>>>(logits, u, idx, (long long)B * cats, classes);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
