// Logistic-mixture action head (hulc2/models/decoders/logistic_decoder_rnn.py:133-152,181-255) and the
// tcp<->world frame transforms (hulc2/models/decoders/utils/gripper_control.py:16-63).
// HBM-bound elementwise/reduction kernels: one thread per (row, action-dim) reads its 3*M mixture
// parameters once, everything else stays in registers; the scalar losses are reduced deterministically.
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

constexpr int MAXM = 16;

struct HeadGeom {
  int B, S, A, M;
  long long ld;
  int time_major;
  int Bt;   // batch size of the heads buffer the rows live in (>= B: a segment [b0, b0+B) of a wider time-major buffer)
};
__device__ __forceinline__ long long head_row(const HeadGeom& g, int b, int s) {
  return g.time_major ? (long long)s * g.Bt + b : (long long)b * g.S + s;
}
__device__ __forceinline__ void row_to_bs(const HeadGeom& g, long long row, int& b, int& s) {
  if (g.time_major) { s = (int)(row / g.B); b = (int)(row - (long long)s * g.B); }
  else { b = (int)(row / g.S); s = (int)(row - (long long)b * g.S); }
}

// log-prob of the discretised logistic for one mixture component + derivatives wrt mean and (clamped) log-scale
struct Comp { float lp, dmu, dls; };
template <bool GRAD>
__device__ __forceinline__ Comp logistic_component(float act, float mu, float ls, float amin, float amax, int num_classes,
                                                   float log_half_classes) {
  Comp c; c.dmu = c.dls = 0.f;
  float centered = act - mu;
  float inv = expf(-ls);
  float half = ((amax - amin) / 2.0f) / (float)(num_classes - 1);
  float plus_in = inv * (centered + half);
  float min_in = inv * (centered - half);
  float mid_in = inv * centered;
  float cdf_plus = sigmoid_t(plus_in), cdf_min = sigmoid_t(min_in);
  float cdf_delta = cdf_plus - cdf_min;
  if (act < amin + 1e-3f) {
    c.lp = plus_in - softplus_t(plus_in);
    if (GRAD) { float d = 1.f - cdf_plus; c.dmu = -inv * d; c.dls = -plus_in * d; }
  } else if (act > amax - 1e-3f) {
    c.lp = -softplus_t(min_in);
    if (GRAD) { float d = -cdf_min; c.dmu = -inv * d; c.dls = -min_in * d; }
  } else if (cdf_delta > 1e-5f) {
    c.lp = logf(fmaxf(cdf_delta, 1e-12f));
    if (GRAD) {
      float sp = cdf_plus * (1.f - cdf_plus), sm = cdf_min * (1.f - cdf_min);
      c.dmu = -inv * (sp - sm) / cdf_delta;
      c.dls = -(plus_in * sp - min_in * sm) / cdf_delta;
    }
  } else {
    c.lp = (mid_in - ls - 2.0f * softplus_t(mid_in)) - log_half_classes;
    if (GRAD) { float d = 1.f - 2.f * sigmoid_t(mid_in); c.dmu = -inv * d; c.dls = -mid_in * d - 1.f; }
  }
  return c;
}

// MODE 0: forward partial sums; MODE 1: backward
template <int MODE>
__global__ void logistic_loss_kernel(const float* __restrict__ heads, HeadGeom g, const float* __restrict__ actions,
                                     const float* __restrict__ act_min, const float* __restrict__ act_max,
                                     int num_classes, float ls_min, float gripper_alpha, float log_half_classes,
                                     float* __restrict__ partial, const float* __restrict__ gout, float* __restrict__ dheads) {
  __shared__ float red[32];
  const long long rows = (long long)g.B * g.S;
  const long long total = rows * g.A;
  const int AM = g.A * g.M;
  float sum_ll = 0.f, sum_ce = 0.f;
  const float gscale = (MODE == 1) ? (gout ? gout[0] : 1.f) / (float)rows : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / g.A;
    int a = (int)(i - row * g.A);
    int b, s;
    row_to_bs(g, row, b, s);
    const long long hrow = head_row(g, b, s);
    const float* h = heads + hrow * g.ld;
    const float* arow = actions + ((long long)b * g.S + s) * (g.A + 1);
    float act = arow[a];
    float lpv[MAXM], dmu[MAXM], dls[MAXM];
    float lmax = -INFINITY;
    for (int m = 0; m < g.M; ++m) lmax = fmaxf(lmax, h[a * g.M + m]);
    float lsum = 0.f;
    for (int m = 0; m < g.M; ++m) lsum += expf(h[a * g.M + m] - lmax);
    float lnorm = lmax + logf(lsum);
    float mx = -INFINITY;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      lpv[m] = -INFINITY; dmu[m] = dls[m] = 0.f;
      if (m < g.M) {
        float raw = h[2 * AM + a * g.M + m];
        float ls = fmaxf(raw, ls_min);
        Comp c = logistic_component<MODE == 1>(act, h[AM + a * g.M + m], ls, act_min[a * g.M + m], act_max[a * g.M + m], num_classes,
                                               log_half_classes);
        lpv[m] = c.lp + (h[a * g.M + m] - lnorm);
        dmu[m] = c.dmu;
        dls[m] = (raw >= ls_min) ? c.dls : 0.f;
        mx = fmaxf(mx, lpv[m]);
      }
    }
    float se = 0.f;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) if (m < g.M) se += expf(lpv[m] - mx);
    float lse = mx + logf(se);
    if (MODE == 0) {
      sum_ll += -lse;
    } else {
      float* d = dheads + hrow * g.ld;
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < g.M) {
          float w = expf(lpv[m] - lse);
          float pi = expf(h[a * g.M + m] - lnorm);
          d[a * g.M + m] = -gscale * (w - pi);
          d[AM + a * g.M + m] = -gscale * w * dmu[m];
          d[2 * AM + a * g.M + m] = -gscale * w * dls[m];
        }
      }
    }
    if (a == 0) {  // gripper cross-entropy for this row (logistic_decoder_rnn.py:141-147)
      float gt = arow[g.A];
      int label = (gt == -1.f) ? 0 : (int)gt;
      float g0 = h[3 * AM], g1 = h[3 * AM + 1];
      float gm = fmaxf(g0, g1);
      float lz = gm + logf(expf(g0 - gm) + expf(g1 - gm));
      if (MODE == 0) {
        sum_ce += lz - (label == 0 ? g0 : g1);
      } else {
        float* d = dheads + hrow * g.ld;
        float p0 = expf(g0 - lz), p1 = expf(g1 - lz);
        d[3 * AM] = gscale * gripper_alpha * (p0 - (label == 0 ? 1.f : 0.f));
        d[3 * AM + 1] = gscale * gripper_alpha * (p1 - (label == 1 ? 1.f : 0.f));
      }
    }
  }
  if (MODE == 0) {
    float a = block_sum(sum_ll, red);
    float c = block_sum(sum_ce, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = c; }
  }
}

__global__ void logistic_loss_final_kernel(const float* __restrict__ partial, int nblocks, float inv_rows, float gripper_alpha,
                                           float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += partial[2 * i]; c += partial[2 * i + 1]; }
  a = block_sum(a, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    float ll = a * inv_rows, ce = c * inv_rows;
    out[0] = ll + gripper_alpha * ce;
    out[1] = ll;
    out[2] = ce;
  }
}

// Sampling (logistic_decoder_rnn.py:231-255).  Separate rounded mul/add (no FMA contraction) so the
// Gumbel argmax follows the reference's op sequence; logs are taken in double and rounded once.
__global__ void logistic_sample_kernel(const float* __restrict__ heads, HeadGeom g, const float* __restrict__ u1,
                                       const float* __restrict__ u2, const float* __restrict__ gripper_bounds,
                                       float* __restrict__ out, float ls_min) {
  const long long rows = (long long)g.B * g.S;
  const long long total = rows * g.A;
  const int AM = g.A * g.M;
  const float r1 = 1e-5f, r2 = (float)(1.0 - 1e-5);
  const float r12 = (float)(1e-5 - (1.0 - 1e-5));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / g.A;
    int a = (int)(i - row * g.A);
    int b, s;
    row_to_bs(g, row, b, s);
    const float* h = heads + row * g.ld;
    long long bs = (long long)b * g.S + s;
    const float* uu = u1 + (bs * g.A + a) * g.M;
    int best = 0;
    float bestv = -INFINITY;
    for (int m = 0; m < g.M; ++m) {
      float t = __fadd_rn(__fmul_rn(r12, uu[m]), r2);
      float gum = (float)log(-(double)(float)log((double)t));
      float v = __fsub_rn(h[a * g.M + m], gum);
      if (v > bestv) { bestv = v; best = m; }
    }
    float ls = fmaxf(h[2 * AM + a * g.M + best], ls_min);
    float mu = h[AM + a * g.M + best];
    float u = __fadd_rn(__fmul_rn(r12, u2[bs * g.A + a]), r2);
    float sc = (float)exp((double)ls);
    float lg = __fsub_rn((float)log((double)u), (float)log((double)__fsub_rn(1.0f, u)));
    out[bs * (g.A + 1) + a] = __fadd_rn(mu, __fmul_rn(sc, lg));
    if (a == 0) {
      float g0 = h[3 * AM], g1 = h[3 * AM + 1];
      out[bs * (g.A + 1) + g.A] = gripper_bounds[(g1 > g0) ? 1 : 0];
    }
    (void)r1;
  }
}

__global__ void heads_unpack_kernel(const float* __restrict__ heads, HeadGeom g, float* __restrict__ lp, float* __restrict__ ls,
                                    float* __restrict__ mu, float* __restrict__ grip, float ls_min) {
  const long long rows = (long long)g.B * g.S;
  const int AM = g.A * g.M;
  const long long total = rows * (AM + 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / (AM + 2);
    int c = (int)(i - row * (AM + 2));
    int b, s;
    row_to_bs(g, row, b, s);
    long long bs = (long long)b * g.S + s;
    const float* h = heads + row * g.ld;
    if (c < AM) {
      lp[bs * AM + c] = h[c];
      mu[bs * AM + c] = h[AM + c];
      ls[bs * AM + c] = fmaxf(h[2 * AM + c], ls_min);
    } else if (grip) {
      grip[bs * 2 + (c - AM)] = h[3 * AM + (c - AM)];
    }
  }
}

// --------------------------------------------------------------------------- frames (double internally)
struct M3 { double m[3][3]; };
__device__ __forceinline__ M3 euler_xyz(double a, double b, double c) {
  // pytorch3d euler_angles_to_matrix(.., "XYZ") = Rx(a) Ry(b) Rz(c)
  double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b), cc = cos(c), sc = sin(c);
  M3 r;
  r.m[0][0] = cb * cc;                 r.m[0][1] = -cb * sc;                r.m[0][2] = sb;
  r.m[1][0] = sa * sb * cc + ca * sc;  r.m[1][1] = -sa * sb * sc + ca * cc; r.m[1][2] = -sa * cb;
  r.m[2][0] = -ca * sb * cc + sa * sc; r.m[2][1] = ca * sb * sc + sa * cc;  r.m[2][2] = ca * cb;
  return r;
}
__device__ __forceinline__ void matrix_to_euler_xyz(const M3& r, double* e) {
  double s = fmin(1.0, fmax(-1.0, r.m[0][2]));
  e[0] = atan2(-r.m[1][2], r.m[2][2]);
  e[1] = asin(s);
  e[2] = atan2(-r.m[0][1], r.m[0][0]);
}
__device__ __forceinline__ float wrap100(float v) {
  // torch.where(x < -pi, x + 2pi, x); torch.where(x > pi, x - 2pi, x); x *= 100   (fp32, scalars cast to fp32)
  const float pi = (float)3.14159265358979323846, two_pi = (float)(2.0 * 3.14159265358979323846);
  if (v < -pi) v = v + two_pi;
  if (v > pi) v = v - two_pi;
  return v * 100.f;
}

template <bool TO_TCP>
__global__ void frame_kernel(const float* __restrict__ action, const float* __restrict__ robot, int robot_dim,
                             float* __restrict__ out, long long rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* a = action + r * 7;
    const float* o = robot + r * robot_dim;
    // the reference adds in fp32 before building the rotation: keep that rounding
    float e0 = o[3], e1 = o[4], e2 = o[5];
    M3 R = euler_xyz((double)e0, (double)e1, (double)e2);
    double p[3] = {(double)a[0], (double)a[1], (double)a[2]};
    float* y = out + r * 7;
    if (TO_TCP) {
      for (int i = 0; i < 3; ++i) y[i] = (float)(R.m[0][i] * p[0] + R.m[1][i] * p[1] + R.m[2][i] * p[2]);  // R^T p
      float n0 = e0 + a[3] * 0.01f, n1 = e1 + a[4] * 0.01f, n2 = e2 + a[5] * 0.01f;
      M3 N = euler_xyz((double)n0, (double)n1, (double)n2);
      M3 rel;  // N^T R
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rel.m[i][j] = N.m[0][i] * R.m[0][j] + N.m[1][i] * R.m[1][j] + N.m[2][i] * R.m[2][j];
      double e[3];
      matrix_to_euler_xyz(rel, e);
      for (int i = 0; i < 3; ++i) y[3 + i] = wrap100((float)e[i]);
    } else {
      for (int i = 0; i < 3; ++i) y[i] = (float)(R.m[i][0] * p[0] + R.m[i][1] * p[1] + R.m[i][2] * p[2]);  // R p
      float t0 = a[3] * 0.01f, t1 = a[4] * 0.01f, t2 = a[5] * 0.01f;
      M3 T = euler_xyz((double)t0, (double)t1, (double)t2);
      M3 W;  // R T^T
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) W.m[i][j] = R.m[i][0] * T.m[j][0] + R.m[i][1] * T.m[j][1] + R.m[i][2] * T.m[j][2];
      double e[3];
      matrix_to_euler_xyz(W, e);
      float ef[3] = {e0, e1, e2};
      for (int i = 0; i < 3; ++i) y[3 + i] = wrap100((float)e[i] - ef[i]);
    }
    y[6] = a[6];
  }
}

// --------------------------------------------------------------------------- validation metrics (hulc2.py:292-302, 559-575)
// One CTA per call (a modality's B windows): thread per (b, dim) takes the mean over the window like torch.mean(l1, 1), then
// block sums give total / position / orientation MAE (means over [B,6] / [B,3] / [B,3]) and the discrete-gripper success rate
// mean(gt == (pred > 0 ? 1 : -1)) over [B,S].  out[4] = {total_mae, pos_mae, orn_mae, grip_sr}.  pred / gt [B,S,A+1].
__global__ void val_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ out, int B, int S, int A) {
  __shared__ float red[32];
  const int A1 = A + 1;
  float tot = 0.f, pos = 0.f, orn = 0.f, hit = 0.f;
  for (int i = threadIdx.x; i < B * A; i += blockDim.x) {
    const int b = i / A, d = i - b * A;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
      const long long o = ((long long)b * S + s) * A1 + d;
      acc += fabsf(pred[o] - gt[o]);
    }
    acc /= (float)S;
    tot += acc;
    if (d < 3) pos += acc;
    else if (d < 6) orn += acc;
  }
  for (int i = threadIdx.x; i < B * S; i += blockDim.x) {
    const long long o = (long long)i * A1 + A;
    const float g = pred[o] > 0.f ? 1.f : -1.f;
    hit += (gt[o] == g) ? 1.f : 0.f;
  }
  tot = block_sum(tot, red);
  pos = block_sum(pos, red);
  orn = block_sum(orn, red);
  hit = block_sum(hit, red);
  if (threadIdx.x == 0) {
    out[0] = tot / (float)(B * A);
    out[1] = pos / (float)(B * 3);
    out[2] = orn / (float)(B * (A >= 6 ? 3 : (A > 3 ? A - 3 : 1)));
    out[3] = hit / (float)(B * S);
  }
}

inline int grid_for(long long n, int block) {
  long long want = (n + block - 1) / block;
  // up to 16 blocks of 128 threads per SM: these kernels are transcendental-heavy (ncu r02 at >= 256 MB: 45 % of the issue slots
  // busy with 4 blocks per SM, warps waiting on fixed-latency dependencies), so occupancy, not memory, is what they need
  long long cap = 148LL * 16;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

extern "C" {

int hulc2_logistic_loss_seg_fwd(const float* heads, long long ld, const float* actions, const float* act_min, const float* act_max,
                                float* out, int B, int S, int A, int M, int num_classes, float log_scale_min, float gripper_alpha,
                                int time_major, int B_total, void* workspace, long long workspace_bytes, cudaStream_t st) {
  if (M > MAXM || M <= 0) { hulc2_set_error("logistic_loss: n_mixtures must be in [1,16]"); return HULC2_EINVAL; }
  if (B_total < B) { hulc2_set_error("logistic_loss: B_total < B"); return HULC2_EINVAL; }
  if ((long long)B * S <= 0) return HULC2_OK;
  HeadGeom g{B, S, A, M, ld, time_major, B_total};
  int blocks = grid_for((long long)B * S * A, 128);
  if (!workspace || workspace_bytes < (long long)blocks * 2 * (long long)sizeof(float)) { hulc2_set_error("logistic_loss: workspace too small"); return HULC2_EWORKSPACE; }
  float lhc = (float)log((double)(num_classes - 1) / 2.0);
  logistic_loss_kernel<0><<<blocks, 128, 0, st>>>(heads, g, actions, act_min, act_max, num_classes, log_scale_min, gripper_alpha, lhc,
                                                  (float*)workspace, nullptr, nullptr);
  HULC2_CHECK_LAUNCH();
  logistic_loss_final_kernel<<<1, 256, 0, st>>>((const float*)workspace, blocks, 1.f / (float)((long long)B * S), gripper_alpha, out);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_logistic_loss_fwd(const float* heads, long long ld, const float* actions, const float* act_min, const float* act_max,
                            float* out, int B, int S, int A, int M, int num_classes, float log_scale_min, float gripper_alpha,
                            int time_major, void* workspace, long long workspace_bytes, cudaStream_t st) {
  return hulc2_logistic_loss_seg_fwd(heads, ld, actions, act_min, act_max, out, B, S, A, M, num_classes, log_scale_min, gripper_alpha,
                                     time_major, B, workspace, workspace_bytes, st);
}

int hulc2_logistic_loss_seg_bwd(const float* heads, long long ld, const float* actions, const float* act_min, const float* act_max,
                                const float* gout, float* dheads, int B, int S, int A, int M, int num_classes, float log_scale_min,
                                float gripper_alpha, int time_major, int B_total, cudaStream_t st) {
  if (M > MAXM || M <= 0) { hulc2_set_error("logistic_loss: n_mixtures must be in [1,16]"); return HULC2_EINVAL; }
  if (B_total < B) { hulc2_set_error("logistic_loss: B_total < B"); return HULC2_EINVAL; }
  if ((long long)B * S <= 0) return HULC2_OK;
  HeadGeom g{B, S, A, M, ld, time_major, B_total};
  int blocks = grid_for((long long)B * S * A, 128);
  float lhc = (float)log((double)(num_classes - 1) / 2.0);
  logistic_loss_kernel<1><<<blocks, 128, 0, st>>>(heads, g, actions, act_min, act_max, num_classes, log_scale_min, gripper_alpha, lhc,
                                                  nullptr, gout, dheads);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_logistic_loss_bwd(const float* heads, long long ld, const float* actions, const float* act_min, const float* act_max,
                            const float* gout, float* dheads, int B, int S, int A, int M, int num_classes, float log_scale_min,
                            float gripper_alpha, int time_major, cudaStream_t st) {
  return hulc2_logistic_loss_seg_bwd(heads, ld, actions, act_min, act_max, gout, dheads, B, S, A, M, num_classes, log_scale_min,
                                     gripper_alpha, time_major, B, st);
}

int hulc2_logistic_sample(const float* heads, long long ld, const float* u1, const float* u2, const float* gripper_bounds,
                          float* act, int B, int S, int A, int M, float log_scale_min, int time_major, cudaStream_t st) {
  if ((long long)B * S <= 0) return HULC2_OK;
  HeadGeom g{B, S, A, M, ld, time_major, B};
  logistic_sample_kernel<<<grid_for((long long)B * S * A, 128), 128, 0, st>>>(heads, g, u1, u2, gripper_bounds, act, log_scale_min);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_heads_unpack(const float* heads, long long ld, float* logit_probs, float* log_scales, float* means, float* gripper,
                       int B, int S, int A, int M, float log_scale_min, int time_major, cudaStream_t st) {
  if ((long long)B * S <= 0) return HULC2_OK;
  HeadGeom g{B, S, A, M, ld, time_major, B};
  heads_unpack_kernel<<<grid_for((long long)B * S * (A * M + 2), 256), 256, 0, st>>>(heads, g, logit_probs, log_scales, means, gripper, log_scale_min);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_val_metrics(const float* pred, const float* actions, float* out, int B, int S, int A, cudaStream_t st) {
  if (B <= 0 || S <= 0) return HULC2_OK;
  if (A < 3) { hulc2_set_error("val_metrics: needs at least 3 continuous action dims"); return HULC2_EINVAL; }
  val_metrics_kernel<<<1, 256, 0, st>>>(pred, actions, out, B, S, A);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_world_to_tcp(const float* action, const float* robot_obs, int robot_dim, float* out, long long rows, cudaStream_t st) {
  if (rows <= 0) return HULC2_OK;
  frame_kernel<true><<<grid_for(rows, 128), 128, 0, st>>>(action, robot_obs, robot_dim, out, rows);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_tcp_to_world(const float* action, const float* robot_obs, int robot_dim, float* out, long long rows, cudaStream_t st) {
  if (rows <= 0) return HULC2_OK;
  frame_kernel<false><<<grid_for(rows, 128), 128, 0, st>>>(action, robot_obs, robot_dim, out, rows);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
