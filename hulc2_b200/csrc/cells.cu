// Gated recurrence cells of the action decoder (hulc2/models/decoders/utils/rnn.py:17-36: nn.LSTM / nn.GRU selected by
// `rnn_model`) and the continuous (diagonal Gaussian) latent plan (hulc2/utils/distributions.py:28-29,55-59,
// hulc2/models/hulc2.py:444-466).
//
// The recurrent contraction h_{t-1} W_hh^T of a step is a GEMM (hulc2_gemm); what is left per step is pure
// elementwise work over [B, H] with 3-4 gate streams -- HBM/L2-bound, one thread per 4 hidden units, float4 loads.
// Algorithmic bytes per (b, unit): GRU fwd 6 gate reads + h_prev + h + 4 saved = 48 B, bwd 4 saved + dh(2) + h_prev
// reads, 7 writes = 56 B; LSTM fwd 8 gate reads + c_prev, h + c + 4 saved = 60 B, bwd 52 B.
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

#define FOR4(expr)                     \
  {                                    \
    { const int q = 0; expr; }         \
    { const int q = 1; expr; }         \
    { const int q = 2; expr; }         \
    { const int q = 3; expr; }         \
  }
__device__ __forceinline__ float& at(float4& v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }
__device__ __forceinline__ float at(const float4& v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }

// ---------------------------------------------------------------- GRU (gate order r, z, n; torch.nn.GRU)
// r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h = (1 - z) n + z h_prev
// gi already holds W_ih x + b_ih, gh holds W_hh h_prev + b_hh.  save [B, 4H] = r | z | n | gh_n.
__global__ void gru_cell_fwd_kernel(const float* __restrict__ gi, long long ldgi, const float* __restrict__ gh,
                                    const float* __restrict__ hprev, float* __restrict__ h, float* __restrict__ save,
                                    int B, int H) {
  const int H4 = H >> 2;
  const long long total = (long long)B * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H4), j = (int)(i - (long long)b * H4) * 4;
    const float* gib = gi + (long long)b * ldgi;
    const float* ghb = gh + (long long)b * 3 * H;
    const float4 ir = ld4(gib + j), iz = ld4(gib + H + j), in = ld4(gib + 2 * H + j);
    const float4 hr = ld4(ghb + j), hz = ld4(ghb + H + j), hn = ld4(ghb + 2 * H + j);
    const float4 hp = hprev ? ld4(hprev + (long long)b * H + j) : zero4();
    float4 r, z, n, o;
    FOR4(at(r, q) = sigmoid_t(at(ir, q) + at(hr, q)); at(z, q) = sigmoid_t(at(iz, q) + at(hz, q));
         at(n, q) = tanhf(at(in, q) + at(r, q) * at(hn, q));
         at(o, q) = (1.f - at(z, q)) * at(n, q) + at(z, q) * at(hp, q));
    st4(h + (long long)b * H + j, o);
    if (save) {
      float* s = save + (long long)b * 4 * H;
      st4(s + j, r); st4(s + H + j, z); st4(s + 2 * H + j, n); st4(s + 3 * H + j, hn);
    }
  }
}

// dh = dh_a + dh_b (upstream gradient of h_t + recurrent gradient from step t+1, either may be null).
// dgi [B, 3H] (row stride lddgi) = d/d(gi), dgh [B, 3H] = d/d(gh), dh_prev [B, H] = direct path dh * z.
__global__ void gru_cell_bwd_kernel(const float* __restrict__ dha, const float* __restrict__ dhb, const float* __restrict__ save,
                                    const float* __restrict__ hprev, float* __restrict__ dgi, long long lddgi,
                                    float* __restrict__ dgh, float* __restrict__ dhprev, int B, int H) {
  const int H4 = H >> 2;
  const long long total = (long long)B * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H4), j = (int)(i - (long long)b * H4) * 4;
    const float* s = save + (long long)b * 4 * H;
    const float4 r = ld4(s + j), z = ld4(s + H + j), n = ld4(s + 2 * H + j), hn = ld4(s + 3 * H + j);
    const float4 hp = hprev ? ld4(hprev + (long long)b * H + j) : zero4();
    float4 dh = dha ? ld4(dha + (long long)b * H + j) : zero4();
    if (dhb) { const float4 t = ld4(dhb + (long long)b * H + j); dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w; }
    float4 dr, dz, dn, dnh, dp;
    FOR4(const float g = at(dh, q); const float dn_pre = g * (1.f - at(z, q)) * (1.f - at(n, q) * at(n, q));
         at(dn, q) = dn_pre; at(dz, q) = g * (at(hp, q) - at(n, q)) * at(z, q) * (1.f - at(z, q));
         at(dr, q) = dn_pre * at(hn, q) * at(r, q) * (1.f - at(r, q)); at(dnh, q) = dn_pre * at(r, q);
         at(dp, q) = g * at(z, q));
    float* a = dgi + (long long)b * lddgi;
    st4(a + j, dr); st4(a + H + j, dz); st4(a + 2 * H + j, dn);
    float* c = dgh + (long long)b * 3 * H;
    st4(c + j, dr); st4(c + H + j, dz); st4(c + 2 * H + j, dnh);
    st4(dhprev + (long long)b * H + j, dp);
  }
}

// ---------------------------------------------------------------- LSTM (gate order i, f, g, o; torch.nn.LSTM)
// c = s(f) c_prev + s(i) tanh(g), h = s(o) tanh(c).  save [B, 4H] = activated i | f | g | o.
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gi, long long ldgi, const float* __restrict__ gh,
                                     const float* __restrict__ cprev, float* __restrict__ h, float* __restrict__ c,
                                     float* __restrict__ save, int B, int H) {
  const int H4 = H >> 2;
  const long long total = (long long)B * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H4), j = (int)(i - (long long)b * H4) * 4;
    const float* gib = gi + (long long)b * ldgi;
    const float* ghb = gh + (long long)b * 4 * H;
    const float4 ai = ld4(gib + j), af = ld4(gib + H + j), ag = ld4(gib + 2 * H + j), ao = ld4(gib + 3 * H + j);
    const float4 bi = ld4(ghb + j), bf = ld4(ghb + H + j), bg = ld4(ghb + 2 * H + j), bo = ld4(ghb + 3 * H + j);
    const float4 cp = cprev ? ld4(cprev + (long long)b * H + j) : zero4();
    float4 I, Fg, G, O, cn, hn;
    FOR4(at(I, q) = sigmoid_t(at(ai, q) + at(bi, q)); at(Fg, q) = sigmoid_t(at(af, q) + at(bf, q));
         at(G, q) = tanhf(at(ag, q) + at(bg, q)); at(O, q) = sigmoid_t(at(ao, q) + at(bo, q));
         at(cn, q) = at(Fg, q) * at(cp, q) + at(I, q) * at(G, q); at(hn, q) = at(O, q) * tanhf(at(cn, q)));
    st4(h + (long long)b * H + j, hn);
    st4(c + (long long)b * H + j, cn);
    if (save) {
      float* s = save + (long long)b * 4 * H;
      st4(s + j, I); st4(s + H + j, Fg); st4(s + 2 * H + j, G); st4(s + 3 * H + j, O);
    }
  }
}

// dc [B, H] is read (gradient into c_t from step t+1, null = 0) and the gradient into c_{t-1} is written to dcprev
// (may alias dc).  dg [B, 4H] (row stride lddg) = gradient of the pre-activation gates (same for gi and gh).
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dha, const float* __restrict__ dhb, const float* dc,
                                     const float* __restrict__ save, const float* __restrict__ c, const float* __restrict__ cprev,
                                     float* __restrict__ dg, long long lddg, float* dcprev, int B, int H) {
  const int H4 = H >> 2;
  const long long total = (long long)B * H4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H4), j = (int)(i - (long long)b * H4) * 4;
    const float* s = save + (long long)b * 4 * H;
    const float4 I = ld4(s + j), Fg = ld4(s + H + j), G = ld4(s + 2 * H + j), O = ld4(s + 3 * H + j);
    const float4 ct = ld4(c + (long long)b * H + j);
    const float4 cp = cprev ? ld4(cprev + (long long)b * H + j) : zero4();
    float4 dh = dha ? ld4(dha + (long long)b * H + j) : zero4();
    if (dhb) { const float4 t = ld4(dhb + (long long)b * H + j); dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w; }
    const float4 dcin = dc ? ld4(dc + (long long)b * H + j) : zero4();
    float4 di, df, dgg, dO, dcp;
    FOR4(const float tc = tanhf(at(ct, q)); const float dct = at(dcin, q) + at(dh, q) * at(O, q) * (1.f - tc * tc);
         at(dO, q) = at(dh, q) * tc * at(O, q) * (1.f - at(O, q));
         at(di, q) = dct * at(G, q) * at(I, q) * (1.f - at(I, q));
         at(df, q) = dct * at(cp, q) * at(Fg, q) * (1.f - at(Fg, q));
         at(dgg, q) = dct * at(I, q) * (1.f - at(G, q) * at(G, q));
         at(dcp, q) = dct * at(Fg, q));
    float* a = dg + (long long)b * lddg;
    st4(a + j, di); st4(a + H + j, df); st4(a + 2 * H + j, dgg); st4(a + 3 * H + j, dO);
    st4(dcprev + (long long)b * H + j, dcp);
  }
}

// ---------------------------------------------------------------- continuous latent plan
// forward_dist (distributions.py:55-59): x [B, 2P] -> mean = x[:, :P], std = softplus(x[:, P:]) + 1e-4
__global__ void gauss_state_fwd_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ std, long long B, int P) {
  const long long total = B * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / P;
    const int k = (int)(i - b * P);
    mean[i] = x[b * 2 * P + k];
    std[i] = softplus_t(x[b * 2 * P + P + k]) + 1e-4f;
  }
}
// dx[:, :P] = dmean, dx[:, P:] = dstd * sigmoid(x[:, P:])  (softplus' ; torch uses 1 above the threshold 20)
__global__ void gauss_state_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dmean, const float* __restrict__ dstd,
                                       float* __restrict__ dx, long long B, int P) {
  const long long total = B * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / P;
    const int k = (int)(i - b * P);
    const float v = x[b * 2 * P + P + k];
    dx[b * 2 * P + k] = dmean ? dmean[i] : 0.f;
    dx[b * 2 * P + P + k] = dstd ? dstd[i] * (v > 20.f ? 1.f : sigmoid_t(v)) : 0.f;
  }
}
// Normal.rsample / sample: plan = mean + std * eps (mean null = 0: the std-gradient g * eps of rsample)
__global__ void gauss_rsample_kernel(const float* __restrict__ mean, const float* __restrict__ std, const float* __restrict__ eps,
                                     float* __restrict__ plan, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    plan[i] = (mean ? mean[i] : 0.f) + std[i] * eps[i];
}
// standard normals from two uniform streams in (0,1]: eps = sqrt(-2 ln u1) cos(2 pi u2)
__global__ void box_muller_kernel(const float* __restrict__ u1, const float* __restrict__ u2, float* __restrict__ eps, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    eps[i] = sqrtf(-2.f * logf(fmaxf(u1[i], 1e-30f))) * cospif(2.f * u2[i]);
}

// KL(N(mp, sp) || N(mq, sq)) per element (torch.distributions.kl._kl_normal_normal):
//   var_ratio = (sp/sq)^2, t1 = ((mp - mq)/sq)^2, kl = 0.5 (var_ratio + t1 - 1 - log var_ratio)
// hulc2.py:444-466: loss = beta (alpha KL(sg(pr)||pp) + (1-alpha) KL(pr||sg(pp))), p = pr (posterior), q = pp (prior);
// summed over the P plan dims (Independent(..., 1)), mean over B.  Single block, deterministic.
__global__ void __launch_bounds__(1024) gauss_kl_fwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq, const float* __restrict__ mp,
                                                            const float* __restrict__ sp, float* __restrict__ loss, long long n, int B,
                                                            float alpha, float beta) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float ratio = sp[i] / sq[i], vr = ratio * ratio, d = (mp[i] - mq[i]) / sq[i];
    acc += 0.5f * (vr + d * d - 1.f - logf(vr));
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float kl = tot / (float)B;
    loss[0] = (alpha * kl + (1.f - alpha) * kl) * beta;
  }
}
// prior (pp = q) receives alpha * d/dq, posterior (pr = p) receives (1 - alpha) * d/dp; all scaled by gout * beta / B
__global__ void gauss_kl_bwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq, const float* __restrict__ mp,
                                    const float* __restrict__ sp, const float* __restrict__ gout, float* __restrict__ dmq,
                                    float* __restrict__ dsq, float* __restrict__ dmp, float* __restrict__ dsp, long long n, int B,
                                    float alpha, float beta) {
  const float g = (gout ? gout[0] : 1.f) * beta / (float)B;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = sp[i], b = sq[i], d = mp[i] - mq[i];
    const float ib2 = 1.f / (b * b);
    if (dmp) dmp[i] = g * (1.f - alpha) * d * ib2;
    if (dsp) dsp[i] = g * (1.f - alpha) * (a * ib2 - 1.f / a);
    if (dmq) dmq[i] = -g * alpha * d * ib2;
    if (dsq) dsq[i] = g * alpha * (-(a * a + d * d) * ib2 / b + 1.f / b);
  }
}

inline int ew_grid(long long n, int threads = 256) {
  long long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int hulc2_gru_cell_fwd(const float* gi, long long ldgi, const float* gh, const float* h_prev, float* h, float* save, int B, int H,
                       cudaStream_t st) {
  if (H <= 0 || (H & 3) || (ldgi & 3)) { hulc2_set_error("gru_cell: hidden size and row stride must be multiples of 4"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  gru_cell_fwd_kernel<<<ew_grid((long long)B * (H >> 2)), 256, 0, st>>>(gi, ldgi, gh, h_prev, h, save, B, H);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_gru_cell_bwd(const float* dh_a, const float* dh_b, const float* save, const float* h_prev, float* dgi, long long lddgi,
                       float* dgh, float* dh_prev, int B, int H, cudaStream_t st) {
  if (H <= 0 || (H & 3) || (lddgi & 3)) { hulc2_set_error("gru_cell: hidden size and row stride must be multiples of 4"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  gru_cell_bwd_kernel<<<ew_grid((long long)B * (H >> 2)), 256, 0, st>>>(dh_a, dh_b, save, h_prev, dgi, lddgi, dgh, dh_prev, B, H);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_lstm_cell_fwd(const float* gi, long long ldgi, const float* gh, const float* c_prev, float* h, float* c, float* save,
                        int B, int H, cudaStream_t st) {
  if (H <= 0 || (H & 3) || (ldgi & 3)) { hulc2_set_error("lstm_cell: hidden size and row stride must be multiples of 4"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  lstm_cell_fwd_kernel<<<ew_grid((long long)B * (H >> 2)), 256, 0, st>>>(gi, ldgi, gh, c_prev, h, c, save, B, H);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_lstm_cell_bwd(const float* dh_a, const float* dh_b, const float* dc, const float* save, const float* c, const float* c_prev,
                        float* dgates, long long lddg, float* dc_prev, int B, int H, cudaStream_t st) {
  if (H <= 0 || (H & 3) || (lddg & 3)) { hulc2_set_error("lstm_cell: hidden size and row stride must be multiples of 4"); return HULC2_EINVAL; }
  if (B <= 0) return HULC2_OK;
  lstm_cell_bwd_kernel<<<ew_grid((long long)B * (H >> 2)), 256, 0, st>>>(dh_a, dh_b, dc, save, c, c_prev, dgates, lddg, dc_prev, B, H);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_gauss_state_fwd(const float* x, float* mean, float* std, int B, int P, cudaStream_t st) {
  if (B <= 0 || P <= 0) return HULC2_OK;
  gauss_state_fwd_kernel<<<ew_grid((long long)B * P), 256, 0, st>>>(x, mean, std, B, P);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_gauss_state_bwd(const float* x, const float* dmean, const float* dstd, float* dx, int B, int P, cudaStream_t st) {
  if (B <= 0 || P <= 0) return HULC2_OK;
  gauss_state_bwd_kernel<<<ew_grid((long long)B * P), 256, 0, st>>>(x, dmean, dstd, dx, B, P);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_gauss_rsample(const float* mean, const float* std, const float* eps, float* plan, long long n, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  gauss_rsample_kernel<<<ew_grid(n), 256, 0, st>>>(mean, std, eps, plan, n);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_box_muller(const float* u1, const float* u2, float* eps, long long n, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  box_muller_kernel<<<ew_grid(n), 256, 0, st>>>(u1, u2, eps, n);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_gauss_kl_fwd(const float* pp_mean, const float* pp_std, const float* pr_mean, const float* pr_std, float* loss, int B, int P,
                       float alpha, float beta, cudaStream_t st) {
  if (B <= 0 || P <= 0) { hulc2_set_error("gauss_kl: empty batch"); return HULC2_EINVAL; }
  gauss_kl_fwd_kernel<<<1, 1024, 0, st>>>(pp_mean, pp_std, pr_mean, pr_std, loss, (long long)B * P, B, alpha, beta);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_gauss_kl_bwd(const float* pp_mean, const float* pp_std, const float* pr_mean, const float* pr_std, const float* gout,
                       float* dpp_mean, float* dpp_std, float* dpr_mean, float* dpr_std, int B, int P, float alpha, float beta,
                       cudaStream_t st) {
  if (B <= 0 || P <= 0) return HULC2_OK;
  gauss_kl_bwd_kernel<<<ew_grid((long long)B * P), 256, 0, st>>>(pp_mean, pp_std, pr_mean, pr_std, gout, dpp_mean, dpp_std, dpr_mean,
                                                                  dpr_std, (long long)B * P, B, alpha, beta);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
