// Visuo-lingual InfoNCE (hulc2/models/hulc2.py:472-508): row L2-normalisation (no eps), B x B similarity
// scaled by exp(logit_scale), symmetric cross-entropy with diagonal labels over the rows selected by
// use_for_aux_lang_loss.  The boolean row selection of the reference (data-dependent shapes + two host
// syncs) becomes a static-shape masked computation.  B x B x 32 is tiny (262 kFLOP at B=64): one CTA,
// deterministic, phases separated by __syncthreads, intermediates in a caller workspace.
#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

struct Ws {
  float *imn, *txn, *sim, *rlse, *clse, *inorm, *tnorm, *dsm;
  int P;   // row pitch of imn / txn = D + 1: the similarity phase reads txn[c * P + d] with c across lanes -- a pitch of D = 32
           // floats would put all 32 lanes on one shared-memory bank (measured: 67 of the kernel's 92 us)
};
// intermediates live in shared memory when they fit (B <= ~140 at D = 32: every shipped batch size), else in the workspace
__device__ __forceinline__ Ws carve(float* w, int B, int D) {
  Ws s;
  s.P = D + 1;
  s.imn = w; s.txn = s.imn + (long long)B * s.P; s.sim = s.txn + (long long)B * s.P;
  s.rlse = s.sim + (long long)B * B; s.clse = s.rlse + B; s.inorm = s.clse + B; s.tnorm = s.inorm + B;
  s.dsm = s.tnorm + B;
  return s;
}
extern __shared__ __align__(16) float infonce_smem[];

__device__ void infonce_forward_phases(const float* img, const float* txt, const unsigned char* use, float scale, int B, int D,
                                       const Ws& w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // row norms + normalised features: one warp per row, lanes over the feature dim (same sqrt / division as x / x.norm())
  for (int r = warp; r < B; r += nwarps) {
    float a = 0.f, b = 0.f;
    for (int d = lane; d < D; d += 32) { float x = img[(long long)r * D + d], y = txt[(long long)r * D + d]; a += x * x; b += y * y; }
    a = sqrtf(warp_sum(a)); b = sqrtf(warp_sum(b));
    if (lane == 0) { w.inorm[r] = a; w.tnorm[r] = b; }
    for (int d = lane; d < D; d += 32) { w.imn[(long long)r * w.P + d] = img[(long long)r * D + d] / a; w.txn[(long long)r * w.P + d] = txt[(long long)r * D + d] / b; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) {
    int r = i / B, c = i - r * B;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(scale * w.imn[(long long)r * w.P + d], w.txn[(long long)c * w.P + d], s);
    w.sim[i] = s;
  }
  __syncthreads();
  // row / column log-sum-exp over the selected columns: one warp per row, lanes over the columns
  for (int r = warp; r < B; r += nwarps) {
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int c = lane; c < B; c += 32)
      if (!use || use[c]) { m1 = fmaxf(m1, w.sim[(long long)r * B + c]); m2 = fmaxf(m2, w.sim[(long long)c * B + r]); }
    m1 = warp_max(m1); m2 = warp_max(m2);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < B; c += 32)
      if (!use || use[c]) { s1 += expf(w.sim[(long long)r * B + c] - m1); s2 += expf(w.sim[(long long)c * B + r] - m2); }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { w.rlse[r] = m1 + logf(s1); w.clse[r] = m2 + logf(s2); }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) infonce_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt,
                                   const unsigned char* __restrict__ use, const float* __restrict__ logit_scale,
                                   float* __restrict__ loss, int B, int D, float* __restrict__ wsp) {
  __shared__ float red[32];
  Ws w = carve(wsp ? wsp : infonce_smem, B, D);
  infonce_forward_phases(img, txt, use, expf(logit_scale[0]), B, D, w);
  float acc = 0.f, cnt = 0.f;
  for (int r = threadIdx.x; r < B; r += blockDim.x)
    if (!use || use[r]) { acc += (w.rlse[r] - w.sim[(long long)r * B + r]) + (w.clse[r] - w.sim[(long long)r * B + r]); cnt += 1.f; }
  acc = block_sum(acc, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) loss[0] = cnt > 0.f ? acc / (2.f * cnt) : 0.f;
}

__global__ void __launch_bounds__(512) infonce_bwd_kernel(const float* __restrict__ img, const float* __restrict__ txt,
                                   const unsigned char* __restrict__ use, const float* __restrict__ logit_scale,
                                   const float* __restrict__ gout, float* __restrict__ dimg, float* __restrict__ dtxt,
                                   float* __restrict__ dlogit_scale, int B, int D, float* __restrict__ wsp) {
  __shared__ float red[32];
  Ws w = carve(wsp ? wsp : infonce_smem, B, D);
  const float scale = expf(logit_scale[0]);
  infonce_forward_phases(img, txt, use, scale, B, D, w);
  float cnt = 0.f;
  for (int r = threadIdx.x; r < B; r += blockDim.x) if (!use || use[r]) cnt += 1.f;
  cnt = block_sum(cnt, red);
  const float g = cnt > 0.f ? (gout ? gout[0] : 1.f) / (2.f * cnt) : 0.f;
  // dsim(i,j) = g * [softmax_row_i(j) + softmax_col_j(i) - 2*delta_ij] on selected rows/cols
  auto dsim = [&](int i, int j) -> float {
    if (use && (!use[i] || !use[j])) return 0.f;
    float s = w.sim[(long long)i * B + j];
    float v = expf(s - w.rlse[i]) + expf(s - w.clse[j]);
    if (i == j) v -= 2.f;
    return g * v;
  };
  float dsc = 0.f;
  for (int i = threadIdx.x; i < B * B; i += blockDim.x) {
    float dv = dsim(i / B, i % B);
    w.dsm[i] = dv;
    if (dv != 0.f) dsc += dv * w.sim[i];
  }
  dsc = block_sum(dsc, red);            // (block_sum's barriers also publish dsm)
  if (threadIdx.x == 0 && dlogit_scale) dlogit_scale[0] += dsc;  // d/d(log scale) = sum dsim*sim
  // gradients wrt the normalised features, then through x/||x||
  for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
    int r = i / D, d = i - r * D;
    float a = 0.f, b = 0.f;
    for (int c = 0; c < B; ++c) {
      if (use && (!use[c] || !use[r])) continue;
      a = fmaf(w.dsm[(long long)r * B + c], w.txn[(long long)c * w.P + d], a);
      b = fmaf(w.dsm[(long long)c * B + r], w.imn[(long long)c * w.P + d], b);
    }
    dimg[i] = scale * a;   // temporarily d(imn)
    dtxt[i] = scale * b;   // temporarily d(txn)
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < B; r += nwarps) {                 // through x / ||x||: one warp per row
    if (use && !use[r]) {
      for (int d = lane; d < D; d += 32) dimg[(long long)r * D + d] = dtxt[(long long)r * D + d] = 0.f;
      continue;
    }
    float da = 0.f, db = 0.f;
    for (int d = lane; d < D; d += 32) { da += w.imn[(long long)r * w.P + d] * dimg[(long long)r * D + d]; db += w.txn[(long long)r * w.P + d] * dtxt[(long long)r * D + d]; }
    da = warp_sum(da); db = warp_sum(db);
    for (int d = lane; d < D; d += 32) {
      dimg[(long long)r * D + d] = (dimg[(long long)r * D + d] - w.imn[(long long)r * w.P + d] * da) / w.inorm[r];
      dtxt[(long long)r * D + d] = (dtxt[(long long)r * D + d] - w.txn[(long long)r * w.P + d] * db) / w.tnorm[r];
    }
  }
}

long long ws_floats(int B, int D) { return 2LL * B * (D + 1) + 2LL * B * B + 4LL * B; }
const long long kSmemMax = 200 * 1024;

}  // namespace

extern "C" {

int hulc2_infonce_fwd(const float* img, const float* txt, const unsigned char* use, const float* logit_scale, float* loss,
                      int B, int D, void* workspace, long long workspace_bytes, cudaStream_t st) {
  if (B <= 0) return HULC2_OK;
  const long long need = ws_floats(B, D) * (long long)sizeof(float);
  if (need <= kSmemMax) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(infonce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax); attr = true; }
    infonce_fwd_kernel<<<1, 512, need, st>>>(img, txt, use, logit_scale, loss, B, D, nullptr);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  if (!workspace || workspace_bytes < need) { hulc2_set_error("infonce: workspace too small"); return HULC2_EWORKSPACE; }
  infonce_fwd_kernel<<<1, 512, 0, st>>>(img, txt, use, logit_scale, loss, B, D, (float*)workspace);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_infonce_bwd(const float* img, const float* txt, const unsigned char* use, const float* logit_scale, const float* gout,
                      float* dimg, float* dtxt, float* dlogit_scale, int B, int D, void* workspace, long long workspace_bytes,
                      cudaStream_t st) {
  if (B <= 0) return HULC2_OK;
  const long long need = ws_floats(B, D) * (long long)sizeof(float);
  if (need <= kSmemMax) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(infonce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax); attr = true; }
    infonce_bwd_kernel<<<1, 512, need, st>>>(img, txt, use, logit_scale, gout, dimg, dtxt, dlogit_scale, B, D, nullptr);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  if (!workspace || workspace_bytes < need) { hulc2_set_error("infonce: workspace too small"); return HULC2_EWORKSPACE; }
  infonce_bwd_kernel<<<1, 512, 0, st>>>(img, txt, use, logit_scale, gout, dimg, dtxt, dlogit_scale, B, D, (float*)workspace);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
