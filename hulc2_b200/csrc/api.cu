// extern "C" entry points that dispatch between the fp32 CUDA-core path and the bf16 tcgen05 path,
// plus the host-side step loops of the decoder recurrence.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

int hulc2_gemm_f32_impl(const hulc2_gemm_args* a, cudaStream_t st);
int hulc2_conv2d_fwd_f32_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_conv2d_wgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_conv2d_dgrad_f32_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_gemm_bf16_impl(const hulc2_gemm_args* a, cudaStream_t st);
int hulc2_conv2d_fwd_bf16_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_conv2d_wgrad_bf16_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_conv2d_dgrad_bf16_impl(const hulc2_conv_args* a, cudaStream_t st);
int hulc2_rnn_persistent_launch(const float* add, const float* w, const float* init, const float* mask, float* out, float* final_out,
                                int S, int B, int H, int relu, int reverse, int transpose_w, void* workspace, long long workspace_bytes,
                                cudaStream_t st);

int hulc2_rnn_cluster_device_error(int clear);
int hulc2_rnn_cluster2_device_error(int clear);
int hulc2_rnn_cluster2_launch(const float* add, const float* w, const float* init, const float* mask, float* out, float* final_out,
                              int S, int B, int H, int relu, int reverse, int transpose_w, void* workspace, long long workspace_bytes,
                              void* states16, cudaStream_t st);
int hulc2_rnn_persistent_device_error(int clear);
int hulc2_rnn_cluster_launch(const float* add, const float* w, const float* init, const float* mask, float* out, float* final_out,
                             int S, int B, int H, int relu, int reverse, int transpose_w, void* workspace, long long workspace_bytes,
                             cudaStream_t st);

unsigned long long g_hulc2_launches = 0;
static thread_local char g_err[512] = "";
void hulc2_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

extern "C" {

const char* hulc2_last_error(void) { return g_err; }
int hulc2_version(void) { return 100; }
unsigned long long hulc2_launch_count(void) { return g_hulc2_launches; }

int hulc2_device_supports_tcgen05(void) {
  static int cached = -1;  // one device per process (one rank per GPU)
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    cached = (major == 10) ? 1 : 0;
  }
  return cached;
}

int hulc2_gemm(const hulc2_gemm_args* a, cudaStream_t st) {
  if (!a) { hulc2_set_error("gemm: null args"); return HULC2_EINVAL; }
  if (a->precision == 0 && a->rowsum) { hulc2_set_error("gemm: rowsum is served by the bf16 TMA path only (precision 1)"); return HULC2_EINVAL; }
  if (a->precision == 0) return hulc2_gemm_f32_impl(a, st);
  if (a->precision == 1) return hulc2_gemm_bf16_impl(a, st);
  hulc2_set_error("gemm: unknown precision");
  return HULC2_EINVAL;
}

static int check_conv(const hulc2_conv_args* a) {
  if (!a || a->F < 0 || a->C <= 0 || a->Cout <= 0 || a->KH <= 0 || a->KW <= 0 || a->stride <= 0 || a->H < a->KH || a->W < a->KW) {
    hulc2_set_error("conv2d: bad geometry");
    return HULC2_EINVAL;
  }
  return HULC2_OK;
}
int hulc2_conv2d_fwd(const hulc2_conv_args* a, cudaStream_t st) {
  if (int e = check_conv(a)) return e;
  if (a->precision == 0) return hulc2_conv2d_fwd_f32_impl(a, st);
  if (a->precision == 1) return hulc2_conv2d_fwd_bf16_impl(a, st);
  hulc2_set_error("conv2d_fwd: unknown precision");
  return HULC2_EINVAL;
}
int hulc2_conv2d_wgrad(const hulc2_conv_args* a, cudaStream_t st) {
  if (int e = check_conv(a)) return e;
  if (a->precision == 0) return hulc2_conv2d_wgrad_f32_impl(a, st);
  if (a->precision == 1) return hulc2_conv2d_wgrad_bf16_impl(a, st);
  hulc2_set_error("conv2d_wgrad: unknown precision");
  return HULC2_EINVAL;
}
int hulc2_conv2d_dgrad(const hulc2_conv_args* a, cudaStream_t st) {
  if (int e = check_conv(a)) return e;
  if (a->precision == 0) return hulc2_conv2d_dgrad_f32_impl(a, st);
  if (a->precision == 1) return hulc2_conv2d_dgrad_bf16_impl(a, st);
  hulc2_set_error("conv2d_dgrad: unknown precision");
  return HULC2_EINVAL;
}

static int g_rnn_path = 0;      // which kernel served the last hulc2_rnn_relu_{fwd,bwd}: 1 = (a'), 2 = (a), 3 = (b), 4 = per-step GEMMs
int g_rnn_v2_reject = 0;        // why (a') last declined: see rnn_cluster2_sm100.cu
int hulc2_rnn_last_path(void) { return g_rnn_path | (g_rnn_v2_reject << 8); }
static int g_rnn_kernel = 0;
static bool rnn_v2() {   // HULC2_RNN_V1=1: skip the TMA-fed cluster kernel (rnn_cluster2_sm100.cu) -- A/B switch, read once
  static int v = -1;
  if (v < 0) { const char* e = getenv("HULC2_RNN_V1"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
int hulc2_rnn_select_kernel(int which) {
  int prev = g_rnn_kernel;
  g_rnn_kernel = which;
  return prev;
}

int hulc2_rnn_device_error(int clear) {
  const int a = hulc2_rnn_cluster_device_error(clear), b = hulc2_rnn_persistent_device_error(clear);
  const int c = hulc2_rnn_cluster2_device_error(clear);
  if (a < 0 || b < 0 || c < 0) return -1;
  return a | (b << 1) | (c << 2);
}

// bf16 mirror of the states for the kernels that do not write it themselves: slot t + 1 of states16 = bf16(state t)
static int mirror_states(const float* states, void* states16, int S, int B, int H, int e, cudaStream_t st) {
  if (e != HULC2_OK || !states16) return e;
  const long long step = (long long)B * H;
  return hulc2_f32_to_bf16(states, reinterpret_cast<unsigned short*>(states16) + step, (long long)S * step, st);
}

// h[t] = relu(pre[t] + h[t-1] W_hh^T)      (nn.RNN, nonlinearity=relu; decoders/utils/rnn.py:5-14)
int hulc2_rnn_relu_fwd_m(const float* pre, const float* w_hh, const float* h0, float* h, void* h16, int S, int B, int H, int precision,
                         void* workspace, long long workspace_bytes, cudaStream_t st) {
  if (S <= 0 || B <= 0) return HULC2_OK;
  if (precision == 1 && hulc2_device_supports_tcgen05()) {
    int e = HULC2_ENOTIMPL;
    if (g_rnn_kernel == 0 && rnn_v2()) e = hulc2_rnn_cluster2_launch(pre, w_hh, h0, nullptr, h, nullptr, S, B, H, 1, 0, 0, workspace, workspace_bytes, h16, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 1; return e; }
    if (g_rnn_kernel < 1) e = hulc2_rnn_cluster_launch(pre, w_hh, h0, nullptr, h, nullptr, S, B, H, 1, 0, 0, workspace, workspace_bytes, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 2; return mirror_states(h, h16, S, B, H, e, st); }
    if (g_rnn_kernel < 2) e = hulc2_rnn_persistent_launch(pre, w_hh, h0, nullptr, h, nullptr, S, B, H, 1, 0, 0, workspace, workspace_bytes, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 3; return mirror_states(h, h16, S, B, H, e, st); }
  }
  g_rnn_path = 4;
  const long long step = (long long)B * H;
  for (int t = 0; t < S; ++t) {
    hulc2_gemm_args g;
    memset(&g, 0, sizeof(g));
    const float* prev = (t == 0) ? h0 : h + (t - 1) * step;
    g.M = B; g.N = H; g.K = prev ? H : 0;
    g.A = prev ? prev : w_hh; g.a_rs = H; g.a_ks = 1;
    g.B = w_hh; g.b_rs = H; g.b_ks = 1;
    g.C = h + t * step; g.ldc = H;
    g.add = pre + t * step; g.ld_add = H;
    g.relu = 1; g.alpha = 1.f; g.keep_scale = 1.f;
    g.precision = precision;
    if (int e = hulc2_gemm(&g, st)) return e;
  }
  return mirror_states(h, h16, S, B, H, HULC2_OK, st);
}
int hulc2_rnn_relu_fwd(const float* pre, const float* w_hh, const float* h0, float* h, int S, int B, int H, int precision,
                       void* workspace, long long workspace_bytes, cudaStream_t st) {
  return hulc2_rnn_relu_fwd_m(pre, w_hh, h0, h, nullptr, S, B, H, precision, workspace, workspace_bytes, st);
}

// in place: dh[t] <- dz[t] = (dh[t] + dz[t+1] W_hh) * (h[t] > 0); optional dh0 = dz[0] W_hh
int hulc2_rnn_relu_bwd_m(float* dh, const float* w_hh, const float* h, float* dh0, void* dz16, int S, int B, int H, int precision,
                         void* workspace, long long workspace_bytes, cudaStream_t st) {
  if (S <= 0 || B <= 0) return HULC2_OK;
  if (precision == 1 && hulc2_device_supports_tcgen05()) {
    int e = HULC2_ENOTIMPL;
    if (g_rnn_kernel == 0 && rnn_v2()) e = hulc2_rnn_cluster2_launch(dh, w_hh, nullptr, h, dh, dh0, S, B, H, 0, 1, 1, workspace, workspace_bytes, dz16, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 1; return e; }
    if (g_rnn_kernel < 1) e = hulc2_rnn_cluster_launch(dh, w_hh, nullptr, h, dh, dh0, S, B, H, 0, 1, 1, workspace, workspace_bytes, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 2; return mirror_states(dh, dz16, S, B, H, e, st); }
    if (g_rnn_kernel < 2) e = hulc2_rnn_persistent_launch(dh, w_hh, nullptr, h, dh, dh0, S, B, H, 0, 1, 1, workspace, workspace_bytes, st);
    if (e != HULC2_ENOTIMPL) { g_rnn_path = 3; return mirror_states(dh, dz16, S, B, H, e, st); }
  }
  g_rnn_path = 4;
  const long long step = (long long)B * H;
  if (int e = hulc2_relu_mask(dh + (S - 1) * step, h + (S - 1) * step, dh + (S - 1) * step, step, st)) return e;
  for (int t = S - 2; t >= -1; --t) {
    if (t < 0 && !dh0) break;
    hulc2_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = B; g.N = H; g.K = H;
    g.A = dh + (t + 1) * step; g.a_rs = H; g.a_ks = 1;
    g.B = w_hh; g.b_rs = 1; g.b_ks = H;                 // B(n=i, k=j) = W_hh[j, i]
    g.alpha = 1.f; g.keep_scale = 1.f; g.precision = precision;
    if (t >= 0) {
      g.C = dh + t * step; g.ldc = H;
      g.add = dh + t * step; g.ld_add = H;
      g.mask = h + t * step; g.ld_mask = H;
    } else {
      g.C = dh0; g.ldc = H;
    }
    if (int e = hulc2_gemm(&g, st)) return e;
  }
  return mirror_states(dh, dz16, S, B, H, HULC2_OK, st);
}
int hulc2_rnn_relu_bwd(float* dh, const float* w_hh, const float* h, float* dh0, int S, int B, int H, int precision,
                       void* workspace, long long workspace_bytes, cudaStream_t st) {
  return hulc2_rnn_relu_bwd_m(dh, w_hh, h, dh0, nullptr, S, B, H, precision, workspace, workspace_bytes, st);
}

}  // extern "C"
