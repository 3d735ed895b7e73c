// Shared device/host helpers for the hulc2_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HULC2_OK 0
#define HULC2_EINVAL (-1)
#define HULC2_ELAUNCH (-2)
#define HULC2_ENOTIMPL (-3)
#define HULC2_EWORKSPACE (-4)

extern unsigned long long g_hulc2_launches;
#define HULC2_CHECK_LAUNCH()                                         \
  do {                                                               \
    ++g_hulc2_launches;                                              \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) { hulc2_set_error(cudaGetErrorString(e__)); return HULC2_ELAUNCH; } \
  } while (0)

void hulc2_set_error(const char* msg);

static inline int hulc2_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in all threads; `red` needs >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}
// torch-compatible softplus (beta=1, threshold=20): x > 20 ? x : log1p(exp(x))
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_t(float x) { return 1.f / (1.f + expf(-x)); }
#endif
