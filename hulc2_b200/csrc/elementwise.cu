// Memory-bound utility kernels: strided copies, column sums (bias gradients), layout shuffles,
// Philox noise, fused Adam.  All are grid-stride, coalesced, sized in multiples of the SM count.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

constexpr int kSMs = 148;

inline int grid_for(long long n, int block, int per_sm = 8) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)kSMs * per_sm;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

__global__ void copy2d_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                              long long rows, int cols, int accumulate) {
  long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols;
    int c = (int)(i - r * cols);
    float v = src[r * lds + c];
    float* d = dst + r * ldd + c;
    *d = accumulate ? (*d + v) : v;
  }
}

// 16-byte variant (aligned rows, cols % 4 == 0): the decoder's broadcast of the per-window term over 32 steps and the in-place
// recurrence's input copy move 33 MB each
__global__ void __launch_bounds__(256) copy2d_vec4_kernel(const float4* __restrict__ src, long long lds4, float4* __restrict__ dst, long long ldd4,
                                                          long long rows, int cols4, int accumulate) {
  const long long total = rows * cols4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4);
    float4 v = __ldg(src + r * lds4 + c);
    float4* d = dst + r * ldd4 + c;
    if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    *d = v;
  }
}

__global__ void fill_kernel(float* __restrict__ dst, long long n, float value) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = value;
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}

// scalar combine: out = sum_i w_i * x_i[0] (forward), out[i] = w_i * g[0] (backward); n <= 8 terms, one thread
struct WeightedTerms { const float* x[8]; float w[8]; int n; };
__global__ void weighted_sum_kernel(const WeightedTerms t, float* __restrict__ out) {
  float s = 0.f;
  for (int i = 0; i < t.n; ++i) s = fmaf(t.w[i], *t.x[i], s);
  *out = s;
}
__global__ void weighted_fanout_kernel(const float* __restrict__ g, const WeightedTerms t, float* __restrict__ out) {
  if ((int)threadIdx.x < t.n) out[threadIdx.x] = t.w[threadIdx.x] * *g;
}

__global__ void relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dz[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// stage 1: block (32 cols x 8 row-lanes) sums its row slab; partial[slab][cols]
__global__ void colsum_partial_kernel(const float* __restrict__ x, long long ld, long long rows, int cols,
                                      double* __restrict__ partial, long long rows_per_slab) {
  // bias gradients sum up to millions of cancelling terms: accumulate in double (the kernel is HBM-bound anyway)
  __shared__ double red[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long long r0 = (long long)blockIdx.y * rows_per_slab;
  long long r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  double s = 0.0;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += (double)x[r * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    partial[(long long)blockIdx.y * cols + c] = t;
  }
}
// vectorised stage 1 (cols % 4 == 0, 16-byte aligned rows): thread = 4 adjacent columns, block = 128 columns x 8 row lanes,
// 4 rows in flight per thread (4 independent 16-byte loads) -- enough bytes in flight to approach HBM bandwidth
__global__ void __launch_bounds__(256) colsum_partial4_kernel(const float* __restrict__ x, long long ld, long long rows, int cols,
                                                              double* __restrict__ partial, long long rows_per_slab) {
  __shared__ double red[8][32][4];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_slab;
  const long long r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < cols) {
    long long r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + (r + 8) * ld + c));
      const float4 d = __ldg(reinterpret_cast<const float4*>(x + (r + 16) * ld + c));
      const float4 e = __ldg(reinterpret_cast<const float4*>(x + (r + 24) * ld + c));
      s0 += ((double)a.x + (double)b.x) + ((double)d.x + (double)e.x);
      s1 += ((double)a.y + (double)b.y) + ((double)d.y + (double)e.y);
      s2 += ((double)a.z + (double)b.z) + ((double)d.z + (double)e.z);
      s3 += ((double)a.w + (double)b.w) + ((double)d.w + (double)e.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      s0 += (double)a.x; s1 += (double)a.y; s2 += (double)a.z; s3 += (double)a.w;
    }
  }
  red[threadIdx.y][threadIdx.x][0] = s0; red[threadIdx.y][threadIdx.x][1] = s1;
  red[threadIdx.y][threadIdx.x][2] = s2; red[threadIdx.y][threadIdx.x][3] = s3;
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;          // 128 columns of the block, one per thread
  if (t < 128 && blockIdx.x * 128 + t < cols) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += red[j][t >> 2][t & 3];
    partial[(long long)blockIdx.y * cols + blockIdx.x * 128 + t] = acc;
  }
}
// Short-and-wide sums (the decoder's sum over time: 32 rows x B*H columns): thread = 4 adjacent columns, all rows, 8 independent
// 16-byte loads in flight; one pass, no partials.
__global__ void __launch_bounds__(256) colsum_wide_kernel(const float* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out,
                                                          int accumulate) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = 0;
  for (; r + 8 <= rows; r += 8) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(x + (long long)(r + k) * ld + c));
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0 += (double)v[k].x; s1 += (double)v[k].y; s2 += (double)v[k].z; s3 += (double)v[k].w; }
  }
  for (; r < rows; ++r) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (long long)r * ld + c));
    s0 += (double)a.x; s1 += (double)a.y; s2 += (double)a.z; s3 += (double)a.w;
  }
  float4* o = reinterpret_cast<float4*>(out + c);
  float4 res = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
  if (accumulate) { const float4 p = *o; res.x += p.x; res.y += p.y; res.z += p.z; res.w += p.w; }
  *o = res;
}
// One-launch column sum for the bias gradients (rows <= a few thousand): a thread-block cluster of 8 CTAs splits the rows of a
// 128-column block, every CTA writes its 128 double partials into rank 0's shared memory (DSMEM), rank 0 adds them in rank order.
// Deterministic, no global scratch, no second kernel (33 of these per train step).
constexpr int CS_CLUSTER = 8;
__global__ void __cluster_dims__(1, CS_CLUSTER, 1) __launch_bounds__(256)
    colsum_cluster_kernel(const float* __restrict__ x, long long ld, long long rows, int cols, float* __restrict__ out, int accumulate) {
  __shared__ double red[8][32][4];
  __shared__ double part[CS_CLUSTER][128];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const long long per = (rows + CS_CLUSTER - 1) / CS_CLUSTER;
  const long long r0 = (long long)blockIdx.y * per;
  const long long r1 = r0 + per < rows ? r0 + per : rows;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < cols) {
    long long r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + (r + 8) * ld + c));
      const float4 d = __ldg(reinterpret_cast<const float4*>(x + (r + 16) * ld + c));
      const float4 e = __ldg(reinterpret_cast<const float4*>(x + (r + 24) * ld + c));
      s0 += ((double)a.x + (double)b.x) + ((double)d.x + (double)e.x);
      s1 += ((double)a.y + (double)b.y) + ((double)d.y + (double)e.y);
      s2 += ((double)a.z + (double)b.z) + ((double)d.z + (double)e.z);
      s3 += ((double)a.w + (double)b.w) + ((double)d.w + (double)e.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
      s0 += (double)a.x; s1 += (double)a.y; s2 += (double)a.z; s3 += (double)a.w;
    }
  }
  red[threadIdx.y][threadIdx.x][0] = s0; red[threadIdx.y][threadIdx.x][1] = s1;
  red[threadIdx.y][threadIdx.x][2] = s2; red[threadIdx.y][threadIdx.x][3] = s3;
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;          // 128 columns of the block, one per thread
  if (t < 128) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += red[j][t >> 2][t & 3];
    uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[blockIdx.y][t]), remote;   // blockIdx.y == rank (grid.y == cluster size)
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(acc) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (blockIdx.y == 0 && t < 128 && blockIdx.x * 128 + t < cols) {
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < CS_CLUSTER; ++j) tot += part[j][t];
    float* o = out + blockIdx.x * 128 + t;
    *o = accumulate ? *o + (float)tot : (float)tot;
  }
}
__global__ void colsum_final_kernel(const double* __restrict__ partial, int slabs, int cols, float* __restrict__ out, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int j = 0; j < slabs; ++j) s += partial[(long long)j * cols + c];
  out[c] = accumulate ? out[c] + (float)s : (float)s;
}

// per-frame [HW, C] <-> [C, HW] transposes through a padded shared tile
__device__ __forceinline__ float ld_el(const float* p) { return *p; }
__device__ __forceinline__ float ld_el(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_el(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_el(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename TI, typename TO, typename TM>
__global__ void transpose_frames_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int rows, int cols,
                                        const TM* __restrict__ mask) {
  // src frame is [rows, cols] row-major, dst frame is [cols, rows]; mask (optional) has dst layout
  __shared__ float tile[32][33];
  long long fbase = (long long)blockIdx.z * rows * cols;
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? ld_el(src + fbase + (long long)r * cols + c) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      long long o = fbase + (long long)c * rows + r;
      float v = tile[threadIdx.x][j];
      if (mask) v = ld_el(mask + o) > 0.f ? v : 0.f;
      st_el(dst + o, v);
    }
  }
}

// Small frames (the gripper trunk's 7 x 7 x 64 maps: the 32 x 32 tiles above are half empty and every access is a scalar): one
// block per frame, the whole frame through shared memory, both the read and the write linear in memory.
template <typename TI, typename TO, typename TM>
__global__ void __launch_bounds__(256) transpose_small_frames_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int rows, int cols,
                                                                     const TM* __restrict__ mask) {
  extern __shared__ float tsf_tile[];                 // [rows][pitch], pitch odd: conflict-free column reads
  const int pitch = cols | 1, n = rows * cols;
  const long long fbase = (long long)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int r = i / cols, c = i - r * cols;
    tsf_tile[r * pitch + c] = ld_el(src + fbase + i);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < n; o += 256) {
    const int c = o / rows, r = o - c * rows;
    float v = tsf_tile[r * pitch + c];
    if (mask) v = ld_el(mask + fbase + o) > 0.f ? v : 0.f;
    st_el(dst + fbase + o, v);
  }
}
template <typename TI, typename TO, typename TM>
static bool transpose_small(const TI* src, TO* dst, int F, int rows, int cols, const TM* mask, cudaStream_t st) {
  const size_t smem = (size_t)rows * (cols | 1) * sizeof(float);
  if (smem > 48 * 1024) return false;
  transpose_small_frames_kernel<TI, TO, TM><<<F, 256, smem, st>>>(src, dst, rows, cols, mask);
  return true;
}

// dst[d1, d0, :D2] (+)= src[d0*src_s0 + d1*src_s1 + :D2]   (batch-major <-> time-major row shuffles)
__global__ void transpose01_kernel(const float* __restrict__ src, long long src_s0, long long src_s1, float* __restrict__ dst, long long dst_ld,
                                   int D0, int D1, int D2, int accumulate) {
  long long total = (long long)D0 * D1 * D2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % D2);
    long long r = i / D2;          // r = d1*D0 + d0 (destination row)
    int d0 = (int)(r % D0), d1 = (int)(r / D0);
    float v = src[(long long)d0 * src_s0 + (long long)d1 * src_s1 + c];
    float* d = dst + r * dst_ld + c;
    *d = accumulate ? (*d + v) : v;
  }
}

__global__ void permute_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I, int KH, int KW,
                                      int dir, int accumulate) {
  int total = O * I * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // i indexes OIHW
    int kw = i % KW, t = i / KW;
    int kh = t % KH; t /= KH;
    int ci = t % I, o = t / I;
    int ohwi = ((o * KH + kh) * KW + kw) * I + ci;
    int hwoi = ((kh * KW + kw) * O + o) * I + ci;
    if (dir == 0) dst[ohwi] = src[i];
    else if (dir == 1) dst[i] = accumulate ? dst[i] + src[ohwi] : src[ohwi];
    else dst[hwoi] = src[i];
  }
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// `epoch` (optional device counter, bumped once per training step) moves the whole counter range: a step captured in
// a CUDA graph draws fresh noise on every replay although its launch arguments are frozen.
__global__ void philox_uniform_kernel(float* __restrict__ out, long long n, uint64_t seed, uint64_t offset,
                                      const unsigned long long* __restrict__ epoch) {
  if (epoch) offset += (uint64_t)(*epoch) << 40;
  long long nq = (n + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4(seed, offset + (uint64_t)q, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long i = q * 4 + j;
      if (i < n) out[i] = u01(r[j]);
    }
  }
}
__global__ void dropout_mask_kernel(unsigned char* __restrict__ out, long long n, float p, uint64_t seed, uint64_t offset,
                                    const unsigned long long* __restrict__ epoch) {
  if (epoch) offset += (uint64_t)(*epoch) << 40;
  long long nq = (n + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4(seed, offset + (uint64_t)q, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long i = q * 4 + j;
      if (i < n) out[i] = u01(r[j]) >= p ? 1 : 0;
    }
  }
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }

// ---------------------------------------------------------------- Adam (torch.optim.Adam, amsgrad=False, maximize=False)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale, const unsigned long long* __restrict__ step_dev, int step_bias,
                            const float* __restrict__ lr_dev) {
  if (lr_dev) lr = *lr_dev;   // learning rate lives on the device: an LR scheduler keeps working across graph replays
  if (step_dev) {  // step number lives on the device (CUDA-graph replays): bias corrections computed here
    const double st = (double)(*step_dev) + (double)step_bias;
    bc1 = (float)(1.0 - pow((double)b1, st));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, st));
  }
  auto upd = [&](float gi, float& pi, float& mi, float& vi) {
    gi *= grad_scale;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = mi + (1.f - b1) * (gi - mi);                     // exp_avg.lerp_(grad, 1-beta1)
    vi = b2 * vi + (1.f - b2) * gi * gi;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - (lr / bc1) * (mi / denom);
  };
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  // 16-byte accesses (the arenas are 32-byte aligned): 4 parameters per thread and iteration, same arithmetic per element
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = tid; i < n4; i += nth) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(g[i], pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// flat fp32 -> bf16 mirror (GEMM operand copies); 8 elements per thread when both pointers are 16-byte aligned
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n, int vec) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n8 = n >> 3;
    for (long long i = tid; i < n8; i += nthreads) {
      float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
      uint4 o;
      o.x = *reinterpret_cast<unsigned*>(&h0); o.y = *reinterpret_cast<unsigned*>(&h1);
      o.z = *reinterpret_cast<unsigned*>(&h2); o.w = *reinterpret_cast<unsigned*>(&h3);
      reinterpret_cast<uint4*>(dst)[i] = o;
    }
    for (long long i = (n8 << 3) + tid; i < n; i += nthreads) dst[i] = __float2bfloat16_rn(src[i]);
  } else {
    for (long long i = tid; i < n; i += nthreads) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// strided rows -> compact bf16 rows: dst[r*ld_dst + c] = bf16(src[r*ld_src + c]), c < cols (pad columns are not written)
__global__ void f32_to_bf16_2d_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                      long long ld_dst, long long rows, int cols, int vec) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const int c4 = cols >> 2;
    const long long total = rows * c4;
    for (long long i = tid; i < total; i += nthreads) {
      const long long r = i / c4;
      const int c = (int)(i - r * c4) << 2;
      float4 a = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
      __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
      uint2 o;
      o.x = *reinterpret_cast<unsigned*>(&h0); o.y = *reinterpret_cast<unsigned*>(&h1);
      *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = o;
    }
  } else {
    const long long total = rows * cols;
    for (long long i = tid; i < total; i += nthreads) {
      const long long r = i / cols;
      const int c = (int)(i - r * cols);
      dst[r * ld_dst + c] = __float2bfloat16_rn(src[r * ld_src + c]);
    }
  }
}

}  // namespace

extern "C" {

int hulc2_copy2d(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols, int accumulate,
                 cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HULC2_OK;
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0 && rows * cols >= 4096) {
    long long blocks = (rows * (cols / 4) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    copy2d_vec4_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), lds / 4, reinterpret_cast<float4*>(dst), ldd / 4, rows,
                                                    cols / 4, accumulate);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  copy2d_kernel<<<grid_for(rows * cols, 256), 256, 0, st>>>(src, lds, dst, ldd, rows, cols, accumulate);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  const int vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  long long blocks = ((vec ? (n >> 3) : n) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  f32_to_bf16_kernel<<<(int)blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n, vec);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_f32_to_bf16_2d(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HULC2_OK;
  const int vec = ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dst) & 7) == 0) && cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0;
  long long work = vec ? rows * (cols >> 2) : rows * (long long)cols;
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_2d_kernel<<<(int)blocks, 256, 0, st>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, vec);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_fill(float* dst, long long n, float value, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  fill_kernel<<<grid_for(n, 256), 256, 0, st>>>(dst, n, value);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_axpy(const float* x, float* y, long long n, float a, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  axpy_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, y, n, a);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_weighted_sum(const float* const* xs, const float* w, int n, float* out, cudaStream_t st) {
  if (n <= 0 || n > 8 || !xs || !w || !out) { hulc2_set_error("weighted_sum: 1..8 terms"); return HULC2_EINVAL; }
  WeightedTerms t{};
  t.n = n;
  for (int i = 0; i < n; ++i) { t.x[i] = xs[i]; t.w[i] = w[i]; }
  weighted_sum_kernel<<<1, 1, 0, st>>>(t, out);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_weighted_fanout(const float* g, const float* w, int n, float* out, cudaStream_t st) {
  if (n <= 0 || n > 8 || !g || !w || !out) { hulc2_set_error("weighted_fanout: 1..8 terms"); return HULC2_EINVAL; }
  WeightedTerms t{};
  t.n = n;
  for (int i = 0; i < n; ++i) t.w[i] = w[i];
  weighted_fanout_kernel<<<1, 32, 0, st>>>(g, t, out);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_relu_mask(const float* dy, const float* y, float* dz, long long n, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  relu_mask_kernel<<<grid_for(n, 256), 256, 0, st>>>(dy, y, dz, n);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_colsum(const float* x, long long ld, long long rows, int cols, float* out, int accumulate, void* workspace,
                 long long workspace_bytes, cudaStream_t st) {
  if (cols <= 0) return HULC2_OK;
  const bool vec4 = (cols % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  int colblocks = hulc2_cdiv(cols, vec4 ? 128 : 32);
  if (vec4 && rows > 0 && rows <= 64 && cols >= 32768 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
    colsum_wide_kernel<<<hulc2_cdiv(cols / 4, 256), 256, 0, st>>>(x, ld, (int)rows, cols, out, accumulate);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  if (vec4 && rows <= 32768 && rows > 0) {
    colsum_cluster_kernel<<<dim3(colblocks, CS_CLUSTER), dim3(32, 8), 0, st>>>(x, ld, rows, cols, out, accumulate);
    HULC2_CHECK_LAUNCH();
    return HULC2_OK;
  }
  long long max_slabs = workspace ? workspace_bytes / ((long long)cols * sizeof(double)) : 0;
  if (max_slabs < 1) { hulc2_set_error("colsum: workspace too small"); return HULC2_EWORKSPACE; }
  long long slabs = ((vec4 ? 6LL : 2LL) * kSMs + colblocks - 1) / colblocks;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs > (rows + 63) / 64) slabs = (rows + 63) / 64;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  long long per = (rows + slabs - 1) / slabs;
  if (per < 1) per = 1;
  slabs = rows > 0 ? (rows + per - 1) / per : 1;
  if (vec4) colsum_partial4_kernel<<<dim3(colblocks, (unsigned)slabs), dim3(32, 8), 0, st>>>(x, ld, rows, cols, (double*)workspace, per);
  else colsum_partial_kernel<<<dim3(colblocks, (unsigned)slabs), dim3(32, 8), 0, st>>>(x, ld, rows, cols, (double*)workspace, per);
  HULC2_CHECK_LAUNCH();
  colsum_final_kernel<<<hulc2_cdiv(cols, 128), 128, 0, st>>>((const double*)workspace, (int)slabs, cols, out, accumulate);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_nhwc_to_nchw(const float* src, float* dst, int F, int HW, int C, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  dim3 grid(hulc2_cdiv(C, 32), hulc2_cdiv(HW, 32), F);
  transpose_frames_kernel<float, float, float><<<grid, dim3(32, 8), 0, st>>>(src, dst, HW, C, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_nchw_to_nhwc(const float* src, float* dst, int F, int HW, int C, const float* mask, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  dim3 grid(hulc2_cdiv(HW, 32), hulc2_cdiv(C, 32), F);
  transpose_frames_kernel<float, float, float><<<grid, dim3(32, 8), 0, st>>>(src, dst, C, HW, mask);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
// bf16 conv trunk <-> fp32 heads: y3 (bf16 NHWC) -> nn.Flatten order (fp32 [F, C*HW]) and its gradient back (ReLU-masked by y3)
int hulc2_nhwc_bf16_to_nchw(const void* src, float* dst, int F, int HW, int C, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (transpose_small<__nv_bfloat16, float, float>((const __nv_bfloat16*)src, dst, F, HW, C, nullptr, st)) { HULC2_CHECK_LAUNCH(); return HULC2_OK; }
  dim3 grid(hulc2_cdiv(C, 32), hulc2_cdiv(HW, 32), F);
  transpose_frames_kernel<__nv_bfloat16, float, float><<<grid, dim3(32, 8), 0, st>>>((const __nv_bfloat16*)src, dst, HW, C, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_nchw_to_nhwc_bf16(const float* src, void* dst, int F, int HW, int C, const void* mask, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  if (transpose_small<float, __nv_bfloat16, __nv_bfloat16>(src, (__nv_bfloat16*)dst, F, C, HW, (const __nv_bfloat16*)mask, st)) { HULC2_CHECK_LAUNCH(); return HULC2_OK; }
  dim3 grid(hulc2_cdiv(HW, 32), hulc2_cdiv(C, 32), F);
  transpose_frames_kernel<float, __nv_bfloat16, __nv_bfloat16><<<grid, dim3(32, 8), 0, st>>>(src, (__nv_bfloat16*)dst, C, HW, (const __nv_bfloat16*)mask);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_transpose01(const float* src, long long src_s0, long long src_s1, float* dst, long long dst_ld, int D0, int D1, int D2,
                      int accumulate,
                      cudaStream_t st) {
  long long total = (long long)D0 * D1 * D2;
  if (total <= 0) return HULC2_OK;
  transpose01_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, src_s0, src_s1, dst, dst_ld, D0, D1, D2, accumulate);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_permute_conv_weight(const float* src, float* dst, int O, int I, int KH, int KW, int dir, int accumulate,
                              cudaStream_t st) {
  int total = O * I * KH * KW;
  if (total <= 0) return HULC2_OK;
  permute_weight_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, dst, O, I, KH, KW, dir, accumulate);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_philox_uniform(float* out, long long n, unsigned long long seed, unsigned long long offset, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  philox_uniform_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(out, n, seed, offset, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_dropout_mask(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                       cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  dropout_mask_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(out, n, p, seed, offset, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  if (step < 1) { hulc2_set_error("adam: step must be >= 1"); return HULC2_EINVAL; }
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                    (float)sqrt(bc2), grad_scale, nullptr, 0, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
// ---- CUDA-graph friendly variants: the step / noise epoch is a device counter
int hulc2_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, const unsigned long long* step_counter, int step_bias,
                        float grad_scale, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  if (!step_counter) { hulc2_set_error("adam_step_dev: null step counter"); return HULC2_EINVAL; }
  adam_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_scale,
                                                    step_counter, step_bias, nullptr);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_adam_step_graph(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float beta1, float beta2,
                          float eps, float weight_decay, const unsigned long long* step_counter, int step_bias,
                          float grad_scale, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  if (!step_counter || !lr_dev) { hulc2_set_error("adam_step_graph: null step counter / learning-rate pointer"); return HULC2_EINVAL; }
  adam_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_scale,
                                                    step_counter, step_bias, lr_dev);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_philox_uniform_ep(float* out, long long n, unsigned long long seed, unsigned long long offset,
                            const unsigned long long* epoch, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  philox_uniform_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(out, n, seed, offset, epoch);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_dropout_mask_ep(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                          const unsigned long long* epoch, cudaStream_t st) {
  if (n <= 0) return HULC2_OK;
  dropout_mask_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(out, n, p, seed, offset, epoch);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
int hulc2_counter_add(unsigned long long* counter, unsigned long long inc, cudaStream_t st) {
  counter_add_kernel<<<1, 1, 0, st>>>(counter, inc);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
