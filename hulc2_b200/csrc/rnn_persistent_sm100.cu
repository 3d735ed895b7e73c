// Persistent decoder recurrence for sm_100a (nn.RNN relu cell, hulc2/models/decoders/utils/rnn.py:5-14).
//
//   forward : h[t]  = relu(pre[t] + h[t-1] W_hh^T)                     t = 0 .. S-1
//   backward: dz[t] = (dh[t] + dz[t+1] W_hh) * (h[t] > 0)              t = S-1 .. 0   (in place over dh)
//
// The per-step GEMM is [B x H] x [H x H] with B <= 128: launched step by step it is pure latency (64
// dependent launches per layer, each re-streaming the 8 MB weight).  Here ONE kernel runs all S steps:
//   * grid = H/16 CTAs (128 for H = 2048, one per SM); CTA c owns output columns [16c, 16c+16) and keeps its
//     W slice [16 x H] resident in shared memory as bf16 K-major SWIZZLE_128B tiles for the whole sequence;
//   * per step the previous state (a bf16 copy written by the epilogue) streams through a 4-stage cp.async ring
//     as the A operand; one thread issues tcgen05.mma (128 x 16 x 16, fp32 accumulate in TMEM);
//   * the epilogue adds pre[t] / dh[t], applies ReLU / the ReLU mask, writes the fp32 state (for the heads and
//     weight-gradient GEMMs) and its bf16 copy (next step's operand);
//   * steps are separated by a grid-wide release/acquire barrier on a global counter (all CTAs co-resident:
//     grid <= SM count, 1 CTA per SM by shared-memory footprint).
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

constexpr int NS = 16;        // output columns per CTA
constexpr int BM = 128;       // UMMA M (rows >= B are zero)
constexpr int KT = 64;        // k per tile (128-byte swizzle row)
constexpr int NT = 256;
constexpr int STAGES = 4;
constexpr uint32_t A_TILE = BM * 128;   // 16 KB
constexpr uint32_t W_TILE = NS * 128;   // 2 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO = 1024, version 1
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct RnnParams {
  const float* add;          // [S,B,H] pre-activations (fwd) / incoming gradients (bwd; aliases `out`)
  const float* w;            // [H,H] W_hh
  const float* init;         // [B,H] fp32 initial state (fwd: h0 or null; bwd: null)
  const float* mask;         // [S,B,H] or null: output is zeroed where mask <= 0 (bwd: h)
  float* out;                // [S,B,H] fp32 states
  __nv_bfloat16* outb;       // [S,B,H] bf16 copy (workspace)
  float* final_out;          // bwd only: dh0 [B,H] = dz[0] W_hh (or null)
  unsigned int* counter;     // grid barrier (zeroed before launch)
  int S, B, H;
  int relu, reverse, transpose_w;
};

__device__ unsigned int g_persistent_rnn_error = 0;   // see rnn_cluster_sm100.cu: flagged instead of trapping the context

// grid-wide barrier: every CTA arrives once per call; generation g waits for g * gridDim.x arrivals
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x == 0) {
    unsigned int v;
    unsigned int spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((++spins & 1023u) == 0) {
        if (spins > (1u << 23)) atomicExch(&g_persistent_rnn_error, 1u);
        if (*reinterpret_cast<volatile unsigned int*>(&g_persistent_rnn_error)) break;
      }
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) rnn_persistent_kernel(const RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, B = p.B, S = p.S;
  const int n0 = blockIdx.x * NS;
  const int nkt = H / KT;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_smem = base;                       // nkt tiles of 2 KB
  const uint32_t a_smem = base + (uint32_t)nkt * W_TILE;   // STAGES tiles of 16 KB (1024-aligned since W_TILE*nkt is)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&mma_done[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident W slice: element (n, k) -> fwd W[(n0+n)*H + k], bwd W[k*H + n0+n]; 16-byte chunks of 8 k
  for (int ch = tid; ch < NS * (H / 8); ch += NT) {
    int n = ch % NS, kc = ch / NS;         // consecutive threads -> consecutive n (coalesced for the transposed read)
    if (!p.transpose_w) { kc = ch % (H / 8); n = ch / (H / 8); }
    int k = kc * 8;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = p.transpose_w ? __ldg(p.w + (long long)(k + e) * H + n0 + n) : __ldg(p.w + (long long)(n0 + n) * H + k + e);
    uint32_t dst = w_smem + (uint32_t)(k / KT) * W_TILE + swz(n, (k % KT) / 8);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2(f[0], f[1])), "r"(pack2(f[2], f[3])), "r"(pack2(f[4], f[5])), "r"(pack2(f[6], f[7])) : "memory");
  }
  // zero the A ring once: rows >= B are never written afterwards
  for (uint32_t o = tid * 16; o < STAGES * A_TILE; o += NT * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_smem + o), "r"(0u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  uint32_t uses[STAGES];
#pragma unroll
  for (int s = 0; s < STAGES; ++s) uses[s] = 0;
  unsigned int generation = 0;
  const long long step = (long long)B * H;
  // epilogue ownership: warps 0..3 cover TMEM lanes (rows) 32w .. 32w+31, all 16 columns
  const int erow = warp * 32 + lane;
  const bool eactive = warp < 4 && erow < B;

  // A-tile loader: chunk c = tid & 7 of rows (tid >> 3) + 32 i
  auto load_tile_bf16 = [&](const __nv_bfloat16* src, int kt, int stage) {
    const int c = tid & 7;
#pragma unroll
    for (int i = 0; i < BM / 32; ++i) {
      int r = (tid >> 3) + 32 * i;
      if (r < B) cp_async16(a_smem + stage * A_TILE + swz(r, c), src + (long long)r * H + kt * KT + c * 8);
    }
  };
  auto load_tile_f32 = [&](const float* src, int kt, int stage) {
    const int c = tid & 7;
#pragma unroll
    for (int i = 0; i < BM / 32; ++i) {
      int r = (tid >> 3) + 32 * i;
      if (r < B) {
        const float4* s4 = reinterpret_cast<const float4*>(src + (long long)r * H + kt * KT + c * 8);
        float4 a = __ldcg(s4), b = __ldcg(s4 + 1);
        uint32_t dst = a_smem + stage * A_TILE + swz(r, c);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2(a.x, a.y)), "r"(pack2(a.z, a.w)), "r"(pack2(b.x, b.y)), "r"(pack2(b.z, b.w)) : "memory");
      }
    }
  };

  for (int it = 0; it <= S; ++it) {
    // it < S: produce state t;  it == S (bwd with final_out only): dh0 = dz[0] W_hh, no add/mask
    const bool final_pass = it == S;
    if (final_pass && !p.final_out) break;
    const int t = p.reverse ? S - 1 - it : it;
    const int tprev = p.reverse ? t + 1 : t - 1;
    const bool has_prev = final_pass ? true : (it > 0 || p.init != nullptr);
    const __nv_bfloat16* prev_b = nullptr;
    const float* prev_f = nullptr;
    if (final_pass) prev_b = p.outb;                       // dz[0]
    else if (it > 0) prev_b = p.outb + (long long)tprev * step;
    else if (p.init) prev_f = p.init;

    // prefetch the epilogue addend for this thread's row (overlaps the k loop)
    float addv[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) addv[j] = 0.f;
    if (eactive && !final_pass) {
      const float4* a4 = reinterpret_cast<const float4*>(p.add + (long long)t * step + (long long)erow * H + n0);
#pragma unroll
      for (int j = 0; j < NS / 4; ++j) {
        float4 v = __ldcg(a4 + j);
        addv[4 * j] = v.x; addv[4 * j + 1] = v.y; addv[4 * j + 2] = v.z; addv[4 * j + 3] = v.w;
      }
    }

    if (has_prev) {
      if (it > 0 || final_pass) grid_wait(p.counter, generation * gridDim.x);   // previous state complete everywhere
      // multistage ring over the nkt k-tiles
      for (int pre = 0; pre < STAGES - 1; ++pre) {
        if (pre < nkt) {
          const int s = pre % STAGES;
          if (uses[s] > 0) mbar_wait(smem_u32(&mma_done[s]), (uses[s] - 1) & 1);
          if (prev_b) load_tile_bf16(prev_b, pre, s); else load_tile_f32(prev_f, pre, s);
        }
        cp_async_commit();
      }
      for (int kt = 0; kt < nkt; ++kt) {
        const int nxt = kt + STAGES - 1;
        if (nxt < nkt) {
          const int s = nxt % STAGES;
          if (uses[s] > 0) mbar_wait(smem_u32(&mma_done[s]), (uses[s] - 1) & 1);   // MMAs that read this stage retired
          if (prev_b) load_tile_bf16(prev_b, nxt, s); else load_tile_f32(prev_f, nxt, s);
        }
        cp_async_commit();
        cp_async_wait<STAGES - 1>();      // tile kt has landed (this thread's part)
        fence_proxy_async();
        __syncthreads();
        const int s = kt % STAGES;
        if (tid == 0) {
          tc_fence_after();
          const uint64_t ad = make_desc(a_smem + s * A_TILE), bd = make_desc(w_smem + kt * W_TILE);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            uint32_t acc = (kt > 0 || k > 0) ? 1u : 0u;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                         ::"r"(tmem_d), "l"(ad + (uint64_t)(k * 2)), "l"(bd + (uint64_t)(k * 2)), "r"(IDESC), "r"(acc) : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[s])) : "memory");
        }
        uses[s] += 1;
      }
      const int ls = (nkt - 1) % STAGES;
      mbar_wait(smem_u32(&mma_done[ls]), (uses[ls] - 1) & 1);   // in-order completion: all MMAs of this step are done
      tc_fence_after();
    }

    // epilogue
    if (warp < 4) {
      float acc[NS];
      if (has_prev) {
        uint32_t r[NS];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem_d + ((uint32_t)(warp * 32) << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[j] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) acc[j] = 0.f;
      }
      if (eactive) {
        float v[NS];
        if (final_pass) {
#pragma unroll
          for (int j = 0; j < NS; ++j) v[j] = acc[j];
          float4* o4 = reinterpret_cast<float4*>(p.final_out + (long long)erow * H + n0);
#pragma unroll
          for (int j = 0; j < NS / 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          float mk[NS];
          if (p.mask) {
            const float4* m4 = reinterpret_cast<const float4*>(p.mask + (long long)t * step + (long long)erow * H + n0);
#pragma unroll
            for (int j = 0; j < NS / 4; ++j) {
              float4 q = __ldcg(m4 + j);
              mk[4 * j] = q.x; mk[4 * j + 1] = q.y; mk[4 * j + 2] = q.z; mk[4 * j + 3] = q.w;
            }
          }
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            float x = acc[j] + addv[j];
            if (p.relu) x = fmaxf(x, 0.f);
            if (p.mask) x = mk[j] > 0.f ? x : 0.f;
            v[j] = x;
          }
          float4* o4 = reinterpret_cast<float4*>(p.out + (long long)t * step + (long long)erow * H + n0);
#pragma unroll
          for (int j = 0; j < NS / 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          uint4* b4 = reinterpret_cast<uint4*>(p.outb + (long long)t * step + (long long)erow * H + n0);
          b4[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          b4[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        }
      }
      tc_fence_before();
    }
    if (!final_pass) {
      generation += 1;
      grid_arrive(p.counter);           // publishes this CTA's slice of state t
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u) : "memory");
}

}  // namespace

// returns HULC2_ENOTIMPL when the shape does not fit the persistent kernel (caller falls back to the step loop)
int hulc2_rnn_persistent_launch(const float* add, const float* w, const float* init, const float* mask, float* out, float* final_out,
                                int S, int B, int H, int relu, int reverse, int transpose_w, void* workspace, long long workspace_bytes,
                                cudaStream_t st) {
  if (B > BM || B <= 0 || S <= 0 || H % KT != 0 || H % NS != 0) return HULC2_ENOTIMPL;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  static int cached_sms = 0;
  if (!cached_sms) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); cached_sms = sms; }
  const int grid = H / NS;
  if (grid > cached_sms) return HULC2_ENOTIMPL;
  const long long need = (long long)S * B * H * 2 + 256;
  if (!workspace || workspace_bytes < need) return HULC2_ENOTIMPL;
  if (((uintptr_t)add | (uintptr_t)out | (uintptr_t)w | (uintptr_t)workspace | (uintptr_t)(mask ? mask : out) | (uintptr_t)(init ? init : out) |
       (uintptr_t)(final_out ? final_out : out)) & 15)
    return HULC2_ENOTIMPL;
  const int smem = (H / KT) * (int)W_TILE + STAGES * (int)A_TILE + 1024;
  static int configured = 0;               // largest dynamic shared memory size opted in so far (it grows with H)
  if (configured < smem) {
    if (cudaFuncSetAttribute(rnn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return HULC2_ENOTIMPL;
    }
    configured = smem;
  }
  RnnParams p;
  p.add = add; p.w = w; p.init = init; p.mask = mask; p.out = out; p.final_out = final_out;
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.outb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  p.S = S; p.B = B; p.H = H; p.relu = relu; p.reverse = reverse; p.transpose_w = transpose_w;
  if (cudaMemsetAsync(p.counter, 0, 4, st) != cudaSuccess) return HULC2_ELAUNCH;
  // cooperative launch: the grid barrier needs every CTA resident; if concurrent work holds SMs the launch fails cleanly
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  memset(&cfg, 0, sizeof(cfg));
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, rnn_persistent_kernel, p);
  if (le != cudaSuccess) {
    cudaGetLastError();
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) return HULC2_ENOTIMPL;
    hulc2_set_error(cudaGetErrorString(le));
    return HULC2_ELAUNCH;
  }
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_rnn_persistent_device_error(int clear) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_persistent_rnn_error, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (clear && v) { const unsigned int z = 0; cudaMemcpyToSymbol(g_persistent_rnn_error, &z, sizeof(z)); }
  return (int)v;
}
