// bf16 dense contractions for sm_100a, second generation: operands are bf16 in HBM (row-major mirrors of the fp32
// master tensors, written once by their producer), moved by TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes) into a
// multi-stage shared-memory ring, multiplied by tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulators in TMEM) and
// drained by four epilogue warps: straight from the TMEM row registers with 256-bit stores when the epilogue reads no
// per-element operand, else turned around through the (by then idle) ring so that the mask / keep / add / accumulate operands,
// the result and the split-K partials all move as coalesced 16-byte accesses.
//
// Warp roles (192 threads):  warp 0 = TMA producer (one elected lane),  warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2-5 = epilogue (TMEM lane quarter = warp & 3).  full/empty mbarriers per stage; the MMA warp recycles a stage
// with tcgen05.commit, so loads run STAGES k-tiles ahead of the tensor core and nothing in the main loop touches
// registers.  One output tile per CTA; small-smem configurations co-reside 2 per SM so one CTA's epilogue overlaps the
// other's main loop.  Skinny problems (weight-streaming M=B layers, weight gradients) are split along K over ~2 waves
// of CTAs and reduced by splitk_reduce_kernel.
//
// The same row-major bf16 matrix serves as a K-major operand (contraction along its rows' fast axis: forward /
// input-gradient A) or as an MN-major operand (contraction along its slow axis: W in the input gradient, both
// operands of a weight gradient); the only difference is the TMA box orientation and the UMMA descriptor.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"
#include "sm100.cuh"

using namespace hulc2;
using namespace sm100;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // bf16 elements per k-tile = one 128-byte swizzle row
constexpr int NT = 192;
constexpr int MAX_STAGES = 8;
constexpr uint32_t A_BYTES = BM * 128;
unsigned long long g_tma_gemms = 0;

struct TmaParams {
  int M, N, K;
  int ktiles, kt_per_split, splits, stages;
  Epilogue E;
  float* partial;               // [splits, M, N] when splits > 1
  __nv_bfloat16* C16;           // optional bf16 mirror of the output (same row addressing with ld16)
  long long ld16;
  int vec;                      // 16-byte epilogue accesses are legal for every pointer involved
  int vec8;                     // ... and so are 32-byte ones (8 fp32 columns per 256-bit access)
  float* rowsum;                // optional [M]: rowsum[m] = sum_k A(m,k) (an nn.Linear's bias gradient out of its weight-gradient GEMM)
  float* rowsum_partial;        // [splits, M] when splits > 1
  int cluster_splitk;           // splits > 1: the CTAs of one output tile form a thread-block cluster (1, 1, splits) and reduce their
                                // accumulators through distributed shared memory -- no partial buffer, no reduce kernel
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}

// one output element through the shared epilogue (+ optional bf16 mirror)
__device__ __forceinline__ void store_scalar(const TmaParams& p, float acc, int m, int n, long long crow) {
  float v = apply_epilogue(p.E, acc, m, n, crow);
  p.E.C[crow + n] = v;
  if (p.C16) p.C16[(long long)m * p.ld16 + n] = __float2bfloat16_rn(v);
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NT, 2) gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const TmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kt0 = blockIdx.z * p.kt_per_split;
  const int nkt = min(p.ktiles, kt0 + p.kt_per_split) - kt0;
  const int ST = p.stages;

  if (tid == 0) {
    // (row sums: a second issuer thread commits on the same barriers, see below)
    const uint32_t issuers = (p.rowsum != nullptr && blockIdx.y == 0) ? 2u : 1u;
    for (int s = 0; s < ST; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), issuers); }
    mbar_init(smem_u32(&acc_bar), issuers);
    mbar_fence_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  // Row sums of A ride along as one more accumulator: the CTAs of the first N-tile column multiply every A k-slice with a
  // constant all-ones B operand as well (N = 16, the narrowest UMMA at M = 128; all elements are equal, so its swizzled layout
  // is immaterial) into TMEM columns [BN, BN + 16): the "ones column" of the weight-gradient contraction, without touching
  // the operand layouts in HBM.
  const bool row_sums = p.rowsum != nullptr && blockIdx.y == 0;
  const uint32_t tmem_cols = p.rowsum ? 2u * BN : (uint32_t)BN;      // power of two >= 32
  const uint32_t ones_tile = tiles + (uint32_t)ST * STAGE_BYTES;     // 2 KB behind the ring (1024-byte aligned)
  if (row_sums) {
    for (int i = tid; i < 2048 / 16; i += NT)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_tile + 16u * i), "r"(0x3F803F80u) : "memory");
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkt; ++i) {
        const int s = i % ST;
        if (i >= ST) mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)((i / ST - 1) & 1));
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t a_tile = tiles + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
        const int k0 = (kt0 + i) * BK;
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
        if (A_MN) {
          tma_load_2d(a_tile, &tmA, bar, m0, k0);
          tma_load_2d(a_tile + 8192, &tmA, bar, m0 + 64, k0);
        } else {
          tma_load_2d(a_tile, &tmA, bar, k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_tile + j * 8192, &tmB, bar, n0 + 64 * j, k0);
        } else {
          tma_load_2d(b_tile, &tmB, bar, k0, n0);
        }
      }
    }
    __syncwarp();                                  // (the cluster barriers of the split-K reduction are warp-aligned)
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(BM, BN, A_MN, B_MN);
      constexpr uint64_t A_STEP = (A_MN ? 2048u : 32u) >> 4, B_STEP = (B_MN ? 2048u : 32u) >> 4;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % ST;
        mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((i / ST) & 1));
        tc_fence_after();
        const uint32_t a_tile = tiles + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
        const uint64_t ad = make_desc(a_tile, A_MN ? 8192u : 0u), bd = make_desc(b_tile, B_MN ? 8192u : 0u);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, ad + A_STEP * k, bd + B_STEP * k, IDESC, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(smem_u32(&empty_bar[s]));     // stage free once these MMAs have read it
      }
      umma_commit(smem_u32(&acc_bar));            // accumulator complete
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue warps
    if (warp == 2 && row_sums) {
      // The ones-operand MMAs come from a SECOND issuer thread (one thread issues a tcgen05.mma only every ~110 cycles: issued
      // behind the main MMAs they doubled the main loop of the 16 row-sum CTAs, and the whole grid waited for those -- ncu r02:
      // 2048 x 2048 x 4096 weight gradient 39 -> 48 us).  Lane 0 of the first epilogue warp, idle until the accumulator is
      // complete, follows the same full barriers; both issuers commit on the stage's empty barrier and on the accumulator barrier.
      if (lane == 0) {
        constexpr uint32_t IDESC_RS = make_idesc(BM, 16, A_MN, false);
        constexpr uint64_t A_STEP = (A_MN ? 2048u : 32u) >> 4;
        const uint64_t od = make_desc(ones_tile, 0u);
        for (int i = 0; i < nkt; ++i) {
          const int s = i % ST;
          mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((i / ST) & 1));
          tc_fence_after();
          const uint64_t ad = make_desc(tiles + s * STAGE_BYTES, A_MN ? 8192u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d + BN, ad + A_STEP * k, od, IDESC_RS, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[s]));
        }
        umma_commit(smem_u32(&acc_bar));
      }
      __syncwarp();
    }
    if (!p.cluster_splitk) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool mvalid = m < p.M;
    const Epilogue& E = p.E;
    const long long crow = mvalid ? c_row_off(E, m) : 0;
    // Transposed epilogue (r02).  tcgen05.ld hands every thread ONE ROW of the tile (32 columns per chunk), so stores straight
    // from those registers put the 32 lanes of an instruction on 32 different rows: 32 L1 wavefronts per instruction, a quarter
    // line each -- the big-output, small-K contractions (FFN 4096 x 2048 x 128: 50 MB out, 41 MB of mask / keep in) were bound by
    // exactly that (38 -> 33 us; the split-K contractions with M = 4096 rows 22 -> 18 us).  Each epilogue warp turns its 32 x 32 chunk around through shared memory (the
    // pipeline ring is idle once the accumulator barrier has fired; row pitch 36 floats: conflict-free both ways): lane ->
    // (row quad lane >> 3, column quad lane & 7), one instruction = 4 rows x 128 contiguous bytes = 4 full lines, every epilogue
    // operand and result moves with coalesced 16-byte accesses.  The ReLU-mask / dropout-keep operands are fetched one chunk
    // AHEAD (first chunk while the main loop still runs), as before.
    const int rq = lane >> 3, cq = lane & 7;
    const uint32_t stg = tiles + (uint32_t)(warp - 2) * (32u * 144u);
    const bool tp = p.vec && p.splits == 1;                                  // CTA-uniform
    const bool direct = p.vec8 && !(E.mask || E.keep || E.add || E.accumulate);
    const int mrow0 = m0 + q * 32 + rq;                                      // this lane's row in iteration it: mrow0 + 4 * it
    uint4 mk_next[8];
    uint32_t kp_next[8];
    auto prefetch = [&](int c) {
      if (!tp || !(E.mask || E.keep) || n0 + c + 32 > p.N || c >= BN) return;
      const int n = n0 + c + 4 * cq;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int mr = mrow0 + 4 * it;
        if (mr < p.M) {
          if (E.mask) mk_next[it] = __ldg(reinterpret_cast<const uint4*>(E.mask + (long long)mr * E.ld_mask + n));
          if (E.keep) kp_next[it] = __ldg(reinterpret_cast<const unsigned int*>(E.keep + (long long)mr * E.ld_keep + n));
        }
      }
    };
    prefetch(0);
    mbar_wait_relaxed(smem_u32(&acc_bar), 0);
    tc_fence_after();
    const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
    if (row_sums) {                               // block-uniform
      uint32_t rs[16];
      tmem_ld16_nowait(trow + (uint32_t)BN, rs);
      tmem_ld_wait();
      if (mvalid) {
        if (p.splits > 1) p.rowsum_partial[(long long)blockIdx.z * p.M + m] = __uint_as_float(rs[0]);
        else p.rowsum[m] = __uint_as_float(rs[0]);
      }
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (n0 + c >= p.N) break;                   // warp-uniform
      uint32_t r[32];
      tmem_ld16_nowait(trow + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
      tmem_ld16_nowait(trow + (uint32_t)(c + 16), *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
      tmem_ld_wait();
      if (direct && n0 + c + 32 <= p.N) {
        // no per-element operand to read (bias / ReLU / bf16 mirror only: plain outputs, weight gradients): the thread's row goes
        // out straight from its registers with 256-bit stores -- every 32-byte sector is written whole, and measured against the
        // turn-around below this is 2-10 us faster per launch (in-graph timeline: 4096 x 2048 x 182 22 vs 33 us, x 2048 56 vs 65)
        if (!mvalid) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int n = n0 + c + j;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = E.alpha * __uint_as_float(r[j + e]);
          uint32_t t[8];
          if (E.bias) { ldg256_nc(E.bias + n, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += __uint_as_float(t[e]); }
          if (E.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = __float_as_uint(v[e]);
          stg256(E.C + crow + n, t);
          if (p.C16)
            *reinterpret_cast<uint4*>(p.C16 + (long long)m * p.ld16 + n) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
        continue;
      }
      if (tp && n0 + c + 32 <= p.N) {             // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 144u + 16u * j), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                       "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
        uint4 mk[8];
        uint32_t kp[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) { mk[it] = mk_next[it]; kp[it] = kp_next[it]; }
        prefetch(c + 32);
        __syncwarp();
        const int n = n0 + c + 4 * cq;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (E.bias) bias4 = __ldg(reinterpret_cast<const float4*>(E.bias + n));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 ad[4], ca[4];
          long long cro[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {            // the loads of four row quads are issued before the first store
            const int mr = mrow0 + 4 * (4 * h + u);
            cro[u] = mr < p.M ? c_row_off(E, mr) : 0;
            if (mr < p.M) {
              if (E.add) ad[u] = *reinterpret_cast<const float4*>(E.add + (long long)mr * E.ld_add + n);
              if (E.accumulate) ca[u] = *reinterpret_cast<const float4*>(E.C + cro[u] + n);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int it = 4 * h + u;
            const int mr = mrow0 + 4 * it;
            float4 a;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                         : "r"(stg + (uint32_t)(4 * it + rq) * 144u + 16u * cq) : "memory");
            if (mr >= p.M) continue;
            float v[4] = {E.alpha * a.x, E.alpha * a.y, E.alpha * a.z, E.alpha * a.w};
            if (E.bias) { v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w; }
            if (E.add) { v[0] += ad[u].x; v[1] += ad[u].y; v[2] += ad[u].z; v[3] += ad[u].w; }
            if (E.accumulate) { v[0] += ca[u].x; v[1] += ca[u].y; v[2] += ca[u].z; v[3] += ca[u].w; }
            if (E.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
            if (E.mask) {
              v[0] = __uint_as_float(mk[it].x) > 0.f ? v[0] : 0.f; v[1] = __uint_as_float(mk[it].y) > 0.f ? v[1] : 0.f;
              v[2] = __uint_as_float(mk[it].z) > 0.f ? v[2] : 0.f; v[3] = __uint_as_float(mk[it].w) > 0.f ? v[3] : 0.f;
            }
            if (E.keep) {
              const uint32_t k4 = kp[it];
              v[0] = (k4 & 0xffu) ? v[0] * E.keep_scale : 0.f; v[1] = (k4 & 0xff00u) ? v[1] * E.keep_scale : 0.f;
              v[2] = (k4 & 0xff0000u) ? v[2] * E.keep_scale : 0.f; v[3] = (k4 & 0xff000000u) ? v[3] * E.keep_scale : 0.f;
            }
            *reinterpret_cast<float4*>(E.C + cro[u] + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.C16) *reinterpret_cast<uint2*>(p.C16 + (long long)mr * p.ld16 + n) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
          }
        }
        __syncwarp();                              // the staging rows are rewritten by the next chunk
        continue;
      }
      if (p.splits > 1 && p.vec && n0 + c + 32 <= p.N) {      // warp-uniform: split-K partials, same turn-around, no operands
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 144u + 16u * j), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                       "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
        __syncwarp();
        const int n = n0 + c + 4 * cq;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int mr = mrow0 + 4 * it;
          float4 a;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                       : "r"(stg + (uint32_t)(4 * it + rq) * 144u + 16u * cq) : "memory");
          if (mr < p.M) *reinterpret_cast<float4*>(p.partial + ((long long)blockIdx.z * p.M + mr) * p.N + n) = a;
        }
        __syncwarp();
        continue;
      }
      if (!mvalid) continue;
      if (p.splits > 1) {
        float* prow = p.partial + ((long long)blockIdx.z * p.M + m) * p.N;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c + j;
          if (n >= p.N) break;
          if (p.vec && n + 4 <= p.N) {
            *reinterpret_cast<float4*>(prow + n) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
            for (int e = 0; e < 4; ++e) if (n + e < p.N) prow[n + e] = __uint_as_float(r[j + e]);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = n0 + c + j;
        if (n >= p.N) break;
        if (p.vec && n + 4 <= p.N) {
          float v[4] = {E.alpha * __uint_as_float(r[j]), E.alpha * __uint_as_float(r[j + 1]), E.alpha * __uint_as_float(r[j + 2]), E.alpha * __uint_as_float(r[j + 3])};
          if (E.bias) { float4 t = __ldg(reinterpret_cast<const float4*>(E.bias + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.add) { float4 t = *reinterpret_cast<const float4*>(E.add + (long long)m * E.ld_add + n); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.accumulate) { float4 t = *reinterpret_cast<const float4*>(E.C + crow + n); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
          if (E.mask) {
            float4 t = *reinterpret_cast<const float4*>(E.mask + (long long)m * E.ld_mask + n);
            v[0] = t.x > 0.f ? v[0] : 0.f; v[1] = t.y > 0.f ? v[1] : 0.f; v[2] = t.z > 0.f ? v[2] : 0.f; v[3] = t.w > 0.f ? v[3] : 0.f;
          }
          if (E.keep) {
            uchar4 t = *reinterpret_cast<const uchar4*>(E.keep + (long long)m * E.ld_keep + n);
            v[0] = t.x ? v[0] * E.keep_scale : 0.f; v[1] = t.y ? v[1] * E.keep_scale : 0.f;
            v[2] = t.z ? v[2] * E.keep_scale : 0.f; v[3] = t.w ? v[3] * E.keep_scale : 0.f;
          }
          *reinterpret_cast<float4*>(E.C + crow + n) = make_float4(v[0], v[1], v[2], v[3]);
          if (p.C16) {
            uint2 h = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            *reinterpret_cast<uint2*>(p.C16 + (long long)m * p.ld16 + n) = h;
          }
        } else {
          for (int e = 0; e < 4; ++e) if (n + e < p.N) store_scalar(p, __uint_as_float(r[j + e]), m, n + e, crow);
        }
      }
    }
    }   // !cluster_splitk
  }
  if (p.cluster_splitk) {
    // ------------------------------------------------------------ split-K reduced through distributed shared memory
    // The `splits` CTAs of an output tile are one cluster.  CTA z finalises rows [z rp, (z + 1) rp) of the tile (rp = ceil(128 /
    // splits)): every CTA pushes each accumulator row into the shared memory of the row's owner (RED[source z][local row][BN + 4],
    // in the pipeline ring, idle by then), one cluster barrier, the owner adds the `splits` strips in source order (deterministic)
    // and runs the full epilogue with coalesced 16-byte rows.  Replaces [splits, M, N] fp32 partials through HBM / L2 plus a
    // second launch (49 split-K contractions per train step) -- implemented, parity-green (195 GPU tests) and OFF by default: slower
    // than what it replaces, see gemm_cluster_splitk().
    const Epilogue& E = p.E;
    const int splits = p.splits;
    const int rp = (BM + splits - 1) / splits;
    constexpr int PITCH = BN + 4;                              // floats; column BN carries the row sum of A
    const uint32_t z = cluster_ctarank();
    if (warp >= 2) {
      mbar_wait_relaxed(smem_u32(&acc_bar), 0);                // this CTA's MMAs have read the ring for the last time
      tc_fence_after();
    }
    cluster_sync_all();                                        // ... and so have those of every CTA of the cluster
    if (warp >= 2) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t remote = mapa(tiles + (uint32_t)(((int)z * rp + row % rp) * PITCH) * 4u, (uint32_t)(row / rp));
      const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= p.N) break;
        uint32_t r[32];
        tmem_ld16_nowait(trow + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        tmem_ld16_nowait(trow + (uint32_t)(c + 16), *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)(c + 4 * j) * 4u), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                       "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
      }
      if (row_sums) {
        uint32_t rs[16];
        tmem_ld16_nowait(trow + (uint32_t)BN, rs);
        tmem_ld_wait();
        asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(remote + (uint32_t)BN * 4u), "r"(rs[0]) : "memory");
      }
      tc_fence_before();
    }
    cluster_sync_all();                                        // every strip has landed
    if (warp >= 2) {
      const int te = tid - 64;                                 // 0 .. 127
      const int row_lo = (int)z * rp;
      const int nrows = min(rp, BM - row_lo);                  // (<= 0 for the CTAs behind a ragged split: nothing to finalise)
      constexpr int G = BN / 4;
      for (int item = te; item < nrows * G; item += 128) {
        const int lr = item / G, c4 = item - lr * G;
        const int mr = m0 + row_lo + lr, n = n0 + 4 * c4;
        if (mr >= p.M || n >= p.N) continue;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int sidx = 0; sidx < splits; ++sidx) {
          float4 a;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                       : "r"(tiles + (uint32_t)((sidx * rp + lr) * PITCH + 4 * c4) * 4u) : "memory");
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
        }
        const long long cro = c_row_off(E, mr);
        if (p.vec && n + 4 <= p.N) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] *= E.alpha;
          if (E.bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(E.bias + n)); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.add) { const float4 t = *reinterpret_cast<const float4*>(E.add + (long long)mr * E.ld_add + n); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.accumulate) { const float4 t = *reinterpret_cast<const float4*>(E.C + cro + n); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
          if (E.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
          if (E.mask) {
            const float4 t = *reinterpret_cast<const float4*>(E.mask + (long long)mr * E.ld_mask + n);
            v[0] = t.x > 0.f ? v[0] : 0.f; v[1] = t.y > 0.f ? v[1] : 0.f; v[2] = t.z > 0.f ? v[2] : 0.f; v[3] = t.w > 0.f ? v[3] : 0.f;
          }
          if (E.keep) {
            const uchar4 t = *reinterpret_cast<const uchar4*>(E.keep + (long long)mr * E.ld_keep + n);
            v[0] = t.x ? v[0] * E.keep_scale : 0.f; v[1] = t.y ? v[1] * E.keep_scale : 0.f;
            v[2] = t.z ? v[2] * E.keep_scale : 0.f; v[3] = t.w ? v[3] * E.keep_scale : 0.f;
          }
          *reinterpret_cast<float4*>(E.C + cro + n) = make_float4(v[0], v[1], v[2], v[3]);
          if (p.C16) *reinterpret_cast<uint2*>(p.C16 + (long long)mr * p.ld16 + n) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
        } else {
          for (int e = 0; e < 4; ++e) if (n + e < p.N) store_scalar(p, v[e], mr, n + e, cro);
        }
      }
      if (row_sums) {
        for (int lr = te; lr < nrows; lr += 128) {
          const int mr = m0 + row_lo + lr;
          if (mr >= p.M) continue;
          float rs = 0.f;
          for (int sidx = 0; sidx < splits; ++sidx) {
            float a;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(tiles + (uint32_t)((sidx * rp + lr) * PITCH + BN) * 4u) : "memory");
            rs += a;
          }
          p.rowsum[mr] = rs;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

// ---------------------------------------------------------------------------------------------- host side
// cuTensorMapEncodeTiled comes from the driver at run time (cudaGetDriverEntryPoint), so the library still loads --
// and exports every symbol -- on hosts without libcuda (build / CPU test containers).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

bool encode_map(CUtensorMap* tm, const void* base, long long inner, long long outer, long long ld_elems, int box_outer) {
  EncodeTiledFn cuTensorMapEncodeTiled = encode_fn();
  if (!cuTensorMapEncodeTiled) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = cuTensorMapEncodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TmaParams& p, int smem, cudaStream_t st) {
  auto kern = gemm_tma_kernel<BN, A_MN, B_MN>;
  static int configured = 0;  // per instantiation: largest dynamic smem opted in so far
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      hulc2_set_error("gemm_tma: cannot raise the dynamic shared memory limit");
      return HULC2_ELAUNCH;
    }
    configured = smem;
  }
  dim3 grid(hulc2_cdiv(p.M, BM), hulc2_cdiv(p.N, BN), p.splits);
  if (p.cluster_splitk) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.splits;
    cfg.gridDim = grid; cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) { cudaGetLastError(); hulc2_set_error(cudaGetErrorString(e)); return HULC2_ELAUNCH; }
    ++g_tma_gemms;
    return HULC2_OK;
  }
  kern<<<grid, NT, smem, st>>>(ta, tb, p);
  HULC2_CHECK_LAUNCH();
  ++g_tma_gemms;
  if (p.splits > 1) {
    launch_splitk_reduce(p.partial, p.splits, p.M, p.N, p.E, p.C16, p.ld16, p.rowsum_partial, p.rowsum, st);
    HULC2_CHECK_LAUNCH();
  }
  return HULC2_OK;
}

template <int BN>
int launch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TmaParams& p, int smem, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(ta, tb, p, smem, st);
  if (!a_mn && b_mn) return launch<BN, false, true>(ta, tb, p, smem, st);
  if (a_mn && b_mn) return launch<BN, true, true>(ta, tb, p, smem, st);
  return launch<BN, true, false>(ta, tb, p, smem, st);
}

bool al(const void* p, uintptr_t a) { return ((uintptr_t)p & (a - 1)) == 0; }
bool gemm_cluster_splitk() {
  static int v = -1;
  // A/B switch, read once.  Default OFF -- measured (r02, 1 x B200, interleaved runs): 221 instead of 315 launches per step but
  // 6.66 vs 6.41 ms: a cluster of 8 CTAs pushing 64 KB each through DSMEM behind two cluster barriers costs 8-11 us per launch
  // (128 x 2048 x 2048: 17 -> 28 us), more than the partial round trip through L2 plus the reduce launch it replaces.
  if (v < 0) { const char* e = getenv("HULC2_GEMM_CLUSTER_SPLITK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
bool gemm_small_k_bn128() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HULC2_GEMM_SMALLK_BN128"); v = (e && e[0] == '1') ? 1 : 0; }   // A/B switch (default off), read once
  return v == 1;
}

}  // namespace

// Returns HULC2_ENOTIMPL when the problem does not fit the TMA path (the caller then uses the gather kernel).
int hulc2_gemm_tma_impl(const hulc2_gemm_args* a, cudaStream_t st) {
  if (!a->A16 || !a->B16 || a->a_inner > 0 || a->K <= 0 || a->M <= 0 || a->N <= 0) return HULC2_ENOTIMPL;
  if (!hulc2_device_supports_tcgen05()) return HULC2_ENOTIMPL;
  // operand orientation: K-major when the contraction index is the unit-stride axis, MN-major when the row index is
  bool a_mn, b_mn;
  long long a_ld, b_ld;
  if (a->a_ks == 1) { a_mn = false; a_ld = a->a_rs; } else if (a->a_rs == 1) { a_mn = true; a_ld = a->a_ks; } else return HULC2_ENOTIMPL;
  if (a->b_ks == 1) { b_mn = false; b_ld = a->b_rs; } else if (a->b_rs == 1) { b_mn = true; b_ld = a->b_ks; } else return HULC2_ENOTIMPL;
  if (!al(a->A16, 16) || !al(a->B16, 16) || a_ld % 8 || b_ld % 8 || a_ld <= 0 || b_ld <= 0) return HULC2_ENOTIMPL;

  TmaParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  fill_epilogue(p.E, a->C, a->ldc);
  p.E.c_inner = a->c_inner; p.E.cs_outer = a->c_rs_outer; p.E.cs_inner = a->c_rs_inner;
  p.E.bias = a->bias; p.E.add = a->add; p.E.ld_add = a->ld_add; p.E.mask = a->mask; p.E.ld_mask = a->ld_mask;
  p.E.keep = a->keep; p.E.ld_keep = a->ld_keep; p.E.keep_scale = a->keep_scale;
  p.E.relu = a->relu; p.E.accumulate = a->accumulate; p.E.alpha = a->alpha;
  p.C16 = (__nv_bfloat16*)a->C16; p.ld16 = a->ld16;
  p.ktiles = hulc2_cdiv(a->K, BK);

  const int mt = hulc2_cdiv(a->M, BM);
  int BN = a->N <= 64 ? 64 : (a->N <= 128 ? 128 : 256);
  if (BN == 256 && (long long)mt * hulc2_cdiv(a->N, 256) < 148) BN = 128;   // more, smaller tiles when the grid is short
  if (BN == 256 && a->rowsum) BN = 128;
  // small K with a big output: the epilogue is the kernel; 128-column tiles put four CTAs (16 epilogue warps) on an SM instead of two
  if (BN == 256 && p.ktiles <= 4 && gemm_small_k_bn128()) BN = 128;                                      // the row-sum accumulator doubles the TMEM allocation
  if (BN == 128 && a->N > 64 && (long long)mt * hulc2_cdiv(a->N, 128) < 74 && p.ktiles <= 8) BN = 64;
  const long long tiles = (long long)mt * hulc2_cdiv(a->N, BN);

  // split-K: skinny outputs with a long contraction get ~2 waves of CTAs, >= 4 k-tiles each
  p.splits = 1; p.kt_per_split = p.ktiles; p.partial = nullptr;
  // (the reduce kernel applies the full epilogue and writes the bf16 mirror, so fused-epilogue layers split too)
  const bool simple = a->c_inner == 0;
  const bool cluster_ok = gemm_cluster_splitk() && BN <= 128;       // (split problems never take 256-column tiles: tiles < 100)
  if (simple && tiles < 100 && p.ktiles >= 8 && (a->workspace || cluster_ok)) {
    int want = (int)((296 + tiles - 1) / tiles);
    int maxs = p.ktiles / 4;
    int s = want < maxs ? want : maxs;
    if (cluster_ok) {
      if (s > 8) s = 8;                              // portable cluster size
    } else {
      const long long per = ((long long)a->M * a->N + (a->rowsum ? a->M : 0)) * (long long)sizeof(float);
      while (s > 1 && (long long)s * per > a->workspace_bytes) --s;
    }
    if (s > 1) {
      p.kt_per_split = hulc2_cdiv(p.ktiles, s);
      p.splits = hulc2_cdiv(p.ktiles, p.kt_per_split);
      if (cluster_ok) {
        p.cluster_splitk = 1;
      } else {
        p.partial = (float*)a->workspace;
        p.rowsum_partial = a->rowsum ? p.partial + (long long)p.splits * a->M * a->N : nullptr;
      }
    }
  }
  const int stage_bytes = (int)A_BYTES + BN * 128;
  const long long ctas = tiles * p.splits;
  const int budget = (ctas > 148 && stage_bytes * 3 <= 110 * 1024) ? 110 * 1024 : 200 * 1024;   // 2 CTAs/SM when they exist
  int stages = budget / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > p.kt_per_split) stages = p.kt_per_split;
  if (stages < 1) stages = 1;
  if (p.cluster_splitk) {
    // RED[splits][ceil(128 / splits)][BN + 4] fp32 lives in the ring: give the ring the stages it needs for that
    const int rp = (BM + p.splits - 1) / p.splits;
    const long long red = (long long)p.splits * rp * (BN + 4) * 4;
    while ((long long)stages * stage_bytes < red) ++stages;
  }
  p.stages = stages;
  const int smem = stages * stage_bytes + 1024 + (a->rowsum ? 2048 : 0);
  p.rowsum = a->rowsum;

  // 16-byte epilogue accesses
  bool vec = al(a->C, 16) && a->ldc % 4 == 0;
  if (a->bias) vec = vec && al(a->bias, 16);
  if (a->add) vec = vec && al(a->add, 16) && a->ld_add % 4 == 0;
  if (a->mask) vec = vec && al(a->mask, 16) && a->ld_mask % 4 == 0;
  if (a->keep) vec = vec && al(a->keep, 4) && a->ld_keep % 4 == 0;
  if (a->C16) vec = vec && al(a->C16, 8) && a->ld16 % 4 == 0;
  if (p.splits > 1 && !p.cluster_splitk) vec = al(p.partial, 16) && a->N % 4 == 0;
  p.vec = vec ? 1 : 0;
  bool vec8 = vec && p.splits == 1 && al(a->C, 32) && a->ldc % 8 == 0 && a->c_inner == 0;
  if (a->bias) vec8 = vec8 && al(a->bias, 32);
  if (a->C16) vec8 = vec8 && al(a->C16, 16) && a->ld16 % 8 == 0;
  p.vec8 = vec8 ? 1 : 0;

  CUtensorMap ta, tb;
  bool ok = a_mn ? encode_map(&ta, a->A16, a->M, a->K, a_ld, 64) : encode_map(&ta, a->A16, a->K, a->M, a_ld, BM);
  ok = ok && (b_mn ? encode_map(&tb, a->B16, a->N, a->K, b_ld, 64) : encode_map(&tb, a->B16, a->K, a->N, b_ld, BN));
  if (!ok) return HULC2_ENOTIMPL;

  switch (BN) {
    case 64: return launch_major<64>(a_mn, b_mn, ta, tb, p, smem, st);
    case 128: return launch_major<128>(a_mn, b_mn, ta, tb, p, smem, st);
    default: return launch_major<256>(a_mn, b_mn, ta, tb, p, smem, st);
  }
}

extern "C" unsigned long long hulc2_tma_gemm_count(void) { return g_tma_gemms; }
