// bf16 tensor-core contractions for sm_100a: tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) with fp32
// accumulators in TMEM, operands staged in shared memory in the canonical K-major SWIZZLE_128B layout.
//
// The same kernel serves nn.Linear fwd/dgrad/wgrad and the implicit-GEMM convolutions (fwd = im2col A,
// wgrad = im2col^T B with split-K, dgrad = strided gather A), sharing operand addressing and the fused
// epilogue with the fp32 backend (gemm_common.cuh).  Operands are fp32 in HBM (master activations /
// weights): the producer threads gather them (im2col / transposes are pure address arithmetic), round to
// bf16 and write 16-byte chunks into the swizzled tile; `fence.proxy.async` hands the tile to the tensor
// core; one elected thread issues the MMAs and commits them to an mbarrier that recycles the stage.
// Global loads of tile t+1 are in flight while tile t is multiplied; 2-5 CTAs co-reside per SM
// (TMEM columns: BN per CTA) to cover load latency.  The accumulator tile is read back with tcgen05.ld
// (32 lanes x 16 columns per instruction) and goes through the shared epilogue.
#include <cuda_bf16.h>

#include "gemm_common.cuh"

using namespace hulc2;

namespace {

constexpr int BM = 128;
constexpr int BKE = 64;     // bf16 elements per k-tile = one 128-byte swizzle row
constexpr int NT = 256;
constexpr int STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) __trap();  // never hang the GPU on a protocol bug
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), version 1.  Atoms are 8 rows x 128 bytes = 1024 B (SBO).
//  K-major : rows = M/N index, 128-byte row = 64 k; LBO unused.
//  MN-major: rows = k index, 128-byte row = 64 M/N elements; LBO = stride between 64-wide M/N blocks (8192 B here).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), a_major @15, b_major @16
// (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Gathers one operand tile (fp32 in HBM) as 16-byte bf16 chunks into the canonical SWIZZLE_128B layout.
//  K-major  (KMAJ): tile = [ROWS mn-rows][64 k]; a thread owns chunk column (tid & 7) of rows (tid >> 3) + 32 i;
//                   8 consecutive k are contiguous in memory (row-major activations / weights, im2col, dgrad gather).
//  MN-major (!KMAJ): tile = ROWS/64 blocks of [64 k-rows][64 mn]; a thread owns mn-chunk (tid & 7) of k-rows
//                   (tid >> 3) + 32 j in every block; 8 consecutive mn are contiguous in memory (transposed
//                   activations for weight gradients, W[k][n] for input gradients, im2col^T).
// Either way a warp reads 4 rows x 256 contiguous bytes and writes 4 x 128-byte swizzled rows (conflict free).
template <int ROWS, int MODE, bool KMAJ>
struct TileGather {
  static constexpr int BLOCKS = (ROWS + 63) / 64;
  static constexpr int CH = KMAJ ? ROWS * 8 / NT : BLOCKS * 2;
  static constexpr int NR = KMAJ ? CH : BLOCKS;
  static constexpr uint32_t BYTES = (ROWS > 64 ? ROWS : 64) * 128;
  uint4 v[CH];
  long long roff[NR];     // K-major: row offsets; MN-major: offset of the first mn element of this thread's chunk, per block
  int rdec[KMAJ ? NR : 1][3];
  int rvalid[NR];         // K-major: row valid (0/1); MN-major: number of valid mn elements in the chunk (0..8)
  int r0, c0, mn0;

  __device__ __forceinline__ void init(const Operand& o, int row_base, int nrows_total) {
    const int tid = threadIdx.x;
    c0 = tid & 7; r0 = tid >> 3;
    if (KMAJ) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        int r = row_base + r0 + 32 * i;
        rvalid[i] = r < nrows_total;
        int rr = rvalid[i] ? r : 0;
        roff[i] = 0;
        if (MODE == OP_DGRAD) {
          dgrad_row_decode(o.g, rr, rdec[i][0], rdec[i][1], rdec[i][2]);
        } else {
          roff[i] = row_off<MODE>(o, rr);
        }
      }
    } else {
      mn0 = row_base;
#pragma unroll
      for (int b = 0; b < BLOCKS; ++b) {
        int mn = row_base + b * 64 + c0 * 8;
        int left = nrows_total - mn;
        rvalid[b] = (b * 64 + c0 * 8 < ROWS) ? (left >= 8 ? 8 : (left > 0 ? left : 0)) : 0;
        roff[b] = rvalid[b] > 0 ? row_off<MODE>(o, mn) : 0;
      }
    }
  }

  __device__ __forceinline__ void load(const Operand& o, int k0, int kend, bool vec) {
    if (KMAJ) {
      const int kc = k0 + c0 * 8;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        if (rvalid[i] && kc < kend) {
          if (MODE == OP_DGRAD) {
            // k = (a, b, co): the 8 k of a chunk share the tap (a, b) because Cout % 8 == 0 (checked on the host)
            int ta, tb, co;
            dgrad_k_decode(o.g, kc, ta, tb, co);
            long long off;
            if (dgrad_src(o.g, rdec[i][0], rdec[i][1], rdec[i][2], ta, tb, off)) {
              const float* src = o.p + off + co;
              float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
              f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
            }
          } else if (vec && kc + 8 <= kend) {
            const float* src = o.p + roff[i] + col_off<MODE>(o, kc);
            float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) if (kc + e < kend) f[e] = __ldg(o.p + roff[i] + col_off<MODE>(o, kc + e));
          }
        }
        v[i] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kg = k0 + r0 + 32 * j;
        const bool kv = kg < kend;
        const long long koff = kv ? col_off<MODE>(o, kg) : 0;
#pragma unroll
        for (int b = 0; b < BLOCKS; ++b) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
          if (kv && rvalid[b] > 0) {
            if (vec && rvalid[b] == 8) {
              const float* src = o.p + roff[b] + koff;
              float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
              f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
            } else {
              const int mn = mn0 + b * 64 + c0 * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) if (e < rvalid[b]) f[e] = __ldg(o.p + row_off<MODE>(o, mn + e) + koff);
            }
          }
          v[b * 2 + j] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
        }
      }
    }
  }

  // tile base is 1024-byte aligned; 128-byte row r, 16-byte chunk c -> (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16)
  __device__ __forceinline__ void store(uint32_t tile_base) const {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      int r = KMAJ ? r0 + 32 * i : r0 + 32 * (i & 1);
      uint32_t blk = KMAJ ? 0u : (uint32_t)(i >> 1) * 8192u;
      uint32_t addr = tile_base + blk + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c0 ^ (r & 7)) << 4));
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z), "r"(v[i].w) : "memory");
    }
  }
};

struct Bf16Params {
  GemmParams g;
  int vecA, vecB;
};

template <int BN, int AMODE, bool AKF, int BMODE, bool BKF>
__global__ void __launch_bounds__(NT, 2) gemm_bf16_kernel(const Bf16Params bp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ uint32_t tmem_base_slot;
  const GemmParams& p = bp.g;
  constexpr uint32_t A_BYTES = TileGather<BM, AMODE, AKF>::BYTES, B_BYTES = TileGather<BN, BMODE, BKF>::BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const int ntiles = (kend - kbeg + BKE - 1) / BKE;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_slot), BN);
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&mma_done[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  TileGather<BM, AMODE, AKF> ga;
  TileGather<BN, BMODE, BKF> gb;
  ga.init(p.A, m0, p.M);
  gb.init(p.B, n0, p.N);
  constexpr uint32_t IDESC = make_idesc(BN, !AKF, !BKF);

  if (ntiles > 0) {
    ga.load(p.A, kbeg, kend, bp.vecA);
    gb.load(p.B, kbeg, kend, bp.vecB);
  }
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % STAGES;
    const uint32_t a_tile = tiles + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
    if (t >= STAGES) mbar_wait(smem_u32(&mma_done[s]), (uint32_t)((t / STAGES - 1) & 1));  // MMAs that read this stage retired
    ga.store(a_tile);
    gb.store(b_tile);
    fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
    if (t + 1 < ntiles) {                // next tile's global loads are in flight during the MMAs
      ga.load(p.A, kbeg + (t + 1) * BKE, kend, bp.vecA);
      gb.load(p.B, kbeg + (t + 1) * BKE, kend, bp.vecB);
    }
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // UMMA_K = 16: K-major advances 32 bytes inside the swizzle atom; MN-major advances 16 k-rows = 2048 bytes
      const uint64_t ad = make_desc(a_tile, AKF ? 0u : 8192u), bd = make_desc(b_tile, BKF ? 0u : 8192u);
      constexpr uint64_t A_STEP = (AKF ? 32u : 2048u) >> 4, B_STEP = (BKF ? 32u : 2048u) >> 4;
#pragma unroll
      for (int k = 0; k < BKE / 16; ++k)
        umma_bf16(tmem_d, ad + A_STEP * k, bd + B_STEP * k, IDESC, (t > 0 || k > 0) ? 1u : 0u);
      umma_commit(smem_u32(&mma_done[s]));
    }
  }
  if (ntiles > 0) {
    const int last = ntiles - 1;
    mbar_wait(smem_u32(&mma_done[last % STAGES]), (uint32_t)((last / STAGES) & 1));
  }
  tc_fence_after();

  // epilogue: warp w reads TMEM lanes 32*(w%4).., columns [(w/4)*BN/2, +BN/2)
  const Epilogue& E = p.E;
  const int q = warp & 3, half = warp >> 2;
  constexpr int COLS = BN / 2;
  const int m = m0 + q * 32 + lane;
  const bool mvalid = m < p.M;
  const int mp = mvalid ? phys_row(E, m) : 0;
  const long long crow = mvalid ? c_row_off(E, mp) : 0;
#pragma unroll 1
  for (int c = 0; c < COLS; c += 16) {
    float acc[16];
    if (ntiles > 0) {
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS + c), acc);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    }
    if (mvalid) {
      if (p.splits > 1) {
        float* prow = p.partial + ((long long)blockIdx.z * p.M + m) * p.N;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int n = n0 + half * COLS + c + j;
          if (n < p.N) prow[n] = acc[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int n = n0 + half * COLS + c + j;
          if (n < p.N) E.C[crow + n] = apply_epilogue(E, acc[j], mp, n, crow);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
}

template <int BN, int AMODE, bool AKF, int BMODE, bool BKF>
int launch_bn(const Bf16Params& bp, cudaStream_t st) {
  const GemmParams& p = bp.g;
  auto kern = gemm_bf16_kernel<BN, AMODE, AKF, BMODE, BKF>;
  const int smem = STAGES * (int)(TileGather<BM, AMODE, AKF>::BYTES + TileGather<BN, BMODE, BKF>::BYTES) + 1024;
  static bool configured = false;  // per instantiation
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      hulc2_set_error("gemm_bf16: cannot raise dynamic shared memory limit");
      return HULC2_ELAUNCH;
    }
    configured = true;
  }
  dim3 grid(hulc2_cdiv(p.M, BM), hulc2_cdiv(p.N, BN), p.splits);
  kern<<<grid, NT, smem, st>>>(bp);
  HULC2_CHECK_LAUNCH();
  if (p.splits > 1) {
    long long total = (long long)p.M * p.N;
    launch_splitk_reduce(p.partial, p.splits, p.M, p.N, p.E, nullptr, 0, nullptr, nullptr, st);
    HULC2_CHECK_LAUNCH();
  }
  return HULC2_OK;
}

int pick_bn(const GemmParams& p) {
  if (p.N <= 32) return 32;
  if (p.N <= 64) return 64;
  if (p.N <= 128) return 128;
  int bn = 256;
  while (bn > 32 && (long long)hulc2_cdiv(p.M, BM) * hulc2_cdiv(p.N, bn) * p.splits < 148) bn >>= 1;
  return bn;
}

template <int AMODE, bool AKF, int BMODE, bool BKF>
int launch_modes(const Bf16Params& bp, cudaStream_t st) {
  switch (pick_bn(bp.g)) {
    case 32: return launch_bn<32, AMODE, AKF, BMODE, BKF>(bp, st);
    case 64: return launch_bn<64, AMODE, AKF, BMODE, BKF>(bp, st);
    case 128: return launch_bn<128, AMODE, AKF, BMODE, BKF>(bp, st);
    default: return launch_bn<256, AMODE, AKF, BMODE, BKF>(bp, st);
  }
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// can 8 consecutive k of a K-fast operand be fetched as two aligned float4?
bool vec_ok(const Operand& o, int mode) {
  if (!aligned16(o.p)) return false;
  if (mode == OP_DENSE) {
    if (o.ks == 1) {   // K-major: 8 consecutive k
      if (o.r_inner > 0) return (o.rs_outer % 4 == 0) && (o.rs_inner % 4 == 0);
      return o.rs % 4 == 0;
    }
    return o.rs == 1 && o.r_inner == 0 && o.ks % 4 == 0;   // MN-major: 8 consecutive rows
  }
  const ConvGeom& g = o.g;
  if (mode == OP_IM2COL || mode == OP_IM2COL_T) {
    if (g.nhwc) return g.C % 8 == 0;
    return g.KW == 8 && g.W % 4 == 0 && g.stride % 4 == 0 && (g.H * g.W) % 4 == 0;
  }
  if (mode == OP_DGRAD) return g.Cout % 8 == 0;
  if (mode == OP_DGRAD_W) return g.C % 8 == 0;
  return false;
}

int dispatch(GemmParams& p, int amode, int bmode, cudaStream_t st) {
  if (!hulc2_device_supports_tcgen05()) { hulc2_set_error("bf16 path needs an sm_100 device (tcgen05)"); return HULC2_ENOTIMPL; }
  Bf16Params bp;
  bp.g = p;
  bool akf = (amode == OP_DENSE) ? (p.A.ks == 1) : true;
  bool bkf = (bmode == OP_DENSE) ? (p.B.ks == 1) : false;
  if (amode == OP_DGRAD && (p.A.g.Cout % 8 != 0 || ((uintptr_t)p.A.p & 15))) { hulc2_set_error("bf16 conv dgrad needs Cout % 8 == 0 and 16-byte aligned dy"); return HULC2_EINVAL; }
  bp.vecA = vec_ok(p.A, amode);
  bp.vecB = vec_ok(p.B, bmode);
  if (amode == OP_DENSE && bmode == OP_DENSE) {
    if (akf && bkf) return launch_modes<OP_DENSE, true, OP_DENSE, true>(bp, st);
    if (akf && !bkf) return launch_modes<OP_DENSE, true, OP_DENSE, false>(bp, st);
    if (!akf && !bkf) return launch_modes<OP_DENSE, false, OP_DENSE, false>(bp, st);
    return launch_modes<OP_DENSE, false, OP_DENSE, true>(bp, st);
  }
  if (amode == OP_IM2COL && bmode == OP_DENSE && bkf) return launch_modes<OP_IM2COL, true, OP_DENSE, true>(bp, st);
  if (amode == OP_DENSE && !akf && bmode == OP_IM2COL_T) return launch_modes<OP_DENSE, false, OP_IM2COL_T, false>(bp, st);
  if (amode == OP_DGRAD && bmode == OP_DGRAD_W) return launch_modes<OP_DGRAD, true, OP_DGRAD_W, false>(bp, st);
  hulc2_set_error("gemm_bf16: unsupported operand mode combination");
  return HULC2_EINVAL;
}

}  // namespace

int hulc2_gemm_tma_impl(const hulc2_gemm_args* a, cudaStream_t st);

int hulc2_gemm_bf16_impl(const hulc2_gemm_args* a, cudaStream_t st) {
  if (a && (a->M == 0 || a->N == 0)) return HULC2_OK;
  if (a && a->A16 && a->B16) {   // bf16 operand mirrors: TMA-fed kernel when the layout allows it
    int e = hulc2_gemm_tma_impl(a, st);
    if (e != HULC2_ENOTIMPL) return e;
  }
  if (a && (!a->A || !a->B)) { hulc2_set_error("gemm: bf16-only operands (A/B NULL) need a TMA-compatible layout: 16-byte aligned base, row stride % 8 == 0, K > 0"); return HULC2_EINVAL; }
  if (a && a->C16) { hulc2_set_error("gemm: a bf16 output mirror (C16) needs the TMA path (aligned bf16 operand mirrors)"); return HULC2_EINVAL; }
  if (a && a->rowsum) { hulc2_set_error("gemm: rowsum needs the TMA path (aligned bf16 operand mirrors)"); return HULC2_EINVAL; }
  GemmParams p;
  if (!dense_params(a, p)) return HULC2_EINVAL;
  if (a->M == 0 || a->N == 0) return HULC2_OK;
  long long out_ctas = (long long)hulc2_cdiv(a->M, BM) * hulc2_cdiv(a->N, 128);
  if (simple_epilogue(a) && out_ctas < 74) plan_splitk(p, out_ctas, 296, 256, BKE, a->workspace, a->workspace_bytes);
  else plan_splitk(p, 1, 1, 1 << 30, BKE, nullptr, 0);
  return dispatch(p, OP_DENSE, OP_DENSE, st);
}
int hulc2_conv2d_fwd_bf16_impl(const hulc2_conv_args* a, cudaStream_t st) {
  GemmParams p;
  conv_fwd_params(a, p);
  if (p.M == 0) return HULC2_OK;
  plan_splitk(p, 1, 1, 1 << 30, BKE, nullptr, 0);
  return dispatch(p, OP_IM2COL, OP_DENSE, st);
}
int hulc2_conv2d_wgrad_bf16_impl(const hulc2_conv_args* a, cudaStream_t st) {
  GemmParams p;
  conv_wgrad_params(a, p);
  if (p.K == 0) return HULC2_OK;
  long long out_ctas = (long long)hulc2_cdiv(p.M, BM) * hulc2_cdiv(p.N, 256);
  plan_splitk(p, out_ctas, 592, 1024, BKE, a->workspace, a->workspace_bytes);
  return dispatch(p, OP_DENSE, OP_IM2COL_T, st);
}
int hulc2_conv2d_dgrad_bf16_impl(const hulc2_conv_args* a, cudaStream_t st) {
  for (int ph = 0; ph < a->stride; ++ph)
    for (int pw = 0; pw < a->stride; ++pw) {
      GemmParams p;
      if (!conv_dgrad_params(a, ph, pw, p)) continue;
      plan_splitk(p, 1, 1, 1 << 30, BKE, nullptr, 0);
      if (int e = dispatch(p, OP_DGRAD, OP_DGRAD_W, st)) return e;
    }
  return HULC2_OK;
}
