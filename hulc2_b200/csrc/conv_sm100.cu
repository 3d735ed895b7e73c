// bf16 convolution trunk for sm_100a: persistent, warp-specialised implicit GEMMs on tcgen05/TMEM.
//
// Activations are bf16 NHWC in HBM (half the bytes of the fp32 path, and every im2col row is made of
// contiguous >= 64-byte segments), so operand tiles are gathered with 16-byte cp.async (LDGSTS) straight into
// the canonical SWIZZLE_128B layout -- no register staging, several stages in flight per SM.
//
//   conv_igemm_kernel (forward and input-gradient):  D[pixels, N] = A[pixels, K] . W[N, K]^T
//     * one CTA per SM, static round-robin over 128-pixel tiles; the whole packed weight matrix (<= 72 KB) is
//       resident in shared memory for the lifetime of the CTA, only A streams;
//     * warps 5-8 produce A stages (6-deep ring, cp.async + per-thread wait_group/fence.proxy.async/arrive),
//       warp 4 issues tcgen05.mma (128 x N x 16) and commits to the ring's empty barriers,
//       warps 0-3 drain the double-buffered TMEM accumulator (tcgen05.ld), apply bias+ReLU (forward) or the
//       ReLU mask of the producing layer (input gradient) and store bf16 rows;
//     * the input gradient of a stride-s conv runs as s*s parity classes (only taps that reach a class are
//       contracted), all classes of a layer in ONE launch; out-of-range taps are zero-filled by cp.async.
//   conv_wgrad_kernel:  dW^T[kconv, Cout] = sum_pixels im2col[pixel, kconv] * dZ[pixel, Cout]
//     * both operands are MN-major (the contraction index = pixel is the slow axis of both tensors);
//       each CTA owns a contiguous pixel range and keeps ALL of dW^T in TMEM (<= 5 x 64 columns);
//     * a constant block of ones appended to the im2col operand yields the bias gradient for free;
//     * per-CTA partials go to the workspace, a second kernel reduces them and scatters into OIHW order.
//
// Layer 1 reads "packed frames": fp32 NCHW images re-tiled once per step by pack_frames_kernel into bf16
// [F, H/4, W/4, 16*C] (space-to-depth by the stride), which turns the k8/s4 conv into a k2/s1 conv over 16*C
// channels whose im2col rows are KH contiguous segments of 2*16*C elements.
// Reference: hulc2/models/perceptual_encoders/vision_network.py:38-48, vision_network_gripper.py:11-26.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_halo.cuh"
#include "../../include/hulc2_b200.h"

using namespace sm100;

namespace {

constexpr int TILE_M = 128;
constexpr int KT = 64;                       // k per stage (one 128-byte swizzle row)
constexpr int MAX_STAGES = 12;               // ring depth is chosen at launch from the shared memory left after the weights
// a stage is published once `lag` younger cp.async groups have been issued by the same thread (lag = stages - 2)
constexpr uint32_t A_STAGE = TILE_M * 128;   // 16 KB
// warp roles: 0-7 epilogue (TMEM lanes 32*(w%4).., column half w/4), 8 MMA issuer, 9-16 producers.
// Profiling the first version (4 producer warps, one per SM sub-partition) showed the producers ~80 % busy issuing
// address arithmetic at ~0.2 IPC per warp with every memory pipe < 30 % utilised: the gather is instruction-latency
// bound, so it is spread over two warps per sub-partition and the per-copy instruction count is kept minimal.
constexpr int EPI_WARPS = 8, PROD_WARPS = 8;
constexpr int N_EPI = EPI_WARPS * 32, N_PROD = PROD_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int NT = N_EPI + 32 + N_PROD;      // 544 threads
constexpr int WG_MMA_WARP2 = EPI_WARPS + 1 + PROD_WARPS;  // weight-gradient kernel: second MMA issuer (warp 17)
constexpr int WG_MAX_ISS = 5;                 // weight-gradient kernel: MMA issuer warps = MMA_WARP and WG_MMA_WARP2 .. + 3
constexpr int NT_WG = NT + 32 * (WG_MAX_ISS - 1);   // 672 threads
constexpr int MAX_CLS = 4;
constexpr int MAX_TAB = 160;

// cp.async.wait_group takes an immediate: dispatch on the (warp-uniform) runtime lag
__device__ __forceinline__ void cp_async_wait_dyn(uint32_t n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    case 7: cp_async_wait<7>(); break;
    case 8: cp_async_wait<8>(); break;
    case 9: cp_async_wait<9>(); break;
    default: cp_async_wait<10>(); break;
  }
}

// exact unsigned division by a runtime constant (Granlund-Montgomery round-up method), n < 2^31
struct FastDiv {
  uint32_t mul, sh1, sh2, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l > 1 ? 1 : l;
  f.sh2 = l > 0 ? l - 1 : 0;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

struct ConvClass {
  int M;                 // rows (pixels) of this class
  FastDiv dHW, dW;       // row r -> (f, i, j): f = r / (PH*PW), i = rem / PW, j = rem % PW
  int K;                 // contraction length (multiple of 64)
  int w_off;             // byte offset of this class's packed weights [N][K]
  int tab_off;           // first chunk-table entry of this class
  int sF, sI, sJ;        // source base of row (f,i,j) in 16-byte units
  int oS, oPh, oPw;      // output pixel of row (f,i,j): (f*oH + i*oS + oPh)*oW + j*oS + oPw
};

struct ConvParams {
  const uint8_t* x;      // bf16 source tensor
  const uint8_t* w;      // packed bf16 weights, all classes
  const float* bias;     // [N] or null
  const uint8_t* mask;   // bf16, same shape as the output: keep where > 0 (or null)
  uint8_t* y;            // bf16 output [*, N]
  int ncls, ntiles, N, relu;  // ntiles = (max tiles of a class) << cls_shift: tile vt -> class vt & (ncls-1), tile-in-class vt >> cls_shift
  int ngroups, grp_shift; // work items of (1 << grp_shift) consecutive tiles: grp_shift = cls_shift, or 0 with HULC2_DGRAD_SPREAD=1 (old order)
  int cls_shift;         // classes are interleaved so the s*s parity classes of one image region run back to back (dZ stays in L2)
  int VH, VW;            // a tap (a, b) of row (f,i,j) is valid iff 0 <= i-a < VH and 0 <= j-b < VW
  int oH, oW;
  int w_bytes;           // total packed weight bytes
  int ntab;
  int stages, lag;       // A ring depth and publish lag (lag <= stages - 2)
  int kts;               // k-tiles (64 k each) per ring slot: 1, or the whole K of a tile when that fits (one barrier round trip,
                         // wait_group and proxy fence per tile instead of per k-tile: the producers are bookkeeping-bound)
  ConvClass cls[MAX_CLS];
  int2 table[MAX_TAB];   // per 16-byte chunk of K: {delta in 16-byte units, (a << 16) | b}
};

// DGRAD = false: forward conv (every tap of a valid row is in range, output rows are dense: pixel = row)
// DGRAD = true : input gradient (per-tap validity -> zero fill, strided output pixels, optional ReLU mask)
template <int BN, bool DGRAD>
__global__ void __launch_bounds__(NT, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int2 tab_s[MAX_TAB];
  __shared__ float bias_s[BN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t STAGES = (uint32_t)p.stages, LAG = (uint32_t)p.lag;
  const uint32_t KTS = (uint32_t)p.kts, SLOT = KTS * A_STAGE;
  const uint32_t a_smem = base;                              // STAGES x KTS x 16 KB
  const uint32_t w_smem = base + STAGES * SLOT;              // per class: K/64 tiles of [BN rows][128 B]
  constexpr uint32_t W_TILE = BN * 128;
  constexpr uint32_t TCOLS = 2 * BN;                         // double-buffered accumulator

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
  if (tid == 32) {
    for (uint32_t s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), N_PROD); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), N_EPI); }
    mbar_fence_init();
  }
  for (int i = tid; i < p.ntab; i += NT) tab_s[i] = p.table[i];
  if (tid < BN) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
  // resident weights: [N][K] bf16 per class -> K-major SWIZZLE_128B tiles
  for (int c = 0; c < p.ncls; ++c) {
    const int K = p.cls[c].K, cpr = K >> 3;                  // 16-byte chunks per weight row
    const uint8_t* src = p.w + p.cls[c].w_off;
    const uint32_t dst0 = w_smem + p.cls[c].w_off;
    for (int ch = tid; ch < BN * cpr; ch += NT) {
      const int n = ch / cpr, q = ch - n * cpr;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)n * K + q * 8) * 2));
      const uint32_t dst = dst0 + (uint32_t)(q >> 3) * W_TILE + swz128(n, q & 7);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  const int cls_mask = (1 << p.cls_shift) - 1;

  if (warp > MMA_WARP) {
    // ===================================================== producers (256 threads: 32 rows x 8 chunks per pass, 4 passes)
    const int t = tid - (N_EPI + 32);
    const int c8 = t & 7, r0 = t >> 3;
    const uint32_t dst_t = swz128(r0, c8);                   // + q * 4096 (32 rows = 4 atoms)
    // ring bookkeeping without runtime divisions: s/ph = stage being filled and its empty-barrier parity,
    // ps = oldest stage not yet published, inflight = committed-but-unpublished groups (<= LAG)
    uint32_t s = 0, ph = 1, ps = 0, inflight = 0;
    // work order of a CTA: its tile groups round-robin, and inside a group the s*s parity classes of the same image
    // region back to back (the classes read the same dZ pixels: the second..fourth pass hit in this SM's L1)
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x)
    for (int tile = grp << p.grp_shift, tend = tile + (1 << p.grp_shift); tile < tend; ++tile) {
      const ConvClass& cl = p.cls[DGRAD ? (tile & cls_mask) : 0];
      const int m0 = (DGRAD ? (tile >> p.cls_shift) : tile) * TILE_M;
      if (m0 >= cl.M) continue;
      const uint8_t* rptr[4];
      uint32_t rij[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t r = (uint32_t)(m0 + r0 + 32 * q);
        const uint32_t f = fdiv(r, cl.dHW), rem = r - f * cl.dHW.d;
        const uint32_t i = fdiv(rem, cl.dW), j = rem - i * cl.dW.d;
        const bool rv = r < (uint32_t)cl.M;
        rptr[q] = p.x + ((long long)(rv ? (int)(f * cl.sF + i * cl.sI + j * cl.sJ) : 0) << 4);
        rij[q] = rv ? ((i << 16) | j) : 0x7fff7fffu;
      }
      const int nkt = cl.K / KT;
      const int2* tab = tab_s + cl.tab_off + c8;
      for (int kt0 = 0; kt0 < nkt; kt0 += (int)KTS) {
        mbar_wait(smem_u32(&empty_bar[s]), ph);
        for (uint32_t kk = 0; kk < KTS; ++kk) {
        const int2 e = tab[(kt0 + (int)kk) * 8];
        const long long doff = (long long)e.x << 4;
        const uint32_t dst = a_smem + s * SLOT + kk * A_STAGE + dst_t;
        if (DGRAD) {
          const int ta = e.y >> 16, tb = e.y & 0xffff;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ii = (int)(rij[q] >> 16) - ta, jj = (int)(rij[q] & 0xffff) - tb;
            const bool ok = (unsigned)ii < (unsigned)p.VH && (unsigned)jj < (unsigned)p.VW;
            cp_async16_ca(dst + q * 4096, ok ? rptr[q] + doff : p.x, ok ? 16u : 0u);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) cp_async16_ca(dst + q * 4096, rptr[q] + doff, rij[q] == 0x7fff7fffu ? 0u : 16u);
        }
        }
        cp_async_commit();
        if (++s == STAGES) { s = 0; ph ^= 1; }
        if (inflight == LAG) {
          cp_async_wait_dyn(LAG);
          fence_proxy_async();
          mbar_arrive(smem_u32(&full_bar[ps]));
          if (++ps == STAGES) ps = 0;
        } else {
          ++inflight;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; inflight > 0; --inflight) {
      mbar_arrive(smem_u32(&full_bar[ps]));
      if (++ps == STAGES) ps = 0;
    }
  } else if (warp == MMA_WARP) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(TILE_M, BN, false, false);
      uint32_t s = 0, ph = 0, ti = 0;
      // work order of a CTA: its tile groups round-robin, and inside a group the s*s parity classes of the same image
    // region back to back (the classes read the same dZ pixels: the second..fourth pass hit in this SM's L1)
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x)
    for (int tile = grp << p.grp_shift, tend = tile + (1 << p.grp_shift); tile < tend; ++tile) {
        const ConvClass& cl = p.cls[DGRAD ? (tile & cls_mask) : 0];
        if ((DGRAD ? (tile >> p.cls_shift) : tile) * TILE_M >= cl.M) continue;
        const uint32_t buf = ti & 1;
        mbar_wait(smem_u32(&tempty_bar[buf]), ((ti >> 1) & 1) ^ 1);      // epilogue drained this accumulator
        tc_fence_after();
        const int nkt = cl.K / KT;
        for (int kt0 = 0; kt0 < nkt; kt0 += (int)KTS) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          for (uint32_t kk = 0; kk < KTS; ++kk) {
            const int kt = kt0 + (int)kk;
            const uint64_t ad = make_desc(a_smem + s * SLOT + kk * A_STAGE, 0), bd = make_desc(w_smem + cl.w_off + kt * W_TILE, 0);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k) umma_bf16(tmem_d + buf * BN, ad + 2 * k, bd + 2 * k, IDESC, (kt > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));                          // slot reusable once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(smem_u32(&tfull_bar[buf]));                          // accumulator complete
        ++ti;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue: warp w <-> TMEM lanes 32*(w%4).., columns (w/4)*BN/2..
    constexpr int HC = BN / 2;                                           // columns per thread
    const int lq = warp & 3, half = warp >> 2;
    uint32_t ti = 0;
    // work order of a CTA: its tile groups round-robin, and inside a group the s*s parity classes of the same image
    // region back to back (the classes read the same dZ pixels: the second..fourth pass hit in this SM's L1)
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x)
    for (int tile = grp << p.grp_shift, tend = tile + (1 << p.grp_shift); tile < tend; ++tile) {
      const ConvClass& cl = p.cls[DGRAD ? (tile & cls_mask) : 0];
      const int m0 = (DGRAD ? (tile >> p.cls_shift) : tile) * TILE_M;
      if (m0 >= cl.M) continue;
      const uint32_t buf = ti & 1;
      const uint32_t r = (uint32_t)(m0 + lq * 32 + lane);
      long long opix = r;
      if (DGRAD) {
        const uint32_t f = fdiv(r, cl.dHW), rem = r - f * cl.dHW.d;
        const uint32_t i = fdiv(rem, cl.dW), j = rem - i * cl.dW.d;
        opix = ((long long)f * p.oH + i * cl.oS + cl.oPh) * p.oW + j * cl.oS + cl.oPw;
      }
      const bool rv = r < (uint32_t)cl.M;
      uint4 mk[HC / 8];
      if (DGRAD && p.mask && rv) {
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) mk[c] = __ldg(reinterpret_cast<const uint4*>(p.mask + opix * (BN * 2) + half * HC * 2) + c);
      }
      mbar_wait_relaxed(smem_u32(&tfull_bar[buf]), (ti >> 1) & 1);
      tc_fence_after();
      uint32_t acc[HC];
#pragma unroll
      for (int c = 0; c < HC; c += 16)
        tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + buf * BN + half * HC + c, *reinterpret_cast<uint32_t(*)[16]>(&acc[c]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[buf]));
      if (rv) {
        uint8_t* out = p.y + opix * (BN * 2) + half * HC * 2;
#pragma unroll
        for (int c = 0; c < HC; c += 16) {                     // 16 channels = 32 bytes per 256-bit store
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float lo = __uint_as_float(acc[c + 2 * e]), hi = __uint_as_float(acc[c + 2 * e + 1]);
            if (!DGRAD) {
              lo += bias_s[half * HC + c + 2 * e]; hi += bias_s[half * HC + c + 2 * e + 1];
              if (p.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            }
            o[e] = pack_bf16x2(lo, hi);
          }
          if (DGRAD && p.mask) {
            const uint32_t mw[8] = {mk[c >> 3].x, mk[c >> 3].y, mk[c >> 3].z, mk[c >> 3].w,
                                    mk[(c >> 3) + 1].x, mk[(c >> 3) + 1].y, mk[(c >> 3) + 1].z, mk[(c >> 3) + 1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              // bf16 activations are >= 0 after ReLU: keep a half-word where the mask half-word is a positive number
              const uint32_t m = mw[e];
              const uint32_t keep = (((m & 0x7fffu) != 0 && !(m & 0x8000u)) ? 0x0000ffffu : 0u) |
                                    (((m & 0x7fff0000u) != 0 && !(m & 0x80000000u)) ? 0xffff0000u : 0u);
              o[e] &= keep;
            }
          }
          stg256(out + c * 2, o);
        }
      }
      ++ti;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TCOLS);
}

// ------------------------------------------------------------------------------------------------ wgrad
// pixels (contraction rows) per stage = template parameter KP: 32, or 64 (two per producer thread, which halves the per-stage
// bookkeeping -- barrier round trip, wait_group, proxy fence -- per byte moved) when four 64-pixel stages fit in shared memory
constexpr int WG_MAX_STAGES = 12;
constexpr int WG_MAX_BLK = 10;                // <= 5 M-tiles of 128 -> 320 TMEM columns

struct WgradParams {
  const uint8_t* x;      // bf16 NHWC input of the conv
  const uint8_t* dz;     // bf16 [P, Cout]
  float* partial;        // [grid][nblk*64][64] fp32
  int P;                 // pixels = F*OH*OW
  FastDiv dHW, dW;       // pixel -> (f, oh, ow)
  int sF, sI, sJ;        // im2col row base in 16-byte units
  int K;                 // kconv (multiple of 64); data blocks = K/64; ones block = K/64; nblk (even) >= K/64 + 1
  int nblk;
  int cout8;             // Cout / 8 (4 or 8)
  int nstages;           // ceil(P / WG_KP)
  int stages, lag;       // ring depth / publish lag (see conv_igemm_kernel)
  int delta[72];         // per 16-byte chunk of kconv: offset in 16-byte units
};

template <int WG_KP>
__global__ void __launch_bounds__(NT_WG, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr uint32_t WG_BLK = WG_KP * 128;      // one 64-wide M/N block of a stage: KP k-rows x 128 B
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)(p.nblk + 1) * WG_BLK;          // A blocks + the dZ block
  const uint32_t WG_STAGES = (uint32_t)p.stages, WG_LAG = (uint32_t)p.lag;
  const int nbd = p.K / 64;                                               // data blocks
  const int nmt = p.nblk / 2;
  // MMA issuers: ONE thread gets a 128 x 64 x 16 MMA with two MN-major operands through only every ~375 cycles (ncu r02: the
  // producers wait for free ring slots, tensor pipe 19-27 %), so the issue is spread over up to WG_MAX_ISS threads: one per
  // M-tile of dW^T, and -- when TMEM has room for a second accumulator set (conv1: 2 M-tiles) -- times two halves of a stage's
  // k-steps, each half accumulating into its own set (added in the dump).
  const int ksplit = (2 * nmt * 64 <= 512 && 2 * nmt <= WG_MAX_ISS) ? 2 : 1;
  const int ni_m = nmt < WG_MAX_ISS / ksplit ? nmt : WG_MAX_ISS / ksplit;
  const int niss = ni_m * ksplit;
  const uint32_t need_cols = (uint32_t)(ksplit * nmt * 64);
  const uint32_t tcols = need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u));

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), tcols);
  if (tid == 32) {
    // every MMA issuer (below) commits once per stage / at the end
    for (uint32_t s = 0; s < WG_STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), N_PROD); mbar_init(smem_u32(&empty_bar[s]), (uint32_t)niss); }
    mbar_init(smem_u32(&done_bar), (uint32_t)niss);
    mbar_fence_init();
  }
  // zero every stage, then fill the ones block (bf16 1.0 = 0x3F80) -- neither is touched by the producers
  for (uint32_t o = tid * 16; o < WG_STAGES * stage_bytes; o += NT_WG * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
  __syncthreads();
  for (uint32_t s = 0; s < WG_STAGES; ++s)
    for (uint32_t o = tid * 16; o < WG_BLK; o += NT_WG * 16)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + s * stage_bytes + nbd * WG_BLK + o), "r"(0x3F803F80u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  // contiguous stage range of this CTA (every CTA gets at least one stage: grid <= nstages)
  const int sbeg = (int)((long long)p.nstages * blockIdx.x / gridDim.x);
  const int send = (int)((long long)p.nstages * (blockIdx.x + 1) / gridDim.x);

  if (warp > MMA_WARP && warp < WG_MMA_WARP2) {
    // producers: 256 threads = 32 pixels x 8 chunks, WG_KP / 32 passes; each thread copies its chunk column of every block of its pixels
    const int t = tid - (N_EPI + 32);
    const int c8 = t & 7, g = t >> 3;
    const uint32_t drow = swz128(g, c8);
    int dl[9];
#pragma unroll
    for (int b = 0; b < 9; ++b) dl[b] = b < nbd ? p.delta[b * 8 + c8] : 0;
    const bool has_dz = c8 < p.cout8;
    uint32_t s = 0, ph = 1, ps = 0, inflight = 0;
    for (int st = sbeg; st < send; ++st) {
      const uint8_t* rp[WG_KP / 32];
      uint32_t pu[WG_KP / 32], nb[WG_KP / 32];
#pragma unroll
      for (int q = 0; q < WG_KP / 32; ++q) {                  // this thread's pixels g, g + 32, ..: rows 32 q + g of every block
        const long long pix = (long long)st * WG_KP + g + 32 * q;
        const bool ok = pix < p.P;
        pu[q] = ok ? (uint32_t)pix : 0u;
        const uint32_t f = fdiv(pu[q], p.dHW), rem = pu[q] - f * p.dHW.d;
        const uint32_t i = fdiv(rem, p.dW), j = rem - i * p.dW.d;
        rp[q] = p.x + ((long long)(int)(f * p.sF + i * p.sI + j * p.sJ) << 4);
        nb[q] = ok ? 16u : 0u;
      }
      mbar_wait(smem_u32(&empty_bar[s]), ph);
      const uint32_t sb = base + s * stage_bytes + drow;
#pragma unroll
      for (int q = 0; q < WG_KP / 32; ++q) {
#pragma unroll
        for (int b = 0; b < 9; ++b)
          if (b < nbd) cp_async16_ca(sb + b * WG_BLK + q * 4096, rp[q] + ((long long)dl[b] << 4), nb[q]);
        if (has_dz) cp_async16(sb + p.nblk * WG_BLK + q * 4096, p.dz + (((long long)pu[q] * p.cout8 + c8) << 4), nb[q]);
      }
      cp_async_commit();
      if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      if (inflight == WG_LAG) {
        cp_async_wait_dyn(WG_LAG);
        fence_proxy_async();
        mbar_arrive(smem_u32(&full_bar[ps]));
        if (++ps == WG_STAGES) ps = 0;
      } else {
        ++inflight;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; inflight > 0; --inflight) {
      mbar_arrive(smem_u32(&full_bar[ps]));
      if (++ps == WG_STAGES) ps = 0;
    }
  } else if (warp == MMA_WARP || warp >= WG_MMA_WARP2) {
    // issuer index: MMA_WARP -> 0, WG_MMA_WARP2 + j -> 1 + j;  issuer = (im, ik): M-tiles mt = im, im + ni_m, ..; k-steps ks = ik, ik + ksplit, ..
    const int me = warp == MMA_WARP ? 0 : 1 + (warp - WG_MMA_WARP2);
    if (lane == 0 && me < niss) {
      constexpr uint32_t IDESC = make_idesc(128, 64, true, true);
      const int im = me % ni_m, ik = me / ni_m;
      const uint32_t acc0 = tmem_d + (uint32_t)(ik * nmt * 64);
      uint32_t s = 0, ph = 0;
      for (int st = sbeg; st < send; ++st) {
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sb = base + s * stage_bytes;
        for (int ks = ik; ks < WG_KP / 16; ks += ksplit) {
          const uint64_t bd = make_desc(sb + p.nblk * WG_BLK + ks * 2048, WG_BLK);
          for (int mt = im; mt < nmt; mt += ni_m) {
            const uint64_t ad = make_desc(sb + mt * 2 * WG_BLK + ks * 2048, WG_BLK);
            umma_bf16(acc0 + mt * 64, ad, bd, IDESC, (st > sbeg || ks > ik) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        if (++s == WG_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(smem_u32(&done_bar));
    }
    __syncwarp();
  } else {
    // dump: warp w <-> TMEM lanes 32*(w%4).., columns (w/4)*32..+32 of every M-tile
    const int lq = warp & 3, half = warp >> 2;
    mbar_wait_relaxed(smem_u32(&done_bar), 0);
    tc_fence_after();
    float* out = p.partial + (size_t)blockIdx.x * p.nblk * 64 * 64;
    for (int mt = 0; mt < nmt; ++mt) {
      uint32_t acc[32];
#pragma unroll
      for (int c = 0; c < 32; c += 16)
        tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + mt * 64 + half * 32 + c, *reinterpret_cast<uint32_t(*)[16]>(&acc[c]));
      tmem_ld_wait();
      if (ksplit == 2) {                       // second accumulator set (the other half of every stage's k-steps)
        uint32_t acc2[32];
#pragma unroll
        for (int c = 0; c < 32; c += 16)
          tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + (nmt + mt) * 64 + half * 32 + c, *reinterpret_cast<uint32_t(*)[16]>(&acc2[c]));
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = __float_as_uint(__uint_as_float(acc[c]) + __uint_as_float(acc2[c]));
      }
      float4* o4 = reinterpret_cast<float4*>(out + (size_t)(mt * 128 + lq * 32 + lane) * 64 + half * 32);
#pragma unroll
      for (int c = 0; c < 8; ++c) o4[c] = make_float4(__uint_as_float(acc[4 * c]), __uint_as_float(acc[4 * c + 1]), __uint_as_float(acc[4 * c + 2]), __uint_as_float(acc[4 * c + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tcols);
}

// dw (OIHW fp32) and db from the per-CTA partials.  layout 0: kconv = (kh, kw, ci);
// layout 1 (packed frames): kconv = (dI, dJ, ci, a, b) -> kh = 4 dI + a, kw = 4 dJ + b of a [Cout, C/16, 4KH, 4KW] weight.
// `pitch`: rows per tap in the partials (C for the gather kernel's dense rows, 64 for the halo kernel's padded tap blocks).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int rows_per_part, int K, int Cout, int C, int KH,
                                    int KW, int layout, float* __restrict__ dw, float* __restrict__ db, int pitch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (k, co), co fastest; k == K is the ones row (bias)
  if (idx >= (K + 1) * Cout) return;
  const int k = idx / Cout, co = idx - k * Cout;
  const int row = (k / C) * pitch + k % C;                 // k == K -> (K / C) * pitch: the block of ones
  float s = 0.f;
  for (int g = 0; g < nparts; ++g) s += partial[((size_t)g * rows_per_part + row) * 64 + co];
  if (k == K) { if (db) db[co] = s; return; }
  if (!dw) return;
  int o;
  if (layout == 0) {
    const int ci = k % C, t = k / C, kw = t % KW, kh = t / KW;
    o = ((co * C + ci) * KH + kh) * KW + kw;
  } else {
    const int c16 = k % C, t = k / C, dJ = t % KW, dI = t / KW;
    const int ci = c16 >> 4, a = (c16 >> 2) & 3, b = c16 & 3;
    o = ((co * (C >> 4) + ci) * (4 * KH) + 4 * dI + a) * (4 * KW) + 4 * dJ + b;
  }
  dw[o] = s;
}

// ------------------------------------------------------------------------------------------------ packing
// fp32 NCHW frames -> bf16 [F, H/4, W/4, 16 C] with channel order (ci, a, b) = x[f, ci, 4I + a, 4J + b]
__global__ void pack_frames_kernel(const float* __restrict__ x, uint8_t* __restrict__ xs, int C, int H, int W, int H4, int W4) {
  extern __shared__ float rows[];                      // [C][4][W4*4]
  const int f = blockIdx.x / H4, I = blockIdx.x - f * H4;
  const int wu = W4 * 4;
  const int nq = C * 4 * W4;                           // float4 loads (when W % 4 == 0 and 16-byte aligned rows) else scalar
  const bool vec = (W & 3) == 0;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    const int j4 = q % W4, t = q / W4, a = t & 3, ci = t >> 2;
    const float* src = x + (((size_t)f * C + ci) * H + 4 * I + a) * W + 4 * j4;
    float4 v;
    if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
    else v = make_float4(src[0], src[1], src[2], src[3]);
    *reinterpret_cast<float4*>(&rows[(ci * 4 + a) * wu + 4 * j4]) = v;
  }
  __syncthreads();
  const int c16 = 16 * C, cpc = c16 / 8;               // 16-byte output chunks per cell
  uint4* out = reinterpret_cast<uint4*>(xs + ((size_t)f * H4 + I) * W4 * c16 * 2);
  for (int q = threadIdx.x; q < W4 * cpc; q += blockDim.x) {
    const int J = q / cpc, e0 = (q - J * cpc) * 8;     // 8 channels: ci = e0/16, a = (e0/4)&3 and a+1, b = 0..3
    const int ci = e0 >> 4, a = (e0 >> 2) & 3;
    const float4 lo = *reinterpret_cast<const float4*>(&rows[(ci * 4 + a) * wu + 4 * J]);
    const float4 hi = *reinterpret_cast<const float4*>(&rows[(ci * 4 + a + 1) * wu + 4 * J]);
    out[q] = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
  }
  // zero the 128 bytes of slack behind the last pixel (read through conv1's 64-element rows; 0-weight x NaN would be NaN)
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 8)
    reinterpret_cast<uint4*>(xs + (size_t)gridDim.x * W4 * c16 * 2)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
}

// mode 0: wp[co][(kh,kw,ci)] = w[co][ci][kh][kw];  mode 1: packed-frames order for a [Cout, Cin, 4KH, 4KW] weight:
// wp[co][(dI,dJ,ci,a,b)] = w[co][ci][4dI+a][4dJ+b];  mode 2: input-gradient classes (ph,pw) of a stride-s conv:
// wp[cls][ci][(a,b,co)] = w[co][ci][ph + a s][pw + b s].
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cout, int Cin, int KH, int KW, int mode,
                                   int stride) {
  const int total = Cout * Cin * KH * KW;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int kw = idx % KW, t = idx / KW, kh = t % KH, t2 = t / KH, ci = t2 % Cin, co = t2 / Cin;
    int o;
    if (mode == 0) {
      o = ((co * KH + kh) * KW + kw) * Cin + ci;
    } else if (mode == 1) {
      const int dI = kh >> 2, a = kh & 3, dJ = kw >> 2, b = kw & 3;
      o = (((co * (KH >> 2) + dI) * (KW >> 2) + dJ) * Cin + ci) * 16 + a * 4 + b;
    } else {
      const int ph = kh % stride, a = kh / stride, pw = kw % stride, b = kw / stride;
      int off = 0;                                       // elements of the classes before (ph, pw)
      for (int c = 0; c < ph * stride + pw; ++c) {
        const int cph = c / stride, cpw = c % stride;
        off += Cin * ((KH - cph + stride - 1) / stride) * ((KW - cpw + stride - 1) / stride) * Cout;
      }
      const int KB = (KW - pw + stride - 1) / stride, KA = (KH - ph + stride - 1) / stride;
      o = off + ci * (KA * KB * Cout) + (a * KB + b) * Cout + co;
    }
    wp[o] = __float2bfloat16(w[idx]);
  }
}

// [N][taps * C] -> [N][taps * 64], zero for the 64 - C trailing positions of every tap (conv1 on the halo path)
__global__ void pad_taps_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ wp, int N, int taps, int C) {
  const int total = N * taps * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 63, t = (i >> 6) % taps, n = i / (64 * taps);
    wp[i] = c < C ? w[((size_t)n * taps + t) * C + c] : __float2bfloat16(0.f);
  }
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int BN, bool DGRAD>
int launch_igemm(const ConvParams& p, cudaStream_t st) {
  auto kern = conv_igemm_kernel<BN, DGRAD>;
  // Measured on B200: ring depth beyond 4 does not help (the gather is issue-bound, not latency-bound) while every
  // 16 KB stage is taken from the unified L1 that serves the overlapping im2col windows (dgrad k4s2: 0.72 -> 0.58 ms).
  // whole-K ring slots when every class has the same, small number of k-tiles and 4 such slots fit next to the weights
  int kts = 1;
  {
    const int nkt0 = p.cls[0].K / KT;
    bool same = nkt0 >= 2 && nkt0 <= 4;
    for (int c = 1; c < p.ncls; ++c) same = same && p.cls[c].K / KT == nkt0;
    if (same && 4 * nkt0 * (int)A_STAGE + p.w_bytes + 4096 <= 227 * 1024 && !(getenv("HULC2_CONV_KTS1") && atoi(getenv("HULC2_CONV_KTS1")))) kts = nkt0;
  }
  int stages = (227 * 1024 - 2048 - p.w_bytes - 1024) / (kts * (int)A_STAGE);
  int want = 4;
  if (const char* e = getenv("HULC2_CONV_STAGES")) { int v = atoi(e); if (v >= 3 && v <= MAX_STAGES) want = v; }
  if (stages > want) stages = want;
  if (stages < 3) { hulc2_set_error("convb: packed weights do not fit in shared memory"); return HULC2_EINVAL; }
  ConvParams q = p;
  q.stages = stages; q.lag = stages - 2; q.kts = kts;
  const int smem = stages * kts * (int)A_STAGE + p.w_bytes + 1024;
  static int configured = 0;
  if (configured < smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      hulc2_set_error("convb: cannot raise dynamic shared memory limit");
      return HULC2_ELAUNCH;
    }
    configured = smem;
  }
  static const bool spread = getenv("HULC2_DGRAD_SPREAD") && atoi(getenv("HULC2_DGRAD_SPREAD")) != 0;
  if (spread) {            // A/B switch: one tile per work item, classes of a region on neighbouring CTAs (L2 reuse only)
    q.ngroups = p.ntiles;
    q.grp_shift = 0;
  } else {
    q.ngroups = p.ntiles >> p.cls_shift;
    q.grp_shift = p.cls_shift;
  }
  const int grid = q.ngroups < sm_count() ? q.ngroups : sm_count();
  kern<<<grid, NT, smem, st>>>(q);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int check_convb(const hulc2_convb_args* a) {
  if (!a || a->F < 0 || a->C <= 0 || a->Cout <= 0 || a->KH <= 0 || a->KW <= 0 || a->stride <= 0 || a->H < a->KH || a->W < a->KW) {
    hulc2_set_error("convb: bad geometry");
    return HULC2_EINVAL;
  }
  if (!hulc2_device_supports_tcgen05()) { hulc2_set_error("convb: needs an sm_100 device (tcgen05)"); return HULC2_ENOTIMPL; }
  return HULC2_OK;
}

}  // namespace

extern "C" {

int hulc2_convb_supported(int C, int Cout, int KH, int KW, int stride) {
  if (C % 8 != 0 || (KH * KW * C) % 64 != 0) return 0;
  if (Cout != 32 && Cout != 64) return 0;
  if (KH * KW * C * Cout * 2 > 100 * 1024) return 0;
  if (KH * KW * C / 8 > 72) return 0;
  if (stride > 2 || stride < 1) return 0;
  return 1;
}

int hulc2_pack_frames_bf16(const float* x, void* xs, int F, int C, int H, int W, cudaStream_t st) {
  if (F <= 0) return HULC2_OK;
  const int H4 = H / 4, W4 = W / 4;
  if (H4 <= 0 || W4 <= 0 || C <= 0 || (16 * C) % 8 != 0) { hulc2_set_error("pack_frames: bad geometry"); return HULC2_EINVAL; }
  const int smem = C * 4 * W4 * 4 * (int)sizeof(float);
  if (smem > 48 * 1024) { hulc2_set_error("pack_frames: frame row too wide"); return HULC2_EINVAL; }
  pack_frames_kernel<<<F * H4, 256, smem, st>>>(x, (uint8_t*)xs, C, H, W, H4, W4);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_convb_pack_weight(const float* w, void* wp, int Cout, int Cin, int KH, int KW, int mode, int stride, cudaStream_t st) {
  if (mode < 0 || mode > 2 || (mode == 1 && ((KH & 3) || (KW & 3)))) { hulc2_set_error("convb_pack_weight: bad mode"); return HULC2_EINVAL; }
  const int total = Cout * Cin * KH * KW;
  if (total <= 0) return HULC2_OK;
  pack_weight_kernel<<<hulc2_cdiv(total, 256), 256, 0, st>>>(w, (__nv_bfloat16*)wp, Cout, Cin, KH, KW, mode, stride > 0 ? stride : 1);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_convb_fwd(const hulc2_convb_args* a, cudaStream_t st) {
  if (int e = check_convb(a)) return e;
  if (!hulc2_convb_supported(a->C, a->Cout, a->KH, a->KW, a->stride)) { hulc2_set_error("convb_fwd: unsupported shape"); return HULC2_ENOTIMPL; }
  const int OH = (a->H - a->KH) / a->stride + 1, OW = (a->W - a->KW) / a->stride + 1;
  if (a->C == 64 && a->Cout == 64 && a->stride == 1 && a->KH * a->KW <= 16 && a->W <= 128 && hulc2_conv_halo_enabled()) {
    // halo-tile path: one TMA box per tile, tap windows addressed by shifted UMMA descriptors (conv_halo_sm100.cu)
    HaloParams h{};
    h.w = (const uint8_t*)a->w; h.bias = a->bias; h.mask = nullptr; h.y = (uint8_t*)a->y; h.relu = a->relu;
    h.NT = h.BNc = 64; h.ncls = 1;
    h.PW = hulc2_conv_halo_pitch(a->W); h.BH = 128 / h.PW < OH ? 128 / h.PW : OH; h.PH = h.BH + a->KH - 1;
    h.i_min = 0; h.j_min = 0; h.ntaps = a->KH * a->KW;
    for (int t = 0; t < h.ntaps; ++t) h.delta[t] = (short)((t / a->KW) * h.PW + t % a->KW);
    h.tiles_per_frame = hulc2_cdiv(OH, h.BH);
    h.oH = OH; h.oW = OW; h.oS = 1;
    h.clsH[0] = (short)OH; h.clsW[0] = (short)OW; h.clsPh[0] = h.clsPw[0] = 0;
    h.mask_out = a->relu ? (uint8_t*)a->mask_bits : nullptr;
    const int rc = hulc2_conv_halo_launch(a->x, a->F, a->H, a->W, h, false, st);
    if (rc == HULC2_OK && h.mask_out) const_cast<hulc2_convb_args*>(a)->mask_bits_written = 1;
    if (rc != HULC2_ENOTIMPL) return rc;
  }
  if (a->C % 8 == 0 && a->C < 64 && a->C >= 32 && a->Cout == 32 && a->stride == 1 && a->KH * a->KW <= 4 && a->W <= 128 &&
      a->workspace && a->workspace_bytes >= (long long)a->Cout * a->KH * a->KW * 128 && hulc2_conv_halo_enabled() &&
      !(getenv("HULC2_CONV1_GATHER") && atoi(getenv("HULC2_CONV1_GATHER")))) {
    // conv1 over space-to-depth pixels of C = 48 channels: 64-element rows at a C-element pitch + weights zero-padded per tap
    // (the caller guarantees slack behind the last pixel: ops.pack_frames)
    const int taps = a->KH * a->KW;
    pad_taps_kernel<<<hulc2_cdiv(a->Cout * taps * 64, 256), 256, 0, st>>>((const __nv_bfloat16*)a->w, (__nv_bfloat16*)a->workspace, a->Cout, taps, a->C);
    HULC2_CHECK_LAUNCH();
    HaloParams h{};
    h.w = (const uint8_t*)a->workspace; h.bias = a->bias; h.mask = nullptr; h.y = (uint8_t*)a->y; h.relu = a->relu;
    h.NT = h.BNc = 32; h.ncls = 1;
    h.PW = a->W; h.BH = 128 / h.PW < OH ? 128 / h.PW : OH; h.PH = h.BH + a->KH - 1;
    h.i_min = 0; h.j_min = 0; h.ntaps = taps;
    for (int t = 0; t < taps; ++t) h.delta[t] = (short)((t / a->KW) * h.PW + t % a->KW);
    h.tiles_per_frame = hulc2_cdiv(OH, h.BH);
    h.oH = OH; h.oW = OW; h.oS = 1;
    h.clsH[0] = (short)OH; h.clsW[0] = (short)OW; h.clsPh[0] = h.clsPw[0] = 0;
    h.mask_out = a->relu ? (uint8_t*)a->mask_bits : nullptr;
    const int rc = hulc2_conv_halo_launch_packed(a->x, a->F, a->H, a->W, a->C, h, st);
    if (rc == HULC2_OK && h.mask_out) const_cast<hulc2_convb_args*>(a)->mask_bits_written = 1;
    if (rc != HULC2_ENOTIMPL) return rc;
  }
  if (a->C == 32 && a->Cout == 64 && a->stride == 2 && a->KH == 4 && a->KW == 4 && OW + 1 <= 128 && hulc2_conv_halo_enabled()) {
    // stride-2 halo path: two pixels = one 128-byte row, even / odd source rows in two sub-tiles (conv_halo_sm100.cu);
    // tap t = kh * 2 + kw / 2 <-> the t-th 64-element chunk of the packed weight rows [(kh, kw, ci)]
    HaloParams h{};
    h.w = (const uint8_t*)a->w; h.bias = a->bias; h.mask = nullptr; h.y = (uint8_t*)a->y; h.relu = a->relu;
    h.NT = h.BNc = 64; h.ncls = 1;
    h.PW = OW + 1; h.BH = 128 / h.PW < OH ? 128 / h.PW : OH; h.PH = h.BH + 1;
    h.i_min = 0; h.j_min = 0; h.ntaps = 8;
    for (int t = 0; t < 8; ++t) {
      const int kh = t >> 1, bp = t & 1;
      h.part[t] = (short)(kh & 1);
      h.delta[t] = (short)((kh >> 1) * h.PW + bp);
    }
    h.tiles_per_frame = hulc2_cdiv(OH, h.BH);
    h.oH = OH; h.oW = OW; h.oS = 1;
    h.clsH[0] = (short)OH; h.clsW[0] = (short)OW; h.clsPh[0] = h.clsPw[0] = 0;
    h.mask_out = a->relu ? (uint8_t*)a->mask_bits : nullptr;
    const int rc = hulc2_conv_halo_launch_s2(a->x, a->F, a->H, a->W, h, st);
    if (rc == HULC2_OK && h.mask_out) const_cast<hulc2_convb_args*>(a)->mask_bits_written = 1;
    if (rc != HULC2_ENOTIMPL) return rc;
  }
  ConvParams p{};
  p.x = (const uint8_t*)a->x; p.w = (const uint8_t*)a->w; p.bias = a->bias; p.mask = nullptr; p.y = (uint8_t*)a->y;
  p.ncls = 1; p.N = a->Cout; p.relu = a->relu;
  p.VH = OH; p.VW = OW; p.oH = OH; p.oW = OW;
  ConvClass& c = p.cls[0];
  c.M = a->F * OH * OW; c.dHW = make_fastdiv(OH * OW); c.dW = make_fastdiv(OW); c.K = a->KH * a->KW * a->C; c.w_off = 0; c.tab_off = 0;
  c.sF = a->H * a->W * a->C / 8; c.sI = a->stride * a->W * a->C / 8; c.sJ = a->stride * a->C / 8;
  c.oS = 1; c.oPh = 0; c.oPw = 0;
  if (c.M == 0) return HULC2_OK;
  p.ntiles = hulc2_cdiv(c.M, TILE_M); p.cls_shift = 0;
  p.w_bytes = a->Cout * c.K * 2;
  const int cpr = a->KW * a->C / 8;                    // chunks per kernel row (contiguous in memory)
  p.ntab = c.K / 8;
  for (int q = 0; q < p.ntab; ++q) {
    const int kh = q / cpr;
    p.table[q] = make_int2(kh * a->W * a->C / 8 + (q - kh * cpr), 0);
  }
  return a->Cout == 32 ? launch_igemm<32, false>(p, st) : launch_igemm<64, false>(p, st);
}

int hulc2_convb_dgrad(const hulc2_convb_args* a, cudaStream_t st) {
  if (int e = check_convb(a)) return e;
  // as a GEMM: N = C (input channels), K = taps * Cout
  if (a->Cout % 8 != 0 || (a->C != 32 && a->C != 64) || a->stride > 2) { hulc2_set_error("convb_dgrad: unsupported shape"); return HULC2_ENOTIMPL; }
  const int s = a->stride;
  const int OH = (a->H - a->KH) / s + 1, OW = (a->W - a->KW) / s + 1;
  if (a->Cout == 64 && hulc2_conv_halo_enabled() &&
      ((s == 1 && a->C == 64 && a->KH * a->KW <= 16) || (s == 2 && a->C == 32 && a->KH == 4 && a->KW == 4))) {
    // halo-tile path over dZ (64 channels per pixel); zero padding = TMA out-of-range fill.  s == 2: the four parity
    // classes (2 x 2 taps each) share one halo tile, their packed weights are stacked along N.
    HaloParams h{};
    h.w = (const uint8_t*)a->w; h.bias = nullptr; h.mask = (const uint8_t*)a->xmask; h.y = (uint8_t*)a->dx; h.relu = 0;
    h.mask_bits = (const uint8_t*)a->mask_bits;
    h.BNc = a->C; h.ncls = s * s; h.NT = h.ncls * h.BNc;
    const int KA = a->KH / s, KB = a->KW / s;                    // taps per class along each axis
    const int mH = (a->H + s - 1) / s, mW = (a->W + s - 1) / s;  // largest class
    h.PW = hulc2_conv_halo_pitch(mW + KB - 1); h.i_min = -(KA - 1); h.j_min = -(KB - 1);
    if (h.PW <= 128) {
      h.BH = 128 / h.PW < mH ? 128 / h.PW : mH; h.PH = h.BH + KA - 1;
      h.ntaps = KA * KB;
      for (int t = 0; t < h.ntaps; ++t) h.delta[t] = (short)((KA - 1 - t / KB) * h.PW + (KB - 1 - t % KB));
      h.tiles_per_frame = hulc2_cdiv(mH, h.BH);
      h.oH = a->H; h.oW = a->W; h.oS = s;
      for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
          const int c = ph * s + pw;
          h.clsH[c] = (short)((a->H - ph + s - 1) / s); h.clsW[c] = (short)((a->W - pw + s - 1) / s);
          h.clsPh[c] = (short)ph; h.clsPw[c] = (short)pw;
        }
      const int rc = hulc2_conv_halo_launch(a->dy, a->F, OH, OW, h, true, st);
      if (rc != HULC2_ENOTIMPL) return rc;
    }
  }
  ConvParams p{};
  p.x = (const uint8_t*)a->dy; p.w = (const uint8_t*)a->w; p.bias = nullptr; p.mask = (const uint8_t*)a->xmask; p.y = (uint8_t*)a->dx;
  p.N = a->C; p.relu = 0;
  p.VH = OH; p.VW = OW; p.oH = a->H; p.oW = a->W;
  int ncls = 0, tiles = 0, woff = 0, tab = 0;
  const int co8 = a->Cout / 8;
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      const int CH = (a->H - ph + s - 1) / s, CW = (a->W - pw + s - 1) / s;
      const int KA = (a->KH - ph + s - 1) / s, KB = (a->KW - pw + s - 1) / s;
      const int K = KA * KB * a->Cout;
      if (CH <= 0 || CW <= 0) { woff += a->C * K * 2; continue; }
      if (K <= 0 || K % 64 != 0 || tab + K / 8 > MAX_TAB || ncls >= MAX_CLS) { hulc2_set_error("convb_dgrad: unsupported class shape"); return HULC2_ENOTIMPL; }
      ConvClass& c = p.cls[ncls++];
      c.M = a->F * CH * CW; c.dHW = make_fastdiv(CH * CW); c.dW = make_fastdiv(CW); c.K = K; c.w_off = woff; c.tab_off = tab;
      c.sF = OH * OW * co8; c.sI = OW * co8; c.sJ = co8;
      c.oS = s; c.oPh = ph; c.oPw = pw;
      for (int q = 0; q < K / 8; ++q) {
        const int tap = q / co8, u = q - tap * co8, ta = tap / KB, tb = tap - ta * KB;
        p.table[tab + q] = make_int2(-(ta * OW + tb) * co8 + u, (ta << 16) | tb);
      }
      tab += K / 8; woff += a->C * K * 2;
      if (hulc2_cdiv(c.M, TILE_M) > tiles) tiles = hulc2_cdiv(c.M, TILE_M);
    }
  if (ncls != s * s) { hulc2_set_error("convb_dgrad: degenerate parity classes"); return HULC2_ENOTIMPL; }
  p.cls_shift = s == 1 ? 0 : 2;
  p.ncls = ncls; p.ntiles = tiles << p.cls_shift; p.w_bytes = woff; p.ntab = tab;
  if (tiles == 0) return HULC2_OK;
  return a->C == 32 ? launch_igemm<32, true>(p, st) : launch_igemm<64, true>(p, st);
}

int hulc2_convb_wgrad(const hulc2_convb_args* a, cudaStream_t st) {
  if (int e = check_convb(a)) return e;
  if (!hulc2_convb_supported(a->C, a->Cout, a->KH, a->KW, a->stride)) { hulc2_set_error("convb_wgrad: unsupported shape"); return HULC2_ENOTIMPL; }
  const int OH = (a->H - a->KH) / a->stride + 1, OW = (a->W - a->KW) / a->stride + 1;
  // Halo-tile weight gradient (conv_halo_sm100.cu): the source tile is staged ONCE per tile and the taps are read through shifted
  // MN-major descriptors, instead of gathering every source pixel once per tap.  HULC2_WGRAD_HALO: unset = the cp.async-producer
  // variant for conv1 over packed frames (48-channel pixels, 32 outputs: 0.80 -> 0.46 ms at F = 4096, the gather kernel moves
  // 448 bytes per output pixel through LDGSTS, this one 160) and the gather kernel for everything else (64-channel sources:
  // measured 0.24 ms gather vs 0.42-0.68 ms halo); 0 = gather everywhere; 1 = TMA-box variant, 2 = cp.async variant wherever
  // the shape fits (both slower on conv3: their 5 M-tiles of shifted-descriptor MMAs pace them).
  static const int wg_halo = getenv("HULC2_WGRAD_HALO") ? atoi(getenv("HULC2_WGRAD_HALO")) : -1;
  const bool conv1_like = a->C == 48 && a->Cout == 32;
  if (a->stride == 1 && a->C >= 32 && a->C <= 64 && (a->Cout == 32 || a->Cout == 64) && a->F * OH * OW > 0 && hulc2_conv_halo_enabled() &&
      (wg_halo > 0 || (wg_halo < 0 && conv1_like))) {
    int grid = 0, nblk = 0;
    const int rc = hulc2_conv_halo_wgrad(a->x, a->C, a->dy, a->Cout, a->F, a->H, a->W, a->KH, a->KW, (float*)a->workspace, a->workspace_bytes,
                                         &grid, &nblk, st);
    if (rc == HULC2_OK) {
      const int K = a->KH * a->KW * a->C, total = (K + 1) * a->Cout;
      wgrad_reduce_kernel<<<hulc2_cdiv(total, 256), 256, 0, st>>>((const float*)a->workspace, grid, nblk * 64, K, a->Cout, a->C, a->KH, a->KW,
                                                                   a->dw_layout, a->dw, a->db, 64);
      HULC2_CHECK_LAUNCH();
      return HULC2_OK;
    }
    if (rc != HULC2_ENOTIMPL) return rc;
  }
  WgradParams p{};
  p.x = (const uint8_t*)a->x; p.dz = (const uint8_t*)a->dy;
  p.P = a->F * OH * OW; p.dHW = make_fastdiv(OH * OW); p.dW = make_fastdiv(OW);
  p.sF = a->H * a->W * a->C / 8; p.sI = a->stride * a->W * a->C / 8; p.sJ = a->stride * a->C / 8;
  p.K = a->KH * a->KW * a->C;
  p.nblk = p.K / 64 + 1; if (p.nblk & 1) ++p.nblk;
  if (p.nblk > WG_MAX_BLK) { hulc2_set_error("convb_wgrad: kconv too large"); return HULC2_ENOTIMPL; }
  p.cout8 = a->Cout / 8;
  // 64-pixel stages when four of them fit (conv1: 5 blocks x 8 KB), else 32-pixel stages
  const int KP = (p.nblk + 1) * 64 * 128 * 4 + 4096 <= 227 * 1024 ? 64 : 32;
  const int WG_BLK = KP * 128;
  p.nstages = hulc2_cdiv(p.P, KP);
  const int cpr = a->KW * a->C / 8;
  for (int q = 0; q < p.K / 8; ++q) {
    const int kh = q / cpr;
    p.delta[q] = kh * a->W * a->C / 8 + (q - kh * cpr);
  }
  if (p.P == 0) {
    if (a->dw) cudaMemsetAsync(a->dw, 0, sizeof(float) * a->Cout * p.K, st);
    if (a->db) cudaMemsetAsync(a->db, 0, sizeof(float) * a->Cout, st);
    return HULC2_OK;
  }
  const int grid = p.nstages < sm_count() ? p.nstages : sm_count();
  const long long need = (long long)grid * p.nblk * 64 * 64 * sizeof(float);
  if (!a->workspace || a->workspace_bytes < need) { hulc2_set_error("convb_wgrad: workspace too small"); return HULC2_EWORKSPACE; }
  p.partial = (float*)a->workspace;
  int stages = (227 * 1024 - 2048 - 1024) / ((p.nblk + 1) * (int)WG_BLK);
  int want = 4;
  if (const char* e = getenv("HULC2_CONV_STAGES")) { int v = atoi(e); if (v >= 3 && v <= WG_MAX_STAGES) want = v; }
  if (stages > want) stages = want;
  p.stages = stages; p.lag = stages - 2;
  const int smem = stages * (p.nblk + 1) * (int)WG_BLK + 1024;
  static int configured = 0;
  if (configured < smem) {
    if (cudaFuncSetAttribute(conv_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      hulc2_set_error("convb_wgrad: cannot raise dynamic shared memory limit");
      return HULC2_ELAUNCH;
    }
    configured = smem;
  }
  if (KP == 64) conv_wgrad_kernel<64><<<grid, NT_WG, smem, st>>>(p);
  else conv_wgrad_kernel<32><<<grid, NT_WG, smem, st>>>(p);
  HULC2_CHECK_LAUNCH();
  const int total = (p.K + 1) * a->Cout;
  wgrad_reduce_kernel<<<hulc2_cdiv(total, 256), 256, 0, st>>>(p.partial, grid, p.nblk * 64, p.K, a->Cout, a->C, a->KH, a->KW, a->dw_layout,
                                                               a->dw, a->db, a->C);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
