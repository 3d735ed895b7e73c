// Halo-tile implicit-GEMM convolutions for sm_100a: the im2col operand is never gathered.
//
// For a stride-1 correlation over a 64-channel bf16 NHWC source (conv3 forward, conv3 input gradient, and -- per stride
// parity class -- conv2's input gradient), the A operand of tap (a, b) is the SAME source tile read at a row offset:
// with the tile stored as a raster of PW pixels per row (128 bytes = 64 channels per pixel, SWIZZLE_128B), output
// position m = il * PW + jl needs source row m + delta(a, b), delta = (a - a_min) * PW + (b - b_min).  So
//   * ONE 4-D TMA box load {64 ch, PW, PH, 1 frame} per tile brings the halo tile into shared memory (out-of-range
//     rows / columns are zero-filled by TMA: the zero padding of the input-gradient correlation comes for free);
//   * the MMA thread issues, per tap, tcgen05.mma with the A descriptor's start address advanced by delta * 128 bytes.
//     The swizzle is a function of the absolute shared-memory address, so a start address that is not a multiple of the
//     8-row atom reads the shifted rows correctly (measured: tools/probes/umma_shift_probe.cu, all offsets 0..127 exact);
//   * positions with jl >= (valid width) are garbage columns (the window wraps into the next raster row): computed,
//     never stored.  M = 128 positions per tile = BH full rows of the (class-)output.
// Each source byte is fetched from HBM/L2 once per tile (+ halo rows) instead of once per tap by 16-byte cp.async:
// no producer warps, no per-thread address arithmetic.  Weights stay resident in shared memory as in conv_sm100.cu.
// The s*s = 4 parity classes of conv2's input gradient share one halo tile: their packed weights are stacked along N
// (4 x 32 input channels = one 128-column accumulator), the epilogue scatters the classes to their strided pixels.
//
// Warp roles (416 threads): warps 0-7 epilogue (TMEM lane quarter = w % 4, column half = w / 4), warp 8 TMA producer,
// warps 9-12 MMA issuers (4 for N = 64, 2 for N = 128; round-robin over the CTA's tiles, one TMEM accumulator each).  Reference: hulc2/models/perceptual_encoders/vision_network.py:38-48 (+ autograd).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_halo.cuh"

using namespace sm100;

namespace {

constexpr int NT_HALO = 416;     // 8 epilogue warps + TMA producer (warp 8) + 4 MMA issuers (warps 9-12)
constexpr int MAX_ST = 8;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

template <int NTOT, bool DGRAD>
__global__ void __launch_bounds__(NT_HALO, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_ST], empty_bar[MAX_ST], tfull_bar[4], tempty_bar[4];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[NTOT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ST = (uint32_t)p.stages;
  const uint32_t a_smem = base;                                 // ST halo tiles
  const uint32_t w_smem = base + ST * (uint32_t)p.stage_bytes;  // ntaps tiles of [NTOT rows][128 B]
  constexpr uint32_t W_TILE = NTOT * 128;
  // Several MMA issuer threads take the CTA's tiles round-robin, each with its own TMEM accumulator: with N <= 128 a
  // 128 x N x 16 MMA occupies the tensor pipe for only 32-64 cycles while ONE thread issues a tcgen05.mma every ~110 cycles
  // (measured: 1 -> 2 issuers, conv3 forward 0.332 -> 0.179 ms, conv2 forward 0.308 -> 0.197 ms).
  // (Interleaving several tiles' MMAs from ONE thread was measured and does not help: 0.283 -> 0.334 ms on conv3 forward.)
  constexpr int NISS = 256 / NTOT >= 4 ? 4 : 2;   // N <= 64: 4 issuers; N = 128: 2 issuers (TMEM: NB x N <= 512 columns)
  constexpr int NB = NISS < 2 ? 2 : NISS;
  constexpr uint32_t TCOLS = NB * NTOT;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
  if (tid == 32) {
    for (uint32_t s = 0; s < ST; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int b = 0; b < NB; ++b) { mbar_init(smem_u32(&tfull_bar[b]), 1); mbar_init(smem_u32(&tempty_bar[b]), 8); }   // tempty: one arrival per epilogue warp
    mbar_fence_init();
  }
  if (tid < NTOT) bias_s[tid] = (!DGRAD && p.bias) ? p.bias[tid] : 0.f;
  {  // resident weights: [NTOT][K] bf16, K = ntaps * 64 -> per tap a K-major SWIZZLE_128B tile
    const int cpr = p.ntaps * 8;                                // 16-byte chunks per weight row
    for (int ch = tid; ch < NTOT * cpr; ch += NT_HALO) {
      const int n = ch / cpr, q = ch - n * cpr;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w + ((size_t)n * cpr + q) * 16));
      const uint32_t dst = w_smem + (uint32_t)(q >> 3) * W_TILE + swz128(n, q & 7);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == 8) {
    // ===================================================== TMA producer (one thread): one halo tile per output tile
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm) : "memory");
      uint32_t s = 0, ph = 1;
      const uint32_t bytes = (uint32_t)(p.PH * p.PW) * 128u * (uint32_t)p.nparts;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int f = tile / p.tiles_per_frame, ti = tile - f * p.tiles_per_frame;
        mbar_wait(smem_u32(&empty_bar[s]), ph);
        const uint32_t bar = smem_u32(&full_bar[s]);
        mbar_expect_tx(bar, bytes);
        const uint32_t dst = a_smem + s * (uint32_t)p.stage_bytes;
        if (p.nparts == 1) {
          tma_load_4d(dst, &tm, bar, 0, p.j_min, ti * p.BH + p.i_min, f);
        } else {            // 5-D view (64 = 2 pixels x 32 ch, pixel pair, row parity, half row, frame): one box per row parity
          tma_load_5d(dst, &tm, bar, 0, p.j_min, 0, ti * p.BH + p.i_min, f);
          tma_load_5d(dst + (uint32_t)p.part_bytes, &tm, bar, 0, p.j_min, 1, ti * p.BH + p.i_min, f);
        }
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 9 && warp < 9 + NISS) {
    // ===================================================== MMA issuers (one thread each): issuer w takes tiles c = w, w + NISS, ..
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(128, NTOT, false, false);
      const uint32_t w = (uint32_t)(warp - 9);
      uint32_t c = 0, s = 0, ph = 0;                             // c: tiles of this CTA so far; s / ph: ring slot and parity of tile c
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++c) {
        if ((c & (NISS - 1)) == w) {
          const uint32_t buf = c % NB;                           // == w: each issuer owns one accumulator
          mbar_wait(smem_u32(&tempty_bar[buf]), ((c / NB) & 1) ^ 1);     // epilogue drained this accumulator
          mbar_wait(smem_u32(&full_bar[s]), ph);                         // halo tile landed
          tc_fence_after();
          const uint32_t tile_addr = a_smem + s * (uint32_t)p.stage_bytes;
          for (int t = 0; t < p.ntaps; ++t) {
            const uint64_t ad = make_desc(tile_addr + (uint32_t)p.delta[t] * 128u + (uint32_t)p.part[t] * (uint32_t)p.part_bytes, 0);
            const uint64_t bd = make_desc(w_smem + t * W_TILE, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_d + buf * NTOT, ad + 2 * k, bd + 2 * k, IDESC, (t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));                          // halo tile reusable once these MMAs retire
          umma_commit(smem_u32(&tfull_bar[buf]));                        // accumulator complete
        }
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ===================================================== epilogue: row m = 32 * (w % 4) + lane, columns (w / 4) * NTOT/2 ..
    constexpr int HC = NTOT / 2;
    const int lq = warp & 3, half = warp >> 2;
    const int m = lq * 32 + lane;
    const int il = m / p.PW, jl = m - il * p.PW;
    const int BNc = p.BNc;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int f = tile / p.tiles_per_frame, tr = tile - f * p.tiles_per_frame;
      const int i = tr * p.BH + il;
      const uint32_t buf = ti % NB, par = (ti / NB) & 1;
      // output addresses and the ReLU-mask words of this thread's chunks BEFORE waiting for the accumulator: the loads'
      // latency hides behind the tile's MMAs
      constexpr int NCH = HC / 16;
      long long off[NCH];
      bool ok[NCH];
      uint32_t mk[NCH][8];
      uint32_t mb[NCH];           // ... or its 16 sign bits (r02: 2 bytes instead of 32 per chunk -- the input gradient of conv2 read
                                  // 629 MB of bf16 activations per step only to test their sign)
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int col = half * HC + q * 16;             // first of 16 accumulator columns: one class, 16 channels
        const int cls = col / BNc, ch = col - cls * BNc;
        ok[q] = il < p.BH && i < p.clsH[cls] && jl < p.clsW[cls];
        const long long opix = ((long long)f * p.oH + i * p.oS + p.clsPh[cls]) * p.oW + jl * p.oS + p.clsPw[cls];
        off[q] = (opix * BNc + ch) * 2;
        if (DGRAD && p.mask_bits) {
          mb[q] = ok[q] ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.mask_bits + (off[q] >> 4))) : 0u;
        } else if (DGRAD && p.mask && ok[q]) ldg256_nc(p.mask + off[q], mk[q]);      // 32 bytes = this chunk's 16 channels
      }
      mbar_wait_relaxed(smem_u32(&tfull_bar[buf]), par);
      tc_fence_after();
      uint32_t acc[HC];
#pragma unroll
      for (int c = 0; c < HC; c += 16)
        tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + buf * NTOT + half * HC + c, *reinterpret_cast<uint32_t(*)[16]>(&acc[c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
      uint32_t mybits = 0;
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        if (!ok[q]) continue;
        const int c = q * 16, col = half * HC + c;
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float lo = __uint_as_float(acc[c + 2 * e]), hi = __uint_as_float(acc[c + 2 * e + 1]);
          if (!DGRAD) {
            lo += bias_s[col + 2 * e]; hi += bias_s[col + 2 * e + 1];
            if (p.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
          }
          o[e] = pack_bf16x2(lo, hi);
        }
        if (!DGRAD && p.mask_out) {
          // sign bits of this chunk's 16 outputs, the same predicate the input gradient applies to the bf16 values
          uint32_t bits = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t mm = o[e];
            bits |= (((mm & 0x7fffu) != 0 && !(mm & 0x8000u)) ? 1u : 0u) << (2 * e);
            bits |= (((mm & 0x7fff0000u) != 0 && !(mm & 0x80000000u)) ? 1u : 0u) << (2 * e + 1);
          }
          mybits |= bits << (16 * (q & 1));
        }
        if (DGRAD && p.mask_bits) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] &= (((mb[q] >> (2 * e)) & 1u) ? 0x0000ffffu : 0u) | (((mb[q] >> (2 * e + 1)) & 1u) ? 0xffff0000u : 0u);
        } else if (DGRAD && p.mask) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            // bf16 activations are >= 0 after ReLU: keep a half-word where the mask half-word is a positive number
            const uint32_t mm = mk[q][e];
            const uint32_t keep = (((mm & 0x7fffu) != 0 && !(mm & 0x8000u)) ? 0x0000ffffu : 0u) |
                                  (((mm & 0x7fff0000u) != 0 && !(mm & 0x80000000u)) ? 0xffff0000u : 0u);
            o[e] &= keep;
          }
        }
        stg256(p.y + off[q], o);
      }
      if (!DGRAD && NCH <= 2 && p.mask_out && ok[0]) {
        // the thread's 16 / 32 sign bits as one aligned 2- / 4-byte store.  Measured: worth it for the 64-channel layers only (conv2
        // forward +15 us, conv3 input gradient -50 us per 4096 frames); on conv1 forward, whose epilogue paces the kernel (16 columns
        // per thread, 860 cycles of HBM time per tile), the extra ~50 ALU instructions cost 0.10 ms -- more than conv2's input
        // gradient gains (0.06 ms) -- and pairing the two column halves of a pixel through shared memory for full-word stores
        // is worse still (0.38 -> 0.54 ms): ops requests the bits for conv2's output only.
        if (NCH == 2) *reinterpret_cast<uint32_t*>(p.mask_out + (off[0] >> 4)) = mybits;
        else *reinterpret_cast<unsigned short*>(p.mask_out + (off[0] >> 4)) = (unsigned short)mybits;
      }
      ++ti;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TCOLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn halo_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

int halo_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int NTOT, bool DGRAD>
int launch_halo(const CUtensorMap& tm, const HaloParams& p, int smem, cudaStream_t st) {
  auto kern = conv_halo_kernel<NTOT, DGRAD>;
  static int configured = 0;
  if (configured < smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      hulc2_set_error("conv_halo: cannot raise dynamic shared memory limit");
      return HULC2_ELAUNCH;
    }
    configured = smem;
  }
  const int grid = p.ntiles < halo_sms() ? p.ntiles : halo_sms();
  kern<<<grid, NT_HALO, smem, st>>>(tm, p);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // namespace

int hulc2_conv_halo_pitch(int pw) {
  // experiment knob: round the raster pitch up to a multiple of HULC2_HALO_PITCH_ALIGN (extra columns are TMA zero fill)
  static const int al = getenv("HULC2_HALO_PITCH_ALIGN") ? atoi(getenv("HULC2_HALO_PITCH_ALIGN")) : 1;
  return al > 1 ? (pw + al - 1) / al * al : pw;
}

bool hulc2_conv_halo_enabled() {
  static const bool off = getenv("HULC2_CONV_HALO") && atoi(getenv("HULC2_CONV_HALO")) == 0;
  return !off && halo_encode_fn() != nullptr;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Common tail of the launchers: ring sizing + launch.  `rank`-D tensor map already described by dims/strides/box.
static int halo_finish(const void* src, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, HaloParams p,
                       bool dgrad, cudaStream_t st) {
  EncodeTiledFn2 encode = (EncodeTiledFn2)halo_encode_fn();
  if (!encode) return HULC2_ENOTIMPL;
  if (p.PW > 256 || p.PH > 256 || p.BH < 1 || p.BH * p.PW > 128 || p.ntaps < 1 || p.ntaps > 16) return HULC2_ENOTIMPL;
  if (p.NT != 32 && p.NT != 64 && p.NT != 128) return HULC2_ENOTIMPL;
  if (((uintptr_t)src & 15) != 0) return HULC2_ENOTIMPL;
  int dmax = 0;
  for (int t = 0; t < p.ntaps; ++t) dmax = p.delta[t] > dmax ? p.delta[t] : dmax;
  const int rows = p.PH * p.PW > dmax + 128 ? p.PH * p.PW : dmax + 128;
  p.part_bytes = ((rows * 128) + 1023) & ~1023;
  p.stage_bytes = p.part_bytes * p.nparts;
  const int w_bytes = p.NT * p.ntaps * 128;
  int stages = (227 * 1024 - 2048 - w_bytes - 1024) / p.stage_bytes;
  if (stages > MAX_ST) stages = MAX_ST;
  // >= one ring slot per MMA issuer (4 for N = 64, 2 for N = 128): an issuer then never waits on a slot whose previous fill is
  // still pending, which is what keeps the parity waits of several issuers unambiguous
  if (stages < (p.NT <= 64 ? 4 : 2)) return HULC2_ENOTIMPL;
  p.stages = stages;
  p.ntiles = p.F * p.tiles_per_frame;
  if (p.ntiles <= 0) return HULC2_OK;
  CUtensorMap tm;
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(src), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return HULC2_ENOTIMPL;
  const int smem = stages * p.stage_bytes + w_bytes + 1024;
  if (!dgrad) return p.NT == 64 ? launch_halo<64, false>(tm, p, smem, st) : (p.NT == 32 ? launch_halo<32, false>(tm, p, smem, st) : HULC2_ENOTIMPL);
  if (p.NT == 32) return HULC2_ENOTIMPL;
  return p.NT == 64 ? launch_halo<64, true>(tm, p, smem, st) : launch_halo<128, true>(tm, p, smem, st);
}

// src: bf16 NHWC [F, Hs, Ws, 64].  The caller sets the geometry (BH, PW, PH, i_min, j_min, ntaps, delta[], cls*[], NT, BNc,
// ncls, oH, oW, oS, tiles_per_frame).  Returns HULC2_ENOTIMPL when the shape does not fit (the caller then uses the
// gather kernel).
int hulc2_conv_halo_launch(const void* src, int F, int Hs, int Ws, HaloParams p, bool dgrad, cudaStream_t st) {
  p.F = F; p.nparts = 1;
  for (int t = 0; t < 16; ++t) p.part[t] = 0;
  cuuint64_t dims[4] = {64u, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)F};
  cuuint64_t strides[3] = {128u, (cuuint64_t)Ws * 128u, (cuuint64_t)Hs * Ws * 128u};
  cuuint32_t box[4] = {64u, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1u};
  return halo_finish(src, 4, dims, strides, box, p, dgrad, st);
}

// src: bf16 NHWC [F, Hs, Ws, 32], stride-2 forward conv.  View (64 = 2 pixels x 32 ch | pixel pair | row parity | half row |
// frame): output column j reads pairs j + b/2, output row i reads half rows i + a/2 of parity a & 1 -- two sub-tiles per
// stage, every tap again a uniform raster shift.  p.part[] / p.delta[] are set by the caller.
int hulc2_conv_halo_launch_s2(const void* src, int F, int Hs, int Ws, HaloParams p, cudaStream_t st) {
  p.F = F; p.nparts = 2;
  const cuuint64_t row = (cuuint64_t)Ws * 64u;      // bytes per source row
  if (row % 16) return HULC2_ENOTIMPL;
  // Hs / 2 half rows and Ws / 2 pairs (floor): every coordinate inside these bounds addresses a pixel of its own frame, and a
  // valid output never needs more (largest source row read is 2 (OH - 1) + 3 <= Hs - 1)
  cuuint64_t dims[5] = {64u, (cuuint64_t)(Ws / 2), 2u, (cuuint64_t)(Hs / 2), (cuuint64_t)F};
  cuuint64_t strides[4] = {128u, row, 2u * row, (cuuint64_t)Hs * row};
  cuuint32_t box[5] = {64u, (cuuint32_t)p.PW, 1u, (cuuint32_t)p.PH, 1u};
  return halo_finish(src, 5, dims, strides, box, p, false, st);
}

// Forward conv over "packed" pixels of `pixel_elems` (< 64) bf16 channels: the tensor map describes 64-element rows at a
// pitch of pixel_elems elements, i.e. every shared-memory row holds one pixel plus the first 64 - pixel_elems channels of its
// right neighbour; the caller's weights are zero for those K positions, so the overlap contributes nothing.  This puts
// conv1 (48 channels per space-to-depth pixel, 2 x 2 taps) on the halo path without changing the packed-frame layout.
// The source must have >= 128 bytes of slack behind its last pixel.
int hulc2_conv_halo_launch_packed(const void* src, int F, int Hs, int Ws, int pixel_elems, HaloParams p, cudaStream_t st) {
  p.F = F; p.nparts = 1;
  for (int t = 0; t < 16; ++t) p.part[t] = 0;
  const cuuint64_t pb = (cuuint64_t)pixel_elems * 2;
  if (pb % 16 || pixel_elems > 64) return HULC2_ENOTIMPL;
  cuuint64_t dims[4] = {64u, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)F};
  cuuint64_t strides[3] = {pb, (cuuint64_t)Ws * pb, (cuuint64_t)Hs * Ws * pb};
  cuuint32_t box[4] = {64u, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1u};
  return halo_finish(src, 4, dims, strides, box, p, false, st);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// Both operands MN-major (tile rows = pixel positions = the contraction index).  Per tile: ONE TMA box of the source
// (BH + KH - 1 rows) and ONE of dZ (BH rows, out-of-range columns / rows zero-filled, so garbage raster positions contribute
// nothing); tap (a, b) is the source tile read at row offset a * PW + b.  An M = 128 MMA covers two taps: the descriptor's
// leading-dimension offset is the distance between the two taps' windows (measured: tools/probes/umma_mn_shift_probe.cu);
// the last block is a constant 16-row tile of ones (its LBO is recomputed per k-step), which yields the bias gradient.
// dW^T stays in TMEM for the CTA's lifetime; up to 3 issuer threads own disjoint M-tiles.
// Warp roles (416 threads): warps 0-7 dump, warp 8 TMA producer, warps 9-11 MMA issuers.
namespace {

template <int DUMMY>
__global__ void __launch_bounds__(NT_HALO, 1) conv_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmz,
                                                                     const __grid_constant__ HaloWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_ST], empty_bar[MAX_ST], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ST = (uint32_t)p.stages;
  const uint32_t stage_bytes = (uint32_t)(p.a_bytes + p.b_bytes);
  const uint32_t ones_smem = base + ST * stage_bytes;           // 32 rows x 128 B of bf16 1.0 (16 used; a dummy second block may start at row 1)
  const int nmt = p.nmt;
  const int niss = nmt < 3 ? nmt : 3;
  const uint32_t tcols = nmt * 64 <= 64 ? 64u : (nmt * 64 <= 128 ? 128u : (nmt * 64 <= 256 ? 256u : 512u));

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), tcols);
  if (tid == 32) {
    for (uint32_t s = 0; s < ST; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), (uint32_t)niss); }
    mbar_init(smem_u32(&done_bar), (uint32_t)niss);
    mbar_fence_init();
  }
  // zero all stages once (rows behind a TMA box are read by the last taps' windows: they must be finite), then the ones tile
  for (uint32_t o = tid * 16; o < ST * stage_bytes; o += NT_HALO * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
  for (uint32_t o = tid * 16; o < 32 * 128; o += NT_HALO * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_smem + o), "r"(0x3F803F80u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int KL = p.BH * p.PW;                                    // contraction positions per tile (multiple of 16)

  if (warp == 8) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmx) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmz) : "memory");
      uint32_t s = 0, ph = 1;
      const uint32_t bytes = (uint32_t)((p.BH + p.KH - 1) * p.PW + p.BH * p.PW) * 128u;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int f = tile / p.tiles_per_frame, ti = tile - f * p.tiles_per_frame;
        mbar_wait(smem_u32(&empty_bar[s]), ph);
        const uint32_t bar = smem_u32(&full_bar[s]);
        mbar_expect_tx(bar, bytes);
        const uint32_t dst = base + s * stage_bytes;
        tma_load_4d(dst, &tmx, bar, 0, 0, ti * p.BH, f);
        tma_load_4d(dst + (uint32_t)p.a_bytes, &tmz, bar, 0, 0, ti * p.BH, f);
        if (++s == ST) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 9 && warp < 9 + niss) {
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(128, 64, true, true);
      const int me = warp - 9;
      // M-tiles of this issuer: a contiguous share
      const int mt0 = (nmt * me) / niss, mt1 = (nmt * (me + 1)) / niss;
      uint32_t s = 0, ph = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t a_tile = base + s * stage_bytes, b_tile = a_tile + (uint32_t)p.a_bytes;
        for (int ks = 0; ks < KL / 16; ++ks) {
          const uint64_t bd = make_desc(b_tile + ks * 2048, 0);
          for (int mt = mt0; mt < mt1; ++mt) {
            const int t0 = 2 * mt, t1 = 2 * mt + 1;
            // block addresses: a tap's window at this k-step, or the (k-step independent) ones tile
            const uint32_t a0 = t0 < p.ntaps ? a_tile + (uint32_t)p.delta[t0] * 128u + ks * 2048 : ones_smem;
            const uint32_t a1 = t1 < p.ntaps ? a_tile + (uint32_t)p.delta[t1] * 128u + ks * 2048 : ones_smem;
            // LBO = byte distance from block 0 to block 1 (the ones tile lies behind every stage, so it is positive; two
            // blocks on the ones tile -- only when ntaps is even and this is the last M-tile -- use a dummy 128)
            const uint32_t lbo = a1 > a0 ? a1 - a0 : 128u;
            umma_bf16(tmem_d + mt * 64, make_desc(a0, lbo), bd, IDESC, (first && ks == 0) ? 0u : 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        first = false;
        if (++s == ST) { s = 0; ph ^= 1; }
      }
      umma_commit(smem_u32(&done_bar));
    }
    __syncwarp();
  } else if (warp < 8) {
    // dump: warp w <-> TMEM lanes 32*(w%4).., columns (w/4)*32..+32 of every M-tile
    const int lq = warp & 3, half = warp >> 2;
    mbar_wait_relaxed(smem_u32(&done_bar), 0);
    tc_fence_after();
    float* out = p.partial + (size_t)blockIdx.x * nmt * 128 * 64;
    for (int mt = 0; mt < nmt; ++mt) {
      uint32_t acc[32];
#pragma unroll
      for (int c = 0; c < 32; c += 16)
        tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + mt * 64 + half * 32 + c, *reinterpret_cast<uint32_t(*)[16]>(&acc[c]));
      tmem_ld_wait();
      float4* o4 = reinterpret_cast<float4*>(out + (size_t)(mt * 128 + lq * 32 + lane) * 64 + half * 32);
#pragma unroll
      for (int c = 0; c < 8; ++c) o4[c] = make_float4(__uint_as_float(acc[4 * c]), __uint_as_float(acc[4 * c + 1]), __uint_as_float(acc[4 * c + 2]), __uint_as_float(acc[4 * c + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tcols);
}


// ---- the same tile scheme with cp.async producers instead of two TMA boxes per tile.  A TMA box over 128-byte rows is one
// request per row (850 per conv1 tile, one every ~7-10 cycles: 0.92 ms for conv1's weight gradient), and it moves 128 bytes per
// pixel although a packed conv1 pixel has 96 and its dZ pixel 64.  Here the 8 warps that dump the accumulators at the end
// copy exactly the live 16-byte chunks of every pixel (6 + 4 per position for conv1 instead of the gather kernel's 28 per output
// pixel, which re-reads every source pixel once per tap) into the same 128-byte-swizzled raster tile; out-of-range positions
// are zero-filled (src-size 0), so the MMAs and the shifted MN-major descriptors are exactly those of the TMA version.
struct HaloWgradCpParams {
  HaloWgradParams g;
  const uint8_t* x;        // bf16 NHWC source [F, H, W, xpe]
  const uint8_t* dz;       // bf16 [F, OH, OW, zpe]
  int H, W, OH, OW;
  int lag;                 // copy groups a producer thread keeps in flight before it publishes the oldest
};

__device__ __forceinline__ void cp_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    default: cp_async_wait<3>(); break;
  }
}

constexpr int WGCP_MAX_ISS = 12;
constexpr int NT_WGCP = 256 + 32 * WGCP_MAX_ISS;     // 8 producer / dump warps + up to 12 MMA issuer warps

template <int CX, int CZ>      // 16-byte chunks per source / dZ pixel (6, 4: conv1 over packed frames; 8, 8: conv3)
__global__ void __launch_bounds__(NT_WGCP, 1) conv_halo_wgrad_cp_kernel(const __grid_constant__ HaloWgradCpParams q) {
  const HaloWgradParams& p = q.g;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_ST], empty_bar[MAX_ST], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ST = (uint32_t)p.stages;
  const uint32_t stage_bytes = (uint32_t)(p.a_bytes + p.b_bytes);
  const uint32_t ones_smem = base + ST * stage_bytes;           // 32 rows x 128 B of bf16 1.0
  const int nmt = p.nmt;
  // issuers: one per M-tile, times two halves of a tile's k-steps when TMEM holds a second accumulator set (added in the dump)
  // (an accumulator is NACC = the live dZ channels wide: 32 columns for conv1, so four sets of its 3 M-tiles fit in 384 columns)
  constexpr int NACC = CZ * 8;
  int ksplit = 1;
  while (2 * ksplit * nmt * NACC <= 512 && 2 * ksplit * nmt <= WGCP_MAX_ISS && 2 * ksplit <= 4 && 2 * ksplit <= (p.BH * p.PW) / 16) ksplit *= 2;
  const int ni_m = nmt < WGCP_MAX_ISS / ksplit ? nmt : WGCP_MAX_ISS / ksplit;
  const int niss = ni_m * ksplit;
  const uint32_t need_cols = (uint32_t)(ksplit * nmt * NACC);
  const uint32_t tcols = need_cols <= 64 ? 64u : (need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u));
  constexpr int NPROD = 256;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), tcols);
  if (tid == 32) {
    for (uint32_t s = 0; s < ST; ++s) { mbar_init(smem_u32(&full_bar[s]), NPROD); mbar_init(smem_u32(&empty_bar[s]), (uint32_t)niss); }
    mbar_init(smem_u32(&done_bar), (uint32_t)niss);
    mbar_fence_init();
  }
  // zero all stages once (chunks beyond a pixel's live channels and the rows behind a tile are read by the MMAs: they must be
  // finite; nobody writes them afterwards), then the ones tile
  for (uint32_t o = tid * 16; o < ST * stage_bytes; o += NT_WGCP * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
  for (uint32_t o = tid * 16; o < 32 * 128; o += NT_WGCP * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_smem + o), "r"(0x3F803F80u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int KL = p.BH * p.PW;                                    // contraction positions per tile (multiple of 16)

  if (warp < 8) {
    // ---------------------------------------------------------------- producers (then dump, below)
    // The (row, chunk) positions a thread copies are the same for every tile: their shared-memory offsets and their global
    // offsets relative to the tile's first source row are computed ONCE (ncu r02: with the index arithmetic inside the tile loop the
    // producers were issue-bound -- 51 % of the issue slots, MMA issuers starving); per tile a copy is one 64-bit add + cp.async.
    // Positions right of the frame (j >= W) are never copied: they stay zero from the initial fill.
    constexpr int MAXS = CX == 6 ? 8 : 16;       // copies per thread and tile and operand (conv1 tiles: 7 + 4)
    uint32_t xg[MAXS], xs_[MAXS], zg[MAXS], zs_[MAXS];      // global byte offset | smem offset (low 17 bits) + tile row (high bits)
    const int xrows = p.BH + p.KH - 1;
    const long long xrow_b = (long long)q.W * CX * 16, xfr_b = (long long)q.H * xrow_b;
    const long long zrow_b = (long long)q.OW * CZ * 16, zfr_b = (long long)q.OH * zrow_b;
    const int nxc = xrows * p.PW * CX, nzc = p.BH * p.PW * CZ;
#pragma unroll
    for (int k = 0; k < MAXS; ++k) {
      const int c = tid + NPROD * k;
      xg[k] = 0xffffffffu; zg[k] = 0xffffffffu; xs_[k] = 0; zs_[k] = 0;
      if (c < nxc) {
        const int i = c / (p.PW * CX), idx = c - i * (p.PW * CX), j = idx / CX, ch = idx - j * CX;
        if (j < q.W) { xg[k] = (uint32_t)(i * xrow_b + idx * 16); xs_[k] = swz128(i * p.PW + j, ch) | ((uint32_t)i << 17); }
      }
      if (c < nzc) {
        const int i = c / (p.PW * CZ), idx = c - i * (p.PW * CZ), j = idx / CZ, ch = idx - j * CZ;
        if (j < q.OW) { zg[k] = (uint32_t)(i * zrow_b + idx * 16); zs_[k] = swz128(i * p.PW + j, ch) | ((uint32_t)i << 17); }
      }
    }
    uint32_t s = 0, ph = 1, ps = 0;
    int inflight = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int f = tile / p.tiles_per_frame, ti = tile - f * p.tiles_per_frame;
      const int y0 = ti * p.BH;
      mbar_wait(smem_u32(&empty_bar[s]), ph);
      const uint32_t a_tile = base + s * stage_bytes, b_tile = a_tile + (uint32_t)p.a_bytes;
      const uint8_t* xb = q.x + f * xfr_b + (long long)y0 * xrow_b;
      const uint8_t* zb = q.dz + f * zfr_b + (long long)y0 * zrow_b;
      const int xlim = q.H - y0, zlim = q.OH - y0;          // tile rows i >= lim lie below the frame: zero-filled
#pragma unroll
      for (int k = 0; k < MAXS; ++k) {
        if (xg[k] != 0xffffffffu) {
          const bool ok = (int)(xs_[k] >> 17) < xlim;
          cp_async16(a_tile + (xs_[k] & 0x1ffffu), ok ? xb + xg[k] : q.x, ok ? 16u : 0u);
        }
      }
#pragma unroll
      for (int k = 0; k < MAXS; ++k) {
        if (zg[k] != 0xffffffffu) {
          const bool ok = (int)(zs_[k] >> 17) < zlim;
          cp_async16(b_tile + (zs_[k] & 0x1ffffu), ok ? zb + zg[k] : q.dz, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (++s == ST) { s = 0; ph ^= 1; }
      if (inflight == q.lag) {
        cp_wait_dyn(q.lag);
        fence_proxy_async();
        mbar_arrive(smem_u32(&full_bar[ps]));
        if (++ps == ST) ps = 0;
      } else {
        ++inflight;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; inflight > 0; --inflight) {
      mbar_arrive(smem_u32(&full_bar[ps]));
      if (++ps == ST) ps = 0;
    }
    // ---------------------------------------------------------------- dump: warp w <-> TMEM lanes 32*(w%4).., columns (w/4)*32..+32
    const int lq = warp & 3, half = warp >> 2;
    mbar_wait_relaxed(smem_u32(&done_bar), 0);
    tc_fence_after();
    float* out = p.partial + (size_t)blockIdx.x * nmt * 128 * 64;
    constexpr int HC = NACC / 2;                       // columns per warp half: 16 or 32
    for (int mt = 0; mt < nmt; ++mt) {
      float acc[HC];
#pragma unroll
      for (int c = 0; c < HC; ++c) acc[c] = 0.f;
      for (int a = 0; a < ksplit; ++a) {               // add the accumulator sets of the k-step groups
        uint32_t r[HC];
#pragma unroll
        for (int c = 0; c < HC; c += 16)
          tmem_ld16_nowait(tmem_d + ((uint32_t)(lq * 32) << 16) + (a * nmt + mt) * NACC + half * HC + c, *reinterpret_cast<uint32_t(*)[16]>(&r[c]));
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HC; ++c) acc[c] += __uint_as_float(r[c]);
      }
      float4* o4 = reinterpret_cast<float4*>(out + (size_t)(mt * 128 + lq * 32 + lane) * 64 + half * HC);
#pragma unroll
      for (int c = 0; c < HC / 4; ++c) o4[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    }
  } else if (warp >= 8 && warp < 8 + niss) {
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(128, CZ * 8, true, true);   // N = the live dZ channels (32 for conv1: half the MMA work)
      const int me = warp - 8;
      const int im = me % ni_m, ik = me / ni_m;
      const uint32_t acc0 = tmem_d + (uint32_t)(ik * nmt * NACC);
      uint32_t s = 0, ph = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t a_tile = base + s * stage_bytes, b_tile = a_tile + (uint32_t)p.a_bytes;
        for (int ks = ik; ks < KL / 16; ks += ksplit) {
          const uint64_t bd = make_desc(b_tile + ks * 2048, 0);
          for (int mt = im; mt < nmt; mt += ni_m) {
            const int t0 = 2 * mt, t1 = 2 * mt + 1;
            const uint32_t a0 = t0 < p.ntaps ? a_tile + (uint32_t)p.delta[t0] * 128u + ks * 2048 : ones_smem;
            const uint32_t a1 = t1 < p.ntaps ? a_tile + (uint32_t)p.delta[t1] * 128u + ks * 2048 : ones_smem;
            const uint32_t lbo = a1 > a0 ? a1 - a0 : 128u;
            umma_bf16(acc0 + mt * NACC, make_desc(a0, lbo), bd, IDESC, (first && ks == ik) ? 0u : 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        first = false;
        if (++s == ST) { s = 0; ph ^= 1; }
      }
      umma_commit(smem_u32(&done_bar));
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tcols);
}

}  // namespace

// Returns HULC2_ENOTIMPL when the shape does not fit.  On success *grid_out CTAs wrote partial[g][nblk*64][64] with
// nblk = 2 * nmt blocks of 64 rows: rows tap * 64 + ch, and row ntaps * 64 = column sums of dZ.
int hulc2_conv_halo_wgrad(const void* x, int xpe, const void* dz, int zpe, int F, int H, int W, int KH, int KW, float* partial,
                          long long partial_bytes, int* grid_out, int* nblk_out, cudaStream_t st) {
  EncodeTiledFn encode = halo_encode_fn();
  if (!encode || KH * KW > 15 || xpe > 64 || zpe > 64 || (xpe * 2) % 16 || (zpe * 2) % 16) return HULC2_ENOTIMPL;
  if (((uintptr_t)x & 15) || ((uintptr_t)dz & 15)) return HULC2_ENOTIMPL;
  const int OH = H - KH + 1, OW = W - KW + 1;
  HaloWgradParams p{};
  p.KH = KH; p.ntaps = KH * KW; p.nmt = (p.ntaps + 2) / 2;
  if (p.nmt > 8) return HULC2_ENOTIMPL;
  // raster pitch >= W and rows per tile such that BH * PW is a multiple of 16 and 2+ stages fit; prefer more rows per tile
  int best_bh = 0, best_pw = 0;
  for (int pw = W; pw <= W + 3 && pw <= 256; ++pw)
    for (int bh = 1; bh <= 16 && bh <= OH + 15; ++bh) {
      if ((bh * pw) % 16) continue;
      const int a_rows = (bh + KH - 1) * pw + 16, b_rows = bh * pw;
      const long long stage = (((long long)a_rows * 128 + 1023) & ~1023LL) + (((long long)b_rows * 128 + 1023) & ~1023LL);
      if (bh + KH - 1 > 256 || 2 * stage + 8192 > 227 * 1024 - 2048) continue;
      // efficiency: useful rows per tile row budget
      const int tiles = (OH + bh - 1) / bh;
      const double eff = (double)OH / (tiles * bh) * (double)W / pw;
      const int btiles = best_bh ? (OH + best_bh - 1) / best_bh : 1;
      const double beff = best_bh ? (double)OH / (btiles * best_bh) * (double)W / best_pw : 0.0;
      if (eff > beff + 1e-9 || (eff > beff - 1e-9 && bh > best_bh)) { best_bh = bh; best_pw = pw; }
    }
  if (!best_bh) return HULC2_ENOTIMPL;
  p.BH = best_bh; p.PW = best_pw;
  for (int t = 0; t < p.ntaps; ++t) p.delta[t] = (short)((t / KW) * p.PW + t % KW);
  const int a_rows = (p.BH + KH - 1) * p.PW + 16, b_rows = p.BH * p.PW;
  p.a_bytes = ((a_rows * 128) + 1023) & ~1023;
  p.b_bytes = ((b_rows * 128) + 1023) & ~1023;
  int stages = (227 * 1024 - 2048 - 8192) / (p.a_bytes + p.b_bytes);
  if (stages > 4) stages = 4;
  if (stages < 2) return HULC2_ENOTIMPL;
  p.stages = stages;
  p.F = F; p.tiles_per_frame = hulc2_cdiv(OH, p.BH); p.ntiles = F * p.tiles_per_frame;
  const int grid = p.ntiles < halo_sms() ? p.ntiles : halo_sms();
  const long long need = (long long)grid * p.nmt * 128 * 64 * sizeof(float);
  if (!partial || partial_bytes < need) return HULC2_ENOTIMPL;
  p.partial = partial;

  static const int variant = getenv("HULC2_WGRAD_HALO") ? atoi(getenv("HULC2_WGRAD_HALO")) : 2;   // 1 = TMA boxes, else cp.async producers
  if (variant != 1 && ((xpe == 48 && zpe == 32) || (xpe == 64 && zpe == 64))) {
    const int cxh = xpe / 8, czh = zpe / 8;
    if ((p.BH + KH - 1) * p.PW * cxh > 16 * 256 || p.BH * p.PW * czh > 16 * 256 || (long long)(p.BH + KH) * W * xpe * 2 > 0x7fffffffLL)
      return HULC2_ENOTIMPL;                      // more than 16 copies per producer thread and tile: the gather kernel takes it
    HaloWgradCpParams q{};
    q.g = p; q.x = (const uint8_t*)x; q.dz = (const uint8_t*)dz; q.H = H; q.W = W; q.OH = OH; q.OW = OW;
    // a producer publishes tile n after issuing tile n + lag; lag = stages - 2 leaves one free slot, so publishing tile n + 1 never
    // waits behind the MMAs of tile n (with lag = stages - 1 the copy issue and the MMAs of consecutive tiles serialise)
    q.lag = stages >= 2 ? (stages - 2 < 3 ? stages - 2 : 3) : 0;
    const int smem = stages * (p.a_bytes + p.b_bytes) + 32 * 128 + 1024;
    static int configured_cp = 0;
    if (configured_cp < smem) {
      if (cudaFuncSetAttribute(conv_halo_wgrad_cp_kernel<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
          cudaFuncSetAttribute(conv_halo_wgrad_cp_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        cudaGetLastError();
        return HULC2_ENOTIMPL;
      }
      configured_cp = smem;
    }
    if (xpe == 48) conv_halo_wgrad_cp_kernel<6, 4><<<grid, NT_WGCP, smem, st>>>(q);
    else conv_halo_wgrad_cp_kernel<8, 8><<<grid, NT_WGCP, smem, st>>>(q);
    HULC2_CHECK_LAUNCH();
    *grid_out = grid;
    *nblk_out = 2 * p.nmt;
    return HULC2_OK;
  }
  CUtensorMap tmx, tmz;
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  {
    const cuuint64_t pb = (cuuint64_t)xpe * 2;
    cuuint64_t dims[4] = {64u, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t strides[3] = {pb, (cuuint64_t)W * pb, (cuuint64_t)H * W * pb};
    cuuint32_t box[4] = {64u, (cuuint32_t)p.PW, (cuuint32_t)(p.BH + KH - 1), 1u};
    if (encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return HULC2_ENOTIMPL;
  }
  {
    const cuuint64_t pb = (cuuint64_t)zpe * 2;
    cuuint64_t dims[4] = {64u, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)F};
    cuuint64_t strides[3] = {pb, (cuuint64_t)OW * pb, (cuuint64_t)OH * OW * pb};
    cuuint32_t box[4] = {64u, (cuuint32_t)p.PW, (cuuint32_t)p.BH, 1u};
    if (encode(&tmz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return HULC2_ENOTIMPL;
  }
  const int smem = stages * (p.a_bytes + p.b_bytes) + 32 * 128 + 1024;
  auto kern = conv_halo_wgrad_kernel<0>;
  static int configured = 0;
  if (configured < smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return HULC2_ENOTIMPL; }
    configured = smem;
  }
  kern<<<grid, NT_HALO, smem, st>>>(tmx, tmz, p);
  HULC2_CHECK_LAUNCH();
  *grid_out = grid;
  *nblk_out = 2 * p.nmt;
  return HULC2_OK;
}
