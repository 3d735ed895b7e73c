// Cluster split-K persistent decoder recurrence for sm_100a, second generation (nn.RNN relu cell,
// hulc2/models/decoders/utils/rnn.py:5-14).  Same tiling as rnn_cluster_sm100.cu (described below); what changed is the per-step
// dependency chain, which is all this kernel's time (ncu r02: tensor pipe 5.7 %, issue slots 14 %, warps parked on barriers):
//   * the state slice of a step (128 rows x H/CS columns of bf16) arrives by TMA -- nkt 3-D boxes issued by one thread, each
//     completing its own mbarrier -- instead of 8192 16-byte cp.async + wait_group + proxy fence + __syncthreads per k-tile;
//     the MMA issuer (another warp) starts on a k-tile as soon as its box has landed;
//   * a CTA publishes its 16 columns per WARP (st bf16 -> __syncwarp -> fence -> atomicAdd by lane 0) instead of
//     __syncthreads + fence + one atomic per CTA, and the fp32 copy of the state (only the caller reads it) is stored AFTER the
//     flag, off the critical path;
//   * the bf16 state buffer has S + 1 slots (slot 0 = initial state, converted by the host-side launcher), so every step loads
//     its operand the same way.
//
//   forward : h[t]  = relu(pre[t] + h[t-1] W_hh^T)                     t = 0 .. S-1
//   backward: dz[t] = (dh[t] + dz[t+1] W_hh) * (h[t] > 0)              t = S-1 .. 0   (in place over dh)
//
// rnn_persistent_sm100.cu gives every CTA 16 output columns and the WHOLE previous state as its A operand
// (128 x H bf16 = 512 KB per CTA per step, 64 MB of L2->SM traffic per step over 128 CTAs): measured 17-19 us per
// step, all of it state broadcast + 32 block-wide syncs + the grid barrier.  Here the step is tiled in 2-D:
//   * a thread-block cluster of 8 CTAs owns 128 output columns; CTA rank r of the cluster owns K slice
//     [r H/8, (r+1) H/8) of the contraction.  Its W block (128 n x H/8 k, 64 KB bf16 at H = 2048) stays in shared
//     memory for the whole sequence; per step it reads only ITS slice of the state (128 x H/8 bf16 = 64 KB, 8 MB per
//     step over the grid, 8x less than before) with cp.async and issues H/128 tcgen05.mma 128x128x16;
//   * the 8 partial accumulators (TMEM, 128 x 128 fp32 each) are reduced through distributed shared memory:
//     every CTA pushes the 16-column strip that rank d finalises into d's shared memory (st.shared::cluster),
//     one cluster barrier, then rank d sums 8 strips locally and runs the epilogue (add, ReLU / mask, fp32 state +
//     bf16 operand copy) for its 16 columns;
//   * the grid barrier is split per K slice: the CTAs that consume slice r wait on counter[r], which only the
//     H/128 CTAs producing those columns increment (16 arrivals per counter instead of 128 on one address).
// H/128 clusters x 8 CTAs = 128 CTAs at H = 2048, all co-resident (checked with cudaOccupancyMaxActiveClusters).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "sm100.cuh"
#include "../../include/hulc2_b200.h"

namespace {

using namespace sm100;

// CS = CTAs per cluster = K slices (8, or 4 where the device cannot co-schedule H/128 clusters of 8: a B200 with 148 SMs
// takes 15).  Every CTA finalises NR = 16 output columns, so a cluster owns NC = 16 CS columns and the grid is H/16 CTAs.
constexpr int NR = 16;         // columns finalised per CTA
constexpr int BM = 128;        // UMMA M (rows >= B are zero)
constexpr int KT = 64;         // k per tile (128-byte swizzle row)
constexpr int NT = 256;
constexpr int MAX_H = 2048;
constexpr uint32_t A_TILE = BM * 128;               // 16 KB: 128 rows x 64 bf16
constexpr uint32_t CTR_STRIDE = 32;                 // counters 128 bytes apart (in u32)
constexpr long long CTR_BYTES = (MAX_H / KT) * CTR_STRIDE * 4;   // 4 KB at the head of the workspace: one counter per k-tile
template <int CS>
struct Geo {
  static constexpr int NC = NR * CS;                     // output columns per cluster (UMMA N)
  static constexpr int MAX_KT = MAX_H / (CS * KT);       // k tiles per slice (4 / 8)
  static constexpr uint32_t W_TILE = NC * 128;           // 16 KB / 8 KB
  static constexpr uint32_t RED_BYTES = CS * BM * NR * 4;
  static constexpr int SMEM = MAX_KT * (int)(W_TILE + A_TILE) + (int)RED_BYTES + 1024;   // 193 KB / 225 KB
};

struct ClusterRnnParams {
  const float* add;          // [S,B,H] pre-activations (fwd) / incoming gradients (bwd; aliases `out`)
  const float* w;            // [H,H] W_hh
  int has_init;              // fwd: slot 0 of outb holds the bf16 initial state (converted before the launch)
  const float* mask;         // [S,B,H] or null: output is zeroed where mask <= 0 (bwd: h)
  float* out;                // [S,B,H] fp32 states
  __nv_bfloat16* outb;       // [S+1,B,H] bf16 states (workspace): slot 0 = initial state, slot t+1 = state t
  float* final_out;          // bwd only: dh0 [B,H] = dz[0] W_hh (or null)
  unsigned int* counters;    // CS (kt_flags: H / 64) counters, CTR_STRIDE apart (zeroed before launch); one arrival per WARP per step
  int S, B, H;
  int relu, reverse, transpose_w;
  unsigned long long* trace; // diagnostic (hulc2_rnn_set_trace): clock64 stamps of CTA 0, 16 slots per step, or null
  int pair_rows;             // 1: the two column halves of a row are adjacent lanes (full-sector state stores)
  int kt_flags;              // 1: one flag per 64-column k-tile (4 producer CTAs) instead of one per K slice (KS / 16 producer CTAs)
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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void wait_pending(int n) {   // cp.async.wait_group needs an immediate
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

// Set when a flag wait gives up (co-residency of all clusters is guaranteed by the cooperative launch, so this only fires
// on a protocol bug or a dying peer).  The waiter then proceeds with whatever state is there -- every barrier of the
// step protocol is still matched, so the kernel terminates with wrong numbers instead of trapping the context; the host
// reads the flag through hulc2_rnn_device_error().
__device__ unsigned int g_cluster2_rnn_error = 0;

static unsigned long long* g_trace = nullptr;
#define RNN_STAMP(slot) do { if (p.trace && blockIdx.x == 0) p.trace[it * 16 + (slot)] = clock64(); } while (0)

template <int CS>
__global__ void __launch_bounds__(NT, 1) rnn_cluster2_kernel(const __grid_constant__ CUtensorMap tm_state, const ClusterRnnParams p) {
  using G = Geo<CS>;
  constexpr int NC = G::NC, MAX_KT = G::MAX_KT;
  constexpr uint32_t W_TILE = G::W_TILE;
  constexpr int DPT = CS / 2;            // destination ranks each thread pushes to
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ __align__(8) uint64_t full_bar[Geo<CS>::MAX_KT];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, B = p.B, S = p.S;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / CS;
  const int KS = H / CS;                  // k extent of this CTA's slice
  const int nkt = KS / KT;                // 1..4
  const int k0 = (int)rank * KS;
  const int n0 = cluster * NC;            // first output column of the cluster
  const int nf = n0 + (int)rank * NR;     // first output column this CTA finalises
  const int producers = KS / NR;          // CTAs whose columns form one K slice of the next step
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_smem = base;                          // MAX_KT tiles of NC rows
  const uint32_t a_smem = base + MAX_KT * W_TILE;        // MAX_KT tiles of 128 rows (1024-aligned: W_TILE is a multiple of 1024)
  const uint32_t red_smem = a_smem + MAX_KT * A_TILE;    // [src rank][16-byte chunk q][row][4 floats]

  // two MMA issuer threads split the k-tiles of a step and accumulate into their own TMEM columns (one thread issues a
  // tcgen05.mma only every ~110 cycles; the partial sums are added when the strips are read for the DSMEM push)
  constexpr int MAXI = (4 * NC <= 512) ? 4 : 2;            // TMEM: issuers x NC columns
  const int issuers = nkt >= MAXI ? MAXI : (nkt >= 2 ? 2 : 1);
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)(MAXI * NC));
  if (tid == 32) {
    mbar_init(smem_u32(&mma_done), (uint32_t)issuers);
    for (int kt = 0; kt < MAX_KT; ++kt) mbar_init(smem_u32(&full_bar[kt]), 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_state) : "memory");
  }
  // resident W block: element (n, k) -> fwd W[(n0+n)*H + k0+k], bwd W[(k0+k)*H + n0+n]; 16-byte chunks of 8 k
  for (int ch = tid; ch < NC * (KS / 8); ch += NT) {
    int n, kc;
    if (p.transpose_w) { n = ch % NC; kc = ch / NC; }        // consecutive threads -> consecutive n (coalesced)
    else { kc = ch % (KS / 8); n = ch / (KS / 8); }          // consecutive threads -> consecutive k
    const int k = kc * 8;
    float f[8];
    if (p.transpose_w) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __ldg(p.w + (long long)(k0 + k + e) * H + n0 + n);
    } else {
      const float4* s4 = reinterpret_cast<const float4*>(p.w + (long long)(n0 + n) * H + k0 + k);
      float4 a = __ldg(s4), b = __ldg(s4 + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    uint32_t dst = w_smem + (uint32_t)(k / KT) * W_TILE + swz128(n, (k % KT) / 8);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(f[0], f[1])), "r"(pack_bf16x2(f[2], f[3])),
                 "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7])) : "memory");
  }
  // (rows >= B of the A tiles are zero-filled by TMA: the box is 128 rows, the tensor has B)
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t IDESC = make_idesc(BM, NC, false, false);
  // every CTA of the cluster is running (its shared memory is live) before the first remote store
  cluster_arrive();
  cluster_wait();

  uint32_t mma_uses = 0;          // steps that ran the contraction so far (parity of mma_done and of every full_bar)
  unsigned int generation = 0;
  bool readers_pending = false;      // a cluster barrier phase "strips consumed" has been arrived on but not waited for
  const long long step = (long long)B * H;
  // epilogue ownership: thread -> row tid >> 1, 8 of the CTA's 16 columns (half = tid & 1): the two halves of a row are adjacent
  // lanes, so the 16-byte bf16 pieces they publish form ONE full 32-byte sector per instruction (with half = tid >> 7 every state
  // store was a partial-sector write from two different warps)
  const int erow = p.pair_rows ? tid >> 1 : tid & 127, ehalf = p.pair_rows ? tid & 1 : tid >> 7;
  const bool eactive = erow < B;
  const int ecol = nf + 8 * ehalf;
  // push ownership: TMEM lane = (warp & 3) * 32 + lane, columns (NC/2) * (warp >> 2) .. + NC/2 - 1  -> destination ranks
  // DPT * (warp >> 2) .. + DPT - 1
  const int prow = (warp & 3) * 32 + lane;
  const int phalf = warp >> 2;
  uint32_t remote[DPT];
#pragma unroll
  for (int j = 0; j < DPT; ++j) remote[j] = mapa(red_smem + rank * (BM * NR * 4) + (uint32_t)prow * 16, (uint32_t)(phalf * DPT + j));
  constexpr unsigned int WARPS = NT / 32;

  for (int it = 0; it <= S; ++it) {
    // it < S: produce state t;  it == S (bwd with final_out only): dh0 = dz[0] W_hh, no add/mask
    const bool final_pass = it == S;
    if (final_pass && !p.final_out) break;
    const int t = p.reverse ? S - 1 - it : it;
    const int tprev = p.reverse ? t + 1 : t - 1;
    const bool has_prev = final_pass ? true : (it > 0 || p.has_init != 0);
    // slot of the operand state: slot 0 = initial state, slot t + 1 = state t; the final pass reads dz[0] = slot 1
    const int slot_prev = final_pass ? 1 : (it > 0 ? tprev + 1 : 0);

    if (tid == 0) RNN_STAMP(0);
    // prefetch the epilogue addend / mask for this thread's 8 outputs (overlaps the flag wait and the k loop)
    float addv[8], mk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { addv[j] = 0.f; mk[j] = 1.f; }
    if (eactive && !final_pass) {
      const float4* a4 = reinterpret_cast<const float4*>(p.add + (long long)t * step + (long long)erow * H + ecol);
      float4 v0 = __ldcg(a4), v1 = __ldcg(a4 + 1);
      addv[0] = v0.x; addv[1] = v0.y; addv[2] = v0.z; addv[3] = v0.w; addv[4] = v1.x; addv[5] = v1.y; addv[6] = v1.z; addv[7] = v1.w;
      if (p.mask) {
        const float4* m4 = reinterpret_cast<const float4*>(p.mask + (long long)t * step + (long long)erow * H + ecol);
        float4 q0 = __ldcg(m4), q1 = __ldcg(m4 + 1);
        mk[0] = q0.x; mk[1] = q0.y; mk[2] = q0.z; mk[3] = q0.w; mk[4] = q1.x; mk[5] = q1.y; mk[6] = q1.z; mk[7] = q1.w;
      }
    }

    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;

    if (has_prev) {
      const uint32_t par = mma_uses & 1;
      if (warp == 0) {
        if (p.kt_flags) {
          // One flag per k-tile: lane kt waits for the 4 CTAs that produce columns [k0 + 64 kt, + 64) of the previous state and
          // issues that box itself, so the boxes of early producers are in flight while the slowest CTA of the slice still publishes.
          if (lane < nkt) {
            if (it > 0 || final_pass) {
              const unsigned int target = generation * (unsigned int)(KT / NR) * WARPS;
              const unsigned int* c = p.counters + ((int)rank * nkt + lane) * CTR_STRIDE;
              unsigned int v, spins = 0;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
                if ((++spins & 1023u) == 0) {
                  if (spins > (1u << 23)) atomicExch(&g_cluster2_rnn_error, 1u);
                  if (*reinterpret_cast<volatile unsigned int*>(&g_cluster2_rnn_error)) break;
                }
              } while (v < target);
              asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            if (lane == 0) RNN_STAMP(1);
            if (lane == nkt - 1) RNN_STAMP(2);
            const uint32_t bar = smem_u32(&full_bar[lane]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(A_TILE) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(a_smem + lane * A_TILE), "l"((uint64_t)&tm_state), "r"(bar), "r"(k0 + lane * KT), "r"(0), "r"(slot_prev) : "memory");
          }
        } else if (lane == 0) {
          if (it > 0 || final_pass) {
            // K slice `rank` of the previous state is complete: every warp of its `producers` CTAs has arrived `generation` times
            const unsigned int target = generation * (unsigned int)producers * WARPS;
            const unsigned int* c = p.counters + rank * CTR_STRIDE;
            unsigned int v, spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
              if ((++spins & 1023u) == 0) {          // never hang the GPU, never kill the context: flag the error and go on
                if (spins > (1u << 23)) atomicExch(&g_cluster2_rnn_error, 1u);
                if (*reinterpret_cast<volatile unsigned int*>(&g_cluster2_rnn_error)) break;
              }
            } while (v < target);
            // the state was written with generic-proxy stores by other SMs; the boxes below are read by the async proxy
            asm volatile("fence.proxy.async.global;" ::: "memory");
          }
          for (int kt = 0; kt < nkt; ++kt) {
            const uint32_t bar = smem_u32(&full_bar[kt]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(A_TILE) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(a_smem + kt * A_TILE), "l"((uint64_t)&tm_state), "r"(bar), "r"(k0 + kt * KT), "r"(0), "r"(slot_prev) : "memory");
          }
        }
        __syncwarp();
      } else if (warp - 1 < issuers) {
        if (lane == 0) {
          const int per = nkt / issuers;                       // nkt is a multiple of `issuers` (1, 2, 4 or 8 k-tiles)
          const int kt_lo = (warp - 1) * per, kt_hi = kt_lo + per;
          const uint32_t acc_t = tmem_d + (uint32_t)((warp - 1) * NC);
          for (int kt = kt_lo; kt < kt_hi; ++kt) {
            mbar_wait(smem_u32(&full_bar[kt]), par);
            if (kt == 0) RNN_STAMP(3);
            if (kt == nkt - 1) RNN_STAMP(4);
            tc_fence_after();
            const uint64_t ad = make_desc(a_smem + kt * A_TILE, 0), bd = make_desc(w_smem + kt * W_TILE, 0);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k) umma_bf16(acc_t, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC, (kt > kt_lo || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&mma_done));
          if (warp == issuers) RNN_STAMP(5);
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&mma_done), par);
      if (tid == 0) RNN_STAMP(6);
      mma_uses += 1;
      tc_fence_after();

      // push: this CTA's partial strip for destination rank d -> d's RED[rank][q][row][4].  The accumulators of all issuers are
      // read behind ONE tcgen05.wait::ld (they used to be read and added issuer by issuer: three more TMEM round trips per step)
      constexpr bool ONE_WAIT = MAXI * DPT * 16 <= 128;        // registers: clusters of 4 (the B200 configuration) yes, of 8 no
      constexpr int NA = ONE_WAIT ? MAXI : 1;
      uint32_t ra[NA][DPT][16];
      uint32_t (&r)[DPT][16] = ra[0];
      if (ONE_WAIT) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
          if (a < issuers) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) tmem_ld16_nowait(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * NC + (phalf * DPT + j) * NR), ra[a][j]);
          }
        tmem_ld_wait();
#pragma unroll
        for (int a = 1; a < NA; ++a)
          if (a < issuers) {
#pragma unroll
            for (int j = 0; j < DPT; ++j)
#pragma unroll
              for (int e = 0; e < 16; ++e) r[j][e] = __float_as_uint(__uint_as_float(r[j][e]) + __uint_as_float(ra[a][j][e]));
          }
      } else {
#pragma unroll
        for (int j = 0; j < DPT; ++j) tmem_ld16_nowait(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((phalf * DPT + j) * NR), r[j]);
        tmem_ld_wait();
        for (int a = 1; a < issuers; ++a) {            // add the other issuers' partial accumulators
          uint32_t r2[DPT][16];
#pragma unroll
          for (int j = 0; j < DPT; ++j) tmem_ld16_nowait(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * NC + (phalf * DPT + j) * NR), r2[j]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < DPT; ++j)
#pragma unroll
            for (int e = 0; e < 16; ++e) r[j][e] = __float_as_uint(__uint_as_float(r[j][e]) + __uint_as_float(r2[j][e]));
        }
      }
      tc_fence_before();
      if (tid == 0) RNN_STAMP(7);
      if (readers_pending) cluster_wait();      // every CTA of the cluster has consumed the strips of the previous step
      if (tid == 0) RNN_STAMP(11);
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) st_cluster_v4(remote[j] + q * (BM * 16), r[j][4 * q], r[j][4 * q + 1], r[j][4 * q + 2], r[j][4 * q + 3]);
      if (tid == 0) RNN_STAMP(12);
      cluster_arrive();
      if (tid == 0) RNN_STAMP(13);
      cluster_wait();
      if (tid == 0) RNN_STAMP(8);
      // local reduce of the CS strips
#pragma unroll
      for (int src = 0; src < CS; ++src) {
        const uint32_t a = red_smem + src * (BM * NR * 4) + (uint32_t)(2 * ehalf) * (BM * 16) + (uint32_t)erow * 16;
        float4 x, y;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(a + BM * 16) : "memory");
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w; acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
      }
      if (tid == 0) RNN_STAMP(14);
      cluster_arrive();                          // strips consumed (waited for before the next push)
      if (tid == 0) RNN_STAMP(15);
      readers_pending = true;
    }

    // epilogue
    if (final_pass) {
      if (eactive) {
        float4* o4 = reinterpret_cast<float4*>(p.final_out + (long long)erow * H + ecol);
        o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = acc[j] + addv[j];
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.mask) x = mk[j] > 0.f ? x : 0.f;
        v[j] = x;
      }
      // 1) the bf16 operand copy the next step's consumers wait for, 2) publish (per warp), 3) the fp32 state for the caller
      if (eactive) {
        uint4* b4 = reinterpret_cast<uint4*>(p.outb + (long long)(t + 1) * step + (long long)erow * H + ecol);
        b4[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
      generation += 1;
      if (tid == 0) RNN_STAMP(9);
      __syncwarp();
      if (lane == 0)   // release: this warp's share of the CTA's 16 columns of K slice nf / KS is visible before the count moves
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters + (p.kt_flags ? nf / KT : nf / KS) * CTR_STRIDE) : "memory");
      if (tid == 0) RNN_STAMP(10);
      if (eactive) {
        float4* o4 = reinterpret_cast<float4*>(p.out + (long long)t * step + (long long)erow * H + ecol);
        o4[0] = make_float4(v[0], v[1], v[2], v[3]);
        o4[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  if (readers_pending) cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)(MAXI * NC));
}

static bool rnn_cooperative() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HULC2_RNN_COOP"); v = (e && e[0] == '0') ? 0 : 1; }   // A/B switch, read once
  return v == 1;
}

static bool rnn_kt_flags() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HULC2_RNN_KT_FLAGS"); v = (e && e[0] == '0') ? 0 : 1; }   // A/B switch, read once
  return v == 1;
}

template <int CS>
static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, cudaStream_t st) {
  memset(cfg, 0, sizeof(*cfg));
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  // cooperative: the launch only succeeds if EVERY cluster of the grid is co-resident (the per-slice flags are a grid-wide
  // dependency); under concurrent work that takes SMs away the launch fails cleanly and the caller falls back
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg->gridDim = dim3(grid); cfg->blockDim = dim3(NT); cfg->dynamicSmemBytes = Geo<CS>::SMEM; cfg->stream = st;
  cfg->attrs = attr; cfg->numAttrs = rnn_cooperative() ? 2 : 1;
}

// how many CS-CTA clusters of the kernel can be co-resident on this device (one device per process); 0 = unusable
template <int CS>
static int cluster_capacity() {
  // per device (a process may drive several): index = cudaGetDevice()
  static int cap[64];
  static bool known[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int& max_clusters = cap[dev];
  if (!known[dev]) {
    known[dev] = true;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    cluster_config<CS>(&cfg, attr, MAX_H / NR, 0);
    cudaError_t e = cudaFuncSetAttribute(rnn_cluster2_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<CS>::SMEM);
    int n = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, rnn_cluster2_kernel<CS>, &cfg);
    if (e != cudaSuccess) {
      hulc2_set_error(cudaGetErrorString(e));
      cudaGetLastError();
      n = 0;
    }
    max_clusters = n;
  }
  return max_clusters;
}

template <int CS>
static int cluster_launch(const CUtensorMap& tm, const ClusterRnnParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  cluster_config<CS>(&cfg, attr, p.H / NR, st);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, rnn_cluster2_kernel<CS>, tm, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    // not all clusters can be co-resident right now (SMs held by concurrent work): let the caller take the next kernel
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) return HULC2_ENOTIMPL;
    hulc2_set_error(cudaGetErrorString(e));
    return HULC2_ELAUNCH;
  }
  return HULC2_OK;
}

__global__ void init_state_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// cuTensorMapEncodeTiled comes from the driver at run time, so the library loads on hosts without libcuda (CPU test container)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

}  // namespace

// diagnostic hook (tools/trace_rnn.py): device buffer of (S + 1) * 16 u64 that receives clock64 stamps of CTA 0, or null
extern "C" int hulc2_rnn_set_trace(void* buf) { g_trace = reinterpret_cast<unsigned long long*>(buf); return HULC2_OK; }

int hulc2_rnn_cluster2_device_error(int clear) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_cluster2_rnn_error, sizeof(v)) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (clear && v) { const unsigned int z = 0; cudaMemcpyToSymbol(g_cluster2_rnn_error, &z, sizeof(z)); }
  return (int)v;
}

// returns HULC2_ENOTIMPL when the shape / device does not fit this kernel (the caller tries rnn_cluster_sm100.cu, then the 1-D
// persistent kernel, then one GEMM per step)
extern int g_rnn_v2_reject;   // 1 shape, 2 workspace, 3 alignment, 4 no driver entry point, 5 clusters not co-resident, 6 tensor map, 7 launch
int hulc2_rnn_cluster2_launch(const float* add, const float* w, const float* init, const float* mask, float* out, float* final_out,
                              int S, int B, int H, int relu, int reverse, int transpose_w, void* workspace, long long workspace_bytes,
                              void* states16, cudaStream_t st) {
  // states16 (optional): caller-owned bf16 [S+1, B, H] that takes the place of the workspace's state buffer -- slot t + 1 = state t
  // is exactly the operand mirror the caller's contractions over the states need, so no separate conversion pass
  g_rnn_v2_reject = 0;
  if (B > BM || B <= 0 || S <= 0 || H > MAX_H || H % 512 != 0) { g_rnn_v2_reject = 1; return HULC2_ENOTIMPL; }
  const long long need = states16 ? CTR_BYTES : (long long)(S + 1) * B * H * 2 + CTR_BYTES;
  if (!workspace || workspace_bytes < need) { g_rnn_v2_reject = 2; return HULC2_ENOTIMPL; }
  if (((uintptr_t)add | (uintptr_t)out | (uintptr_t)w | (uintptr_t)workspace | (uintptr_t)(mask ? mask : out) | (uintptr_t)(init ? init : out) |
       (uintptr_t)(final_out ? final_out : out) | (uintptr_t)states16) & 15) {
    g_rnn_v2_reject = 3;
    return HULC2_ENOTIMPL;
  }
  EncodeTiledFn encode = encode_fn();
  if (!encode) { g_rnn_v2_reject = 4; return HULC2_ENOTIMPL; }
  // the per-slice flags need every cluster co-resident: H / (16 CS) clusters of CS CTAs
  int cs = 0;
  if (cluster_capacity<8>() >= H / (NR * 8)) cs = 8;
  else if (cluster_capacity<4>() >= H / (NR * 4)) cs = 4;
  if (!cs) { g_rnn_v2_reject = 5; return HULC2_ENOTIMPL; }
  ClusterRnnParams p;
  p.add = add; p.w = w; p.has_init = init ? 1 : 0; p.mask = mask; p.out = out; p.final_out = final_out;
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.outb = states16 ? reinterpret_cast<__nv_bfloat16*>(states16) : reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + CTR_BYTES);
  p.S = S; p.B = B; p.H = H; p.relu = relu; p.reverse = reverse; p.transpose_w = transpose_w;
  p.kt_flags = rnn_kt_flags() ? 1 : 0;
  p.trace = g_trace;
  { static int v = -1; if (v < 0) { const char* e = getenv("HULC2_RNN_PAIR_ROWS"); v = (e && e[0] == '0') ? 0 : 1; } p.pair_rows = v; }   // A/B switch
  // bf16 states [S+1, B, H]: one box = 64 k x 128 rows of one slot, 128-byte swizzled = one A k-tile; rows >= B read as zeros
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)(S + 1)};
  cuuint64_t strides[2] = {(cuuint64_t)H * 2, (cuuint64_t)B * H * 2};
  cuuint32_t box[3] = {(cuuint32_t)KT, (cuuint32_t)BM, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.outb, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    g_rnn_v2_reject = 6;
    return HULC2_ENOTIMPL;
  }
  if (cudaMemsetAsync(p.counters, 0, CTR_BYTES, st) != cudaSuccess) return HULC2_ELAUNCH;
  if (init) {
    const long long n = (long long)B * H;
    init_state_bf16_kernel<<<hulc2_cdiv(n / 4, 256), 256, 0, st>>>(init, p.outb, n);
    HULC2_CHECK_LAUNCH();
  }
  if (int e = (cs == 8 ? cluster_launch<8>(tm, p, st) : cluster_launch<4>(tm, p, st))) {
    if (e == HULC2_ENOTIMPL) g_rnn_v2_reject = 7;
    return e;
  }
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}
