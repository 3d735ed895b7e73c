// Hardware probe (not product code): MN-major SWIZZLE_128B operands with shifted start addresses, for a halo-tile
// weight-gradient kernel.  Tile rows = contraction index k (pixel position), one 128-byte row = 64 M (or N) elements.
//   D[m, n] = sum_k A[k + delta(m / 64), m % 64] * B[k, n],   M = 128 = two 64-wide blocks (taps) whose row offsets differ:
//   block 0 starts at row d0, block 1 at row d1  ->  descriptor start = base + d0 * 128, LBO = (d1 - d0) * 128 bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hulc2_b200/csrc tools/probes/umma_mn_shift_probe.cu -o tools/probes/umma_mn_shift_probe
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "sm100.cuh"

using namespace sm100;

constexpr int ROWS = 320, KLEN = 64;   // contraction over 64 positions (4 UMMA_K steps)
__host__ __device__ inline float a_val(int r, int m) { return (float)(((r * 3 + m * 5) % 13) - 6); }
__host__ __device__ inline float b_val(int r, int n) { return (float)(((r + 2 * n) % 7) - 3); }

__global__ void __launch_bounds__(128) probe_kernel(float* out, const int* d0s, const int* d1s, int nd) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + ROWS * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int R = i >> 3, c = i & 7;
    __nv_bfloat16 v[8], w[8];
    for (int e = 0; e < 8; ++e) { v[e] = __float2bfloat16(a_val(R, c * 8 + e)); w[e] = __float2bfloat16(b_val(R, c * 8 + e)); }
    const uint4 u = *reinterpret_cast<uint4*>(v), x = *reinterpret_cast<uint4*>(w);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_smem + swz128(R, c)), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_smem + swz128(R, c)), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  if (tid == 32) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, 64, true, true);
  for (int d = 0; d < nd; ++d) {
    if (tid == 0) {
      const uint32_t lbo = (uint32_t)(d1s[d] - d0s[d]) * 128u;
      for (int ks = 0; ks < KLEN / 16; ++ks) {
        const uint64_t ad = make_desc(a_smem + d0s[d] * 128 + ks * 2048, lbo), bd = make_desc(b_smem + ks * 2048, 0);
        umma_bf16(tmem_d, ad, bd, IDESC, ks > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), d & 1);
    tc_fence_after();
    uint32_t acc[64];
    for (int c = 0; c < 64; c += 16) tmem_ld16_nowait(tmem_d + ((uint32_t)(warp * 32) << 16) + c, *reinterpret_cast<uint32_t(*)[16]>(&acc[c]));
    tmem_ld_wait();
    tc_fence_before();
    for (int n = 0; n < 64; ++n) out[((size_t)d * 128 + warp * 32 + lane) * 64 + n] = __uint_as_float(acc[n]);
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

int main() {
  std::vector<int> d0 = {0, 0, 0, 0, 50, 3, 8, 100}, d1 = {64, 1, 8, 50, 51, 77, 9, 101};   // first: the canonical layout (LBO = 8 KB)
  const int nd = (int)d0.size();
  int *g0, *g1; float* d_out;
  cudaMalloc(&g0, nd * 4); cudaMalloc(&g1, nd * 4);
  cudaMalloc(&d_out, (size_t)nd * 128 * 64 * 4);
  cudaMemcpy(g0, d0.data(), nd * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(g1, d1.data(), nd * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * ROWS * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(d_out, g0, g1, nd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out((size_t)nd * 128 * 64);
  cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
  for (int d = 0; d < nd; ++d) {
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        const int dl = m < 64 ? d0[d] : d1[d];
        float ref = 0.f;
        for (int k = 0; k < KLEN; ++k) ref += a_val(k + dl, m & 63) * b_val(k, n);
        if (out[((size_t)d * 128 + m) * 64 + n] != ref) ++bad;
      }
    printf("block offsets (%3d, %3d): %s (%d mismatches)\n", d0[d], d1[d], bad ? "FAIL" : "ok", bad);
  }
  return 0;
}
