// Hardware probe (not product code): does a SWIZZLE_128B K-major UMMA operand descriptor whose start address is
// base + delta * 128 B (delta NOT a multiple of the 8-row swizzle atom) read rows delta .. delta+127 of a tile that was
// written with the swizzle pattern anchored at the 1024-byte aligned base?  This is what a "halo tile + shifted
// descriptor" implicit-GEMM convolution needs: the im2col operand of tap (a, b) is the SAME
// shared-memory tile read at row offset a * pitch + b.
//   variant 0: descriptor base_offset field = 0;  variant 1: base_offset = delta & 7 (bits 49-51).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hulc2_b200/csrc tools/probes/umma_shift_probe.cu -o tools/probes/umma_shift_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "sm100.cuh"

using namespace sm100;

constexpr int ROWS = 256, N = 32, K = 64;
__host__ __device__ inline float a_val(int r, int k) { return (float)(((r * 3 + k * 5) % 13) - 6); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n + 2 * k) % 7) - 3); }

__global__ void __launch_bounds__(128) probe_kernel(float* out, const int* deltas, int nd, int variant, int mn_major) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + ROWS * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A tile: row R (128 bytes = 64 bf16 along k), chunk c at swz128(R, c)
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int R = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(a_val(R, c * 8 + e));
    const uint32_t dst = a_smem + swz128(R, c);
    const uint4 u = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(b_val(n, c * 8 + e));
    const uint32_t dst = b_smem + swz128(n, c);
    const uint4 u = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 32);
  if (tid == 32) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t IDESC = make_idesc(128, N, false, false);
  for (int d = 0; d < nd; ++d) {
    const int delta = deltas[d];
    if (tid == 0) {
      uint64_t ad = make_desc(a_smem + delta * 128, 0);
      if (variant == 1) ad |= (uint64_t)(delta & 7) << 49;
      const uint64_t bd = make_desc(b_smem, 0);
      for (int k = 0; k < K / 16; ++k) umma_bf16(tmem_d, ad + 2 * k, bd + 2 * k, IDESC, k > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), d & 1);
    tc_fence_after();
    uint32_t acc[32];
    tmem_ld16_nowait(tmem_d + ((uint32_t)(warp * 32) << 16), *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
    tmem_ld16_nowait(tmem_d + ((uint32_t)(warp * 32) << 16) + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
    tmem_ld_wait();
    tc_fence_before();
    for (int n = 0; n < N; ++n) out[((size_t)d * 128 + warp * 32 + lane) * N + n] = __uint_as_float(acc[n]);
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 32);
}

int main() {
  std::vector<int> deltas = {0, 8, 1, 2, 3, 5, 7, 9, 23, 24, 25, 46, 47, 48, 50, 100, 127};
  const int nd = (int)deltas.size();
  int* d_deltas; float* d_out;
  cudaMalloc(&d_deltas, nd * sizeof(int));
  cudaMalloc(&d_out, (size_t)nd * 128 * N * sizeof(float));
  cudaMemcpy(d_deltas, deltas.data(), nd * sizeof(int), cudaMemcpyHostToDevice);
  const int smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out((size_t)nd * 128 * N);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d_out, 0, out.size() * sizeof(float));
    probe_kernel<<<1, 128, smem>>>(d_out, d_deltas, nd, variant, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    printf("variant %d (base_offset = %s):\n", variant, variant ? "delta & 7" : "0");
    for (int d = 0; d < nd; ++d) {
      int bad = 0; int first_r = -1;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int k = 0; k < K; ++k) ref += a_val(r + deltas[d], k) * b_val(n, k);
          if (out[((size_t)d * 128 + r) * N + n] != ref) { if (!bad) first_r = r; ++bad; }
        }
      printf("  delta %3d: %s (%d mismatches%s)\n", deltas[d], bad ? "FAIL" : "ok", bad, bad ? "" : "");
      if (bad && first_r >= 0) printf("            first bad row %d\n", first_r);
    }
  }
  return 0;
}
