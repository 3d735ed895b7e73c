// bf16 tcgen05/TMEM tensor-core GEMM for sm_100a (placeholder until the kernel lands).
#include "common.cuh"
#include "../../include/hulc2_b200.h"

int hulc2_gemm_bf16_impl(const hulc2_gemm_args* a, cudaStream_t st) {
  (void)a; (void)st;
  hulc2_set_error("gemm: bf16 tcgen05 path not built yet");
  return HULC2_ENOTIMPL;
}
