// Datamodule-side kernels: camera frames stay uint8 HWC -- the layout CALVIN episodes are stored in --
// all the way into HBM; ONE kernel per camera does what the reference does on CPU dataloader workers per window:
//   window gather + pad-with-repetition      hulc2/datasets/base_dataset.py:121-163 (pad_sequence / pad_with_repetition)
//   HWC -> CHW                               hulc2/datasets/utils/episode_utils.py:78-82 (process_rgb)
//   RandomShiftsAug(pad)                     hulc2/utils/transforms.py:85-106
//   ScaleImageTensor, Normalize(0.5, 0.5)    hulc2/utils/transforms.py:8-19, conf/datamodule/transforms/rand_shift.yaml:5-10
// and writes either the conv trunk's packed bf16 operand directly (no fp32 frame ever exists) or fp32 NCHW frames.
//
// RandomShiftsAug is an integer crop of the replicate-padded frame: base grid coordinate i + integer shift s maps to padded
// pixel i + s exactly (linspace step 2/(h+2p), align_corners=False), so out[y,x] = in[clamp(y+dy), clamp(x+dx)] with
// (dx,dy) = (sx-pad, sy-pad) in [-pad, pad]; the shift draw is an INPUT (int32 [F,2] = (dx,dy) per frame), like all noise.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/hulc2_b200.h"

namespace {

__device__ __forceinline__ float norm_u8(unsigned v) {
  // float().div(255) then Normalize(mean .5, std .5): (x - 0.5) / 0.5, same operation order as the reference (IEEE, no fast-math)
  const float x = __fdiv_rn((float)v, 255.0f);
  return __fdiv_rn(__fsub_rn(x, 0.5f), 0.5f);
}

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// source frame of output frame f = (window b, step t): start[b] + min(t, len[b]-1) -- pad_with_repetition of the last frame
__device__ __forceinline__ long long src_frame(int f, int S, const long long* __restrict__ start, const int* __restrict__ len) {
  if (!start) return f;
  const int b = f / S, t = f - b * S;
  const int L = len ? len[b] : S;
  return start[b] + (long long)min(t, max(L, 1) - 1);
}

// Loads the source rows y0 .. y0+nrows-1 (shifted + clamped in y) of a frame into shared memory as raw bytes; the x
// shift/clamp is applied when reading.  rows[r][x*C + c].  All loads of a block are issued before the first use, so a
// block keeps nrows * W * C bytes in flight.
__device__ __forceinline__ void load_rows(const uint8_t* __restrict__ frame, uint8_t* rows, int nrows, int y0, int dy, int H, int W, int C) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;      // 1-D and 2-D blocks
  const int rb = W * C;                                      // bytes per row
  const bool vec8 = (rb & 7) == 0 && ((reinterpret_cast<uintptr_t>(frame) & 7) == 0);
  const bool vec = (rb & 3) == 0 && ((reinterpret_cast<uintptr_t>(frame) & 3) == 0);
  if (vec8) {                                              // 8-byte loads: half the load/store instructions of the 4-byte path
    const int nw = rb >> 3;
    for (int q = tid; q < nrows * nw; q += nth) {
      const int a = q / nw, i = q - a * nw;
      const int y = min(max(y0 + a + dy, 0), H - 1);
      reinterpret_cast<uint2*>(rows)[a * nw + i] = __ldg(reinterpret_cast<const uint2*>(frame + (size_t)y * rb) + i);
    }
  } else if (vec) {
    const int nw = rb >> 2;
    for (int q = tid; q < nrows * nw; q += nth) {
      const int a = q / nw, i = q - a * nw;
      const int y = min(max(y0 + a + dy, 0), H - 1);
      reinterpret_cast<uint32_t*>(rows)[a * nw + i] = __ldg(reinterpret_cast<const uint32_t*>(frame + (size_t)y * rb) + i);
    }
  } else {
    for (int q = tid; q < nrows * rb; q += nth) {
      const int a = q / rb, i = q - a * rb;
      const int y = min(max(y0 + a + dy, 0), H - 1);
      rows[a * rb + i] = __ldg(frame + (size_t)y * rb + i);
    }
  }
}

// uint8 [*,H,W,C] -> bf16 [F, H/4, W/4, 16 C], channel (ci, a, b) = norm(x[src(f), clamp(4I+a+dy), clamp(4J+b+dx), ci]).
// One block packs R consecutive packed rows (4R image rows) of one frame: blockIdx.x = f * nseg + segment.
template <int CT>    // CT = 3: RGB (compile-time divisors for the per-element index arithmetic), 0: any channel count
__global__ void __launch_bounds__(320) frames_u8_pack_kernel(const uint8_t* __restrict__ store, const long long* __restrict__ start,
                                                             const int* __restrict__ len, const int* __restrict__ shift,
                                                             uint8_t* __restrict__ xs, int S, int Crt, int H, int W, int H4, int W4, int R,
                                                             int nseg) {
  const int C = CT ? CT : Crt;
  extern __shared__ __align__(16) uint8_t rows[];           // [4R][W*C]
  const int f = blockIdx.x / nseg, I0 = (blockIdx.x - f * nseg) * R;
  const int nI = min(R, H4 - I0);
  const uint8_t* frame = store + (size_t)src_frame(f, S, start, len) * H * W * C;
  const int dx = shift ? shift[2 * f] : 0, dy = shift ? shift[2 * f + 1] : 0;
  load_rows(frame, rows, 4 * nI, 4 * I0, dy, H, W, C);
  __syncthreads();
  const int rb = W * C, c16 = 16 * C, cpc = c16 / 8, per_row = W4 * cpc;
  uint4* out = reinterpret_cast<uint4*>(xs + ((size_t)f * H4 + I0) * W4 * c16 * 2);
  // Thread (x, y): 16-byte chunk r = (packed pixel J, 8 of its 16 C channels) of a packed row, rows Il = y, y + blockDim.y, ...
  // Everything that depends only on r -- channel, the two source rows inside the 4-row group, the four clamped x offsets -- is
  // computed once and reused for every packed row (ncu r02: 68 % of the issue slots busy on a division by per_row, the clamps and
  // the address arithmetic of EVERY chunk: ~100 instructions per 16 bytes written; now ~45).
  for (int r = threadIdx.x; r < per_row; r += blockDim.x) {
    const int J = r / cpc, e0 = (r - J * cpc) * 8;          // cpc is a compile-time constant for RGB frames
    const int ci = e0 >> 4, a0 = (e0 >> 2) & 3;
    int xo[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) xo[b] = min(max(4 * J + b + dx, 0), W - 1) * C + ci;
    for (int Il = threadIdx.y; Il < nI; Il += blockDim.y) {
      const uint8_t* r0 = rows + (4 * Il + a0) * rb;
      const uint8_t* r1 = r0 + rb;
      // bf16 output: (2 v - 255) * fp32(1/255) rounds to the SAME bf16 as the reference chain ((v / 255) - 0.5) / 0.5 for all 256
      // grey levels (checked exhaustively on the host and by test_all_256_grey_levels_exact), without a table lookup per element
      float v[8];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        v[b] = __fmul_rn(__int2float_rn(2 * (int)r0[xo[b]] - 255), 1.0f / 255.0f);
        v[4 + b] = __fmul_rn(__int2float_rn(2 * (int)r1[xo[b]] - 255), 1.0f / 255.0f);
      }
      out[Il * per_row + r] = make_uint4(bf16x2(v[0], v[1]), bf16x2(v[2], v[3]), bf16x2(v[4], v[5]), bf16x2(v[6], v[7]));
    }
  }
  // The 128 bytes of slack behind the last pixel: conv1's halo path reads every 48-channel pixel as a 64-element row, so the
  // last pixel's row ends in the slack.  Those positions meet zero weights, but 0 x NaN = NaN: they must hold finite values.
  if (blockIdx.x == gridDim.x - 1 && threadIdx.y == 0 && threadIdx.x < 8)
    reinterpret_cast<uint4*>(xs + (size_t)gridDim.x / nseg * H4 * W4 * c16 * 2)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
}

// uint8 [*,H,W,C] -> fp32 NCHW [F,C,H,W] (the reference batch contract; fp32 parity path).  One block per (frame, row).
__global__ void frames_u8_to_f32_kernel(const uint8_t* __restrict__ store, const long long* __restrict__ start, const int* __restrict__ len,
                                        const int* __restrict__ shift, float* __restrict__ out, int S, int C, int H, int W) {
  extern __shared__ __align__(16) uint8_t rows[];           // [1][W*C]
  __shared__ float lut[256];
  lut[threadIdx.x & 255] = norm_u8(threadIdx.x & 255);
  const int f = blockIdx.x / H, y = blockIdx.x - f * H;
  const uint8_t* frame = store + (size_t)src_frame(f, S, start, len) * H * W * C;
  const int dx = shift ? shift[2 * f] : 0, dy = shift ? shift[2 * f + 1] : 0;
  load_rows(frame, rows, 1, y, dy, H, W, C);
  __syncthreads();
  for (int q = threadIdx.x; q < C * W; q += blockDim.x) {
    const int c = q / W, x = q - c * W;
    const int xs_ = min(max(x + dx, 0), W - 1);
    out[(((size_t)f * C + c) * H + y) * W + x] = lut[rows[xs_ * C + c]];
  }
}

// Per-step vectors of a window batch: out[b,t,:] = store[start[b] + t] for t < len[b]; padded steps by `mode`
// (base_dataset.py:129-150): 0 = repeat the last valid row, 1 = zeros, 2 = relative actions: zeros except the last
// component (gripper), which repeats.
__global__ void window_gather_kernel(const float* __restrict__ store, const long long* __restrict__ start, const int* __restrict__ len,
                                     float* __restrict__ out, int B, int S, int D, int mode) {
  const long long total = (long long)B * S * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long r = i / D;
    const int t = (int)(r % S), b = (int)(r / S);
    const int L = max(len ? len[b] : S, 1);
    const bool pad = t >= L;
    const float last = store[(start[b] + min(t, L - 1)) * D + d];
    float v = last;
    if (pad && (mode == 1 || (mode == 2 && d != D - 1))) v = 0.f;
    out[i] = v;
  }
}

}  // namespace

extern "C" {

int hulc2_frames_u8_pack_bf16(const void* store, const long long* win_start, const int* win_len, const int* shift, void* xs, int F, int S,
                              int C, int H, int W, cudaStream_t st) {
  if (F == 0) return HULC2_OK;
  const int H4 = H / 4, W4 = W / 4;
  if (!store || !xs || F < 0 || S <= 0 || H4 <= 0 || W4 <= 0 || C <= 0 || (16 * C) % 8 != 0) { hulc2_set_error("frames_u8_pack: bad geometry"); return HULC2_EINVAL; }
  if ((size_t)4 * W * C > 24 * 1024) { hulc2_set_error("frames_u8_pack: frame row too wide"); return HULC2_EINVAL; }
  // packed rows per block: as many as keep the staging buffer <= 12 KB (8 resident blocks per SM), at most 8
  int R = (int)((12 * 1024) / ((size_t)4 * W * C));
  R = R < 1 ? 1 : (R > 8 ? 8 : R);
  if (R > H4) R = H4;
  const int nseg = (H4 + R - 1) / R;
  const size_t smem = (size_t)4 * R * W * C;
  // block = (chunks of a packed row rounded up to a warp, at most 320) x (as many packed rows as fit 256..320 threads)
  const int per_row = W4 * (16 * C / 8);
  int tx = ((per_row + 31) / 32) * 32;
  if (tx > 320) tx = 320;
  int ty = 320 / tx;
  if (ty > R) ty = R;
  if (ty < 1) ty = 1;
  const dim3 block(tx, ty);
  if (C == 3) frames_u8_pack_kernel<3><<<F * nseg, block, smem, st>>>((const uint8_t*)store, win_start, win_len, shift, (uint8_t*)xs, S, C, H, W, H4, W4, R, nseg);
  else frames_u8_pack_kernel<0><<<F * nseg, block, smem, st>>>((const uint8_t*)store, win_start, win_len, shift, (uint8_t*)xs, S, C, H, W, H4, W4, R, nseg);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_frames_u8_to_f32(const void* store, const long long* win_start, const int* win_len, const int* shift, float* out, int F, int S,
                           int C, int H, int W, cudaStream_t st) {
  if (F == 0) return HULC2_OK;
  if (!store || !out || F < 0 || S <= 0 || H <= 0 || W <= 0 || C <= 0) { hulc2_set_error("frames_u8_to_f32: bad geometry"); return HULC2_EINVAL; }
  const size_t smem = (size_t)W * C;
  if (smem > 48 * 1024) { hulc2_set_error("frames_u8_to_f32: frame row too wide"); return HULC2_EINVAL; }
  frames_u8_to_f32_kernel<<<F * H, 256, smem, st>>>((const uint8_t*)store, win_start, win_len, shift, out, S, C, H, W);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

int hulc2_window_gather_f32(const float* store, const long long* win_start, const int* win_len, float* out, int B, int S, int D, int mode,
                            cudaStream_t st) {
  if (B == 0) return HULC2_OK;
  if (!store || !win_start || !out || B < 0 || S <= 0 || D <= 0 || mode < 0 || mode > 2) { hulc2_set_error("window_gather: bad arguments"); return HULC2_EINVAL; }
  const long long total = (long long)B * S * D;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  window_gather_kernel<<<blocks, 256, 0, st>>>(store, win_start, win_len, out, B, S, D, mode);
  HULC2_CHECK_LAUNCH();
  return HULC2_OK;
}

}  // extern "C"
