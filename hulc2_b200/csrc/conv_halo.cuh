// Parameters of the halo-tile convolution kernels (conv_halo_sm100.cu); geometry is filled by the C entry points in
// conv_sm100.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HaloParams {
  const uint8_t* w;        // packed bf16 weights [NT][ntaps * 64], tap-major contraction index
  const float* bias;       // [NT] or null (forward only)
  const uint8_t* mask;     // bf16, output-shaped: keep where > 0 (input gradient) or null
  const uint8_t* mask_bits;  // the same mask as 1 bit per output element (input gradient; takes precedence over `mask`) or null
  uint8_t* mask_out;       // forward: (y > 0) as 1 bit per output element, or null
  uint8_t* y;              // bf16 output [F, oH, oW, BNc]
  int NT;                  // accumulator columns = ncls * BNc (64 or 128)
  int BNc;                 // channels per output pixel
  int ncls;                // 1, or 4 = the stride-2 parity classes stacked along N
  int relu;
  int F, tiles_per_frame, ntiles;
  int BH, PW, PH;          // a tile = BH rows of the (class-)output; halo tile = PH x PW source pixels (raster pitch PW)
  int i_min, j_min;        // source coordinates of the halo tile's first pixel relative to (tile row 0, column 0)
  int ntaps;
  int oH, oW, oS;          // output tensor dims; pixel stride of a class (1, or 2 for stacked classes)
  int stages, stage_bytes;
  int nparts, part_bytes;  // sub-tiles per stage (2 = even / odd source rows of a stride-2 forward conv) and their byte pitch
  short delta[16];         // per tap: raster offset of its window inside its sub-tile
  short part[16];          // per tap: sub-tile index
  short clsH[4], clsW[4], clsPh[4], clsPw[4];  // valid rows / columns of each class and its pixel parity offsets
};

bool hulc2_conv_halo_enabled();
int hulc2_conv_halo_pitch(int pw);
int hulc2_conv_halo_launch(const void* src, int F, int Hs, int Ws, HaloParams p, bool dgrad, cudaStream_t st);
// stride-2 forward conv over a 32-channel source (two pixels = one 128-byte row; even / odd source rows = two sub-tiles)
// forward conv over packed pixels narrower than 64 channels (overlapping 64-element rows, zero weights for the overlap)
int hulc2_conv_halo_launch_packed(const void* src, int F, int Hs, int Ws, int pixel_elems, HaloParams p, cudaStream_t st);
int hulc2_conv_halo_launch_s2(const void* src, int F, int Hs, int Ws, HaloParams p, cudaStream_t st);

// Halo-tile weight gradient (stride-1 convs): dW^T[(tap, ch), co] = sum over pixels of X[pixel + delta(tap), ch] * dZ[pixel, co].
struct HaloWgradParams {
  float* partial;          // [grid][nblk * 64][64] fp32 per-CTA partial sums (rows: tap * 64 + ch; row ntaps * 64 = sum of dZ)
  int F, tiles_per_frame, ntiles;
  int BH, PW;              // a tile = BH rows of PW raster positions (PW >= source width; the extra columns are TMA zero fill)
  int KH;                  // source rows per tile = BH + KH - 1
  int ntaps, nmt;          // taps; M-tiles of 128 = 2 blocks of 64: ceil((ntaps + 1) / 2), the +1 is the block of ones
  int a_bytes, b_bytes;    // sub-tile sizes (1024-byte multiples); stage = a_bytes + b_bytes
  int stages;
  short delta[16];         // per tap: raster offset of its window in the source tile
};
int hulc2_conv_halo_wgrad(const void* x, int x_pixel_elems, const void* dz, int dz_pixel_elems, int F, int H, int W, int KH, int KW,
                          float* partial, long long partial_bytes, int* grid_out, int* nblk_out, cudaStream_t st);
