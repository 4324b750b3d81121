// Activation "slab": the HBM layout of the 64-channel hidden activations inside a conv coupling layer.
//
// One row = one spatial position = 64 bf16 channels = 128 B, stored with the 128-byte UMMA/TMA swizzle already applied
// (16-byte chunk j of buffer row R lives at chunk j ^ (R & 7)), so a contiguous run of rows can be brought into shared
// memory with ONE 1-D bulk async copy and used directly as a K-major SW128 tcgen05 operand.
//
// Zero padding is part of the layout: images are laid out with a row pitch of P = slab_pitch(W) >= W + 1 (at least one shared
// zero column between the right edge of a line and the left edge of the next; normally a multiple of 8 so that every image line
// starts at the same swizzle phase and a line can be copied to any 8-row-aligned place in shared memory -- the line-tile
// convolution kernel gathers one line of several images into one 128-row MMA tile) and H + 1 lines per image (the zero
// line above image n is the zero line below image n - 1).
// Pixel (n, y, x) sits at data row  n * IMG + (y + 1) * P + (x + 1),  IMG = (H + 1) * P,
// and its 3x3 neighbour (dy, dx) at  +dy * P + dx  -- always inside the slab, always zero where the reference's SAME
// padding reads zero.  The 3x3 implicit GEMM therefore needs no im2col: the nine taps are nine row-shifted views of the
// same shared-memory slab.  Buffer row = HALO + data row; HALO = round_up(P + 1, 8) guard rows at both ends stay zero.
#pragma once
#include <stdint.h>

namespace bruno {

constexpr int SLAB_CH = 64;
constexpr int SLAB_ROW_BYTES = 128;
constexpr int TILE_M = 128;

struct SlabGeom {
  int N, H, W;
  int P;          // row pitch (positions per image line)
  int IMG;        // positions per image
  int HALO;       // guard rows before the first data row (multiple of 8)
  int data_rows;  // N * IMG
  int NT;         // number of 128-row tiles covering the data rows
  int rows;       // total buffer rows
};

// Row pitch: W + 1 rounded up to a multiple of 8 (what the line-tile kernels need) unless that wastes more than a quarter
// of the rows (W = 16 -> 24, W = 8 -> 16: such widths keep the exact pitch W + 1 and run the per-tap kernels).
__host__ __device__ inline int slab_pitch(int W) {
  const int p8 = ((W + 1 + 7) / 8) * 8;
  return 4 * p8 <= 5 * (W + 1) ? p8 : W + 1;
}

__host__ __device__ inline SlabGeom slab_geom(int N, int H, int W) {
  SlabGeom g;
  g.N = N; g.H = H; g.W = W;
  g.P = slab_pitch(W);
  g.IMG = (H + 1) * g.P;
  g.HALO = ((g.P + 1 + 7) / 8) * 8;
  g.data_rows = N * g.IMG;
  g.NT = (g.data_rows + TILE_M - 1) / TILE_M;
  g.rows = (g.NT + 2) * TILE_M + 2 * g.HALO;   // two spare zero tiles: kernels with 126-row tiles overrun slightly
  return g;
}

// byte offset of 16-byte chunk j (channels 8j..8j+7) of buffer row R
__host__ __device__ inline size_t slab_chunk_off(size_t R, int j) {
  return R * SLAB_ROW_BYTES + static_cast<size_t>((j ^ static_cast<int>(R & 7)) << 4);
}

// data row -> is it a pixel (vs. a zero-padding position)?
__host__ __device__ inline bool slab_row_is_pixel(int data_row, const SlabGeom& g) {
  if (data_row >= g.data_rows) return false;
  const int q = data_row % g.IMG;
  return (q / g.P) >= 1 && (q % g.P) >= 1 && (q % g.P) <= g.W;
}

}  // namespace bruno
