// tmap.cuh -- host-side TMA descriptor (CUtensorMap) construction without linking libcuda.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace avexk {

// bf16 row-major [rows, cols] with row pitch `ld` elements; box = [box_rows, box_cols]; 128-byte swizzle
// (box_cols * 2 must be 128); out-of-bounds elements are zero-filled.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows,
                      int box_cols);
// fp32 row-major [rows, cols] with row pitch `ld` elements; box = [box_rows, 32] (128-byte rows, 128-byte swizzle).
int make_tmap_2d_f32(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
// row-major [rows, cols] of 2-byte (bf16) or 4-byte (fp32) elements; box = [box_rows, 64 bytes] with 64-byte swizzle.
int make_tmap_2d_64B(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int elem_bytes, int box_rows);
// bf16 [d2, d1, d0] (d0 innermost) with pitches ld1 (elements between d1 steps) and ld2; box [b2, b1, b0].
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, long long d0, long long d1, long long d2, long long ld1,
                      long long ld2, int b0, int b1, int b2, bool swizzle128);

// 16-bit NHWC activations [B, H, W, C] as a 4-D map (C innermost); box = [1, 1, box_w, box_c], no swizzle.  Coordinates outside
// the image (negative or >= W / H) read as zero: the zero padding of a convolution costs nothing.
int make_tmap_nhwc16(CUtensorMap* map, const void* base, int B, int H, int W, int C, int box_c, int box_w);

}  // namespace avexk
