// mask_bits.h — bit-mask helpers of the k-NN kernels, usable from host code too (tests/cpp/mask_equivalence.cpp checks
// them exhaustively against the straightforward per-bit formulation).
#pragma once
#if defined(__CUDACC__)
#define FLB_HD __host__ __device__ __forceinline__
#else
#define FLB_HD inline
#endif

namespace flb {

// 64-bit voxel mask of a 4x4x4 block (slot order s = (z*4 + y)*4 + x) from three 4-bit per-axis masks.  The mask is
// separable: (x pattern) & (y pattern) & (z pattern), each built with a multiply that replicates a small bit group
// (no carries: the replicated groups never overlap), on 32-bit halves (z = 0,1 | z = 2,3).
FLB_HD unsigned xpat32(unsigned xm) { return xm * 0x11111111u; }   // xm in every nibble
FLB_HD unsigned ypat32(unsigned ym) {                              // nibble y of every 16-bit group
  const unsigned sp = (ym & 1u) | ((ym & 2u) << 3) | ((ym & 4u) << 6) | ((ym & 8u) << 9);   // bit y -> bit 4y
  return (sp * 0xFu) * 0x00010001u;
}
FLB_HD unsigned zpat32(unsigned z2) {   // two z bits of one half: bit 0 -> low 16 bits, bit 1 -> high 16
  return ((0u - (z2 & 1u)) & 0x0000FFFFu) | ((0u - ((z2 >> 1) & 1u)) & 0xFFFF0000u);
}
FLB_HD unsigned long long mask_from_xy(unsigned xy, unsigned zm) {
  return ((unsigned long long)(xy & zpat32(zm >> 2)) << 32) | (unsigned long long)(xy & zpat32(zm & 3u));
}
FLB_HD unsigned long long mask_from_axes(unsigned xm, unsigned ym, unsigned zm) {
  return mask_from_xy(xpat32(xm) & ypat32(ym), zm);
}
// 4-bit mask of the voxels of block b (per axis) that lie inside the 5-wide stencil around voxel cv: the stencil spans
// exactly the two blocks (cv-2)>>2 and (cv-2)>>2 + 1; inside them it is the run 31 << ((cv-2)&3) cut into nibbles.
FLB_HD unsigned stencil_axis_bits(int b, int cv) {
  const int h = b - ((cv - 2) >> 2);
  if ((unsigned)h >= 2u) return 0u;
  return ((31u << ((cv - 2) & 3)) >> (h << 2)) & 15u;
}

}  // namespace flb
