// Exhaustive host-side check of csrc/mask_bits.h (the separable voxel-mask construction used by the k-NN kernels)
// against the straightforward per-bit formulation: bit s = (z*4 + y)*4 + x of the 64-bit block mask is set iff
// xm[x] & ym[y] & zm[z]; stencil_axis_bits against the interval intersection it replaces; and the cached inner-stencil
// patterns (x&y per block quarter, z per block half and word) exactly as k_knn_stencil stages them.
// Built and run by tests/test_mask_bits.py:  g++ -O1 -I better_fastlio2_b200/csrc tests/cpp/mask_equivalence.cpp
#include <cstdio>

#include "mask_bits.h"

static unsigned long long ref_mask(unsigned xm, unsigned ym, unsigned zm) {
  unsigned long long m = 0ull;
  for (int z = 0; z < 4; ++z)
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (((xm >> x) & 1u) && ((ym >> y) & 1u) && ((zm >> z) & 1u)) m |= 1ull << ((z * 4 + y) * 4 + x);
  return m;
}
static unsigned ref_axis(int b, int cv) {   // voxels cv-2..cv+2 intersected with block b = voxels 4b..4b+3
  unsigned m = 0u;
  for (int l = 0; l < 4; ++l) {
    const int v = 4 * b + l;
    if (v >= cv - 2 && v <= cv + 2) m |= 1u << l;
  }
  return m;
}

int main() {
  long bad = 0;
  for (unsigned x = 0; x < 16; ++x)
    for (unsigned y = 0; y < 16; ++y)
      for (unsigned z = 0; z < 16; ++z)
        if (flb::mask_from_axes(x, y, z) != ref_mask(x, y, z)) ++bad;
  for (int cv = -70; cv <= 70; ++cv)
    for (int b = -22; b <= 22; ++b)
      if (flb::stencil_axis_bits(b, cv) != ref_axis(b, cv)) ++bad;
  // inner 3-wide stencil staged as in k_knn_stencil: ix = 14 << ox over the two blocks of an axis
  for (int ox = 0; ox < 4; ++ox)
    for (int oy = 0; oy < 4; ++oy)
      for (int oz = 0; oz < 4; ++oz) {
        const unsigned ix = 14u << ox, iy = 14u << oy, iz = 14u << oz;
        unsigned xy3[4], z3[4];
        for (int q = 0; q < 4; ++q) {
          xy3[q] = flb::xpat32((ix >> ((q & 1) << 2)) & 15u) & flb::ypat32((iy >> ((q >> 1) << 2)) & 15u);
          z3[q] = flb::zpat32((((iz >> ((q >> 1) << 2)) & 15u) >> ((q & 1) << 1)) & 3u);
        }
        for (int b = 0; b < 8; ++b) {
          const unsigned long long ref = ref_mask((ix >> ((b & 1) << 2)) & 15u, (iy >> (((b >> 1) & 1) << 2)) & 15u, (iz >> ((b >> 2) << 2)) & 15u);
          const unsigned xy = xy3[b & 3];
          const int zq = (b >> 2) << 1;
          const unsigned long long got = ((unsigned long long)(xy & z3[zq + 1]) << 32) | (unsigned long long)(xy & z3[zq]);
          if (ref != got) ++bad;
        }
      }
  std::printf("MASK_BITS %s (%ld mismatches)\n", bad ? "FAIL" : "OK", bad);
  return bad != 0;
}
