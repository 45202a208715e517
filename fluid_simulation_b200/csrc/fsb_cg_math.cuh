// Per-float4 arithmetic of the CG sweeps (fsb_cg.cu, fsb_cg_one.cu): the 5-point operator, the
// direction update and the dot-product partials.  Plain C++ on float4, so tests/cpu_emul compiles
// the same source for the host.  Coefficients: src/FluidSolver.cpp:378-410.
#pragma once

#include <stdint.h>

namespace {

// stencil-code word of four LIQUID cells with four non-SOLID neighbours each (the bulk of any scene)
constexpr uint32_t kInterior4 = 0x05050505u;

// new direction for four cells: z + beta p_old with z = invdiag r
__device__ __forceinline__ float4 direction4(const float4 r4, const float4 p4, uint32_t c4,
                                             const float4* lut, float inv5, float beta)
{
  float i0 = inv5, i1 = inv5, i2 = inv5, i3 = inv5;
  if (c4 != kInterior4)
  {
    i0 = lut[c4 & 0xff].x;
    i1 = lut[(c4 >> 8) & 0xff].x;
    i2 = lut[(c4 >> 16) & 0xff].x;
    i3 = lut[c4 >> 24].x;
  }
  return make_float4(fmaf(beta, p4.x, i0 * r4.x), fmaf(beta, p4.y, i1 * r4.y),
                     fmaf(beta, p4.z, i2 * r4.z), fmaf(beta, p4.w, i3 * r4.w));
}

// q = A p for four cells: centre pc, west / east scalars, south / north float4
__device__ __forceinline__ float4 apply_a4(const float4 pc, float w, float e, const float4 s4,
                                           const float4 n4, uint32_t c4, const float4* lut,
                                           float diag5, float off)
{
  const float a0 = (w + pc.y) + (s4.x + n4.x);
  const float a1 = (pc.x + pc.z) + (s4.y + n4.y);
  const float a2 = (pc.y + pc.w) + (s4.z + n4.z);
  const float a3 = (pc.z + e) + (s4.w + n4.w);
  if (c4 == kInterior4)
    return make_float4(fmaf(diag5, pc.x, off * a0), fmaf(diag5, pc.y, off * a1),
                       fmaf(diag5, pc.z, off * a2), fmaf(diag5, pc.w, off * a3));
  const float4 k0 = lut[c4 & 0xff], k1 = lut[(c4 >> 8) & 0xff], k2 = lut[(c4 >> 16) & 0xff],
               k3 = lut[c4 >> 24];
  return make_float4(fmaf(k0.y, pc.x, k0.z * a0), fmaf(k1.y, pc.y, k1.z * a1),
                     fmaf(k2.y, pc.z, k2.z * a2), fmaf(k3.y, pc.w, k3.z * a3));
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b)
{
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

__device__ __forceinline__ float4 fma4(float a, const float4 b, const float4 c)
{
  return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}

} // namespace
