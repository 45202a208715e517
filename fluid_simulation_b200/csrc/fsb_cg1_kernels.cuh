// Single-reduction (Chronopoulos-Gear) form of the Jacobi-preconditioned CG: ONE sweep and ONE
// reduction point per iteration instead of two.  Same Krylov iteration as Eigen's (identical
// iteration counts in fp32 on the bench scenes, tools/studies/cg_single_reduction_study.py); the
// price is two more vectors (s = A p and w = A z kept by recurrence) and 45 instead of 32 B per
// cell and iteration.  Meant for the regimes where the two synchronisations per iteration, not the
// bytes, bound the solve: L2-resident grids and (later) short slabs of a sharded solve.
// Opt-in (FSB_CG_MODE=single), single GPU.  Status: verified on the CPU through the host
// emulation of this very source (tests/test_cpu_emul.py); NOT yet run on a GPU.
//
// One sweep, per cell (z = D^-1 r, all vectors exactly zero outside LIQUID cells):
//   p  <- z_old + beta p          s' <- w_old + beta s_old
//   x  <- x + alpha p             r' <- r_old - alpha s'
//   z' <- D^-1 r'                 w' <- A z'       (z' of the four neighbours is recomputed from their
//                                                   r_old, s_old, w_old: no second pass, no halo exchange)
//   partial sums: gamma' = r'.z', delta' = z'.w', |r'|^2
// then (last CTA): stop if |r'|^2 < threshold, else beta' = gamma'/gamma,
// alpha' = gamma' / (delta' - beta' gamma' / alpha).  r, s, w are double-buffered (a neighbour's old
// values are still needed while a cell is being updated); p and x are updated in place.
#pragma once

#include "fsb.h"
#include "fsb_device.cuh"

namespace {

struct Cg1Coef
{
  float off;        // (float)(1 / dx^2)            src/FluidSolver.cpp:382
  float diag[5];    // (float)(-n / dx^2)           :409-410
  float invdiag[5]; // Eigen DiagonalPreconditioner: diag != 0 ? 1 / diag : 1
};

__device__ __forceinline__ float cg1_invdiag(const Cg1Coef& k, uint32_t code)
{
  // code = 1 + n for a LIQUID cell with n non-SOLID neighbours; 0 -> masked (z = 0)
  return code == 5 ? k.invdiag[4] : code == 4 ? k.invdiag[3] : code == 3 ? k.invdiag[2]
       : code == 2 ? k.invdiag[1] : code == 1 ? k.invdiag[0] : 0.0f;
}
__device__ __forceinline__ float cg1_diag(const Cg1Coef& k, uint32_t code)
{
  return code == 5 ? k.diag[4] : code == 4 ? k.diag[3] : code == 3 ? k.diag[2]
       : code == 2 ? k.diag[1] : 0.0f;
}

// the new residual and its preconditioned form at one cell, from the OLD vectors
__device__ __forceinline__ void cg1_cell(float r_old, float s_old, float w_old, uint32_t code,
                                         float alpha, float beta, const Cg1Coef& k, float* s_new,
                                         float* r_new, float* z_new)
{
  const float sn = fmaf(beta, s_old, w_old);
  const float rn = fmaf(-alpha, sn, r_old);
  *s_new = sn;
  *r_new = rn;
  *z_new = cg1_invdiag(k, code) * rn;
}

__device__ __forceinline__ uint32_t cg1_code_at(const uint8_t* __restrict__ code, int ld, int nx, int ny,
                                                int i, int j)
{
  return (i >= 0 && i < nx && j >= 0 && j < ny) ? code[i + (size_t)j * ld] : 0u;
}

// One float4 group (cells i0..i0+3 of row j).  Reads the old r / s / w of the group, of the rows
// above and below and of the two cells left and right; writes p, x (in place) and the new r / s / w.
__device__ __forceinline__ void cg1_group(const float* __restrict__ r_old, const float* __restrict__ s_old,
                                          const float* __restrict__ w_old, float* __restrict__ r_new,
                                          float* __restrict__ s_new, float* __restrict__ w_new,
                                          float* __restrict__ p, float* __restrict__ x,
                                          const uint8_t* __restrict__ code, int nx, int ny, int ld,
                                          int i0, int j, float alpha, float beta, const Cg1Coef& k,
                                          double* acc_gamma, double* acc_delta, double* acc_r2)
{
  const size_t g = i0 + (size_t)j * ld;
  const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + g);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 == 0u)
  {
    // four masked cells: every vector stays exactly zero
    *reinterpret_cast<float4*>(r_new + g) = zero;
    *reinterpret_cast<float4*>(s_new + g) = zero;
    *reinterpret_cast<float4*>(w_new + g) = zero;
    return;
  }
  const float4 ro = *reinterpret_cast<const float4*>(r_old + g);
  const float4 so = *reinterpret_cast<const float4*>(s_old + g);
  const float4 wo = *reinterpret_cast<const float4*>(w_old + g);
  float4 pc = *reinterpret_cast<const float4*>(p + g);
  float4 xc = *reinterpret_cast<const float4*>(x + g);
  float sn[4], rn[4], zn[4];
#pragma unroll
  for (int t = 0; t < 4; ++t)
    cg1_cell(f4_get(ro, t), f4_get(so, t), f4_get(wo, t), (c4 >> (8 * t)) & 0xffu, alpha, beta, k,
             &sn[t], &rn[t], &zn[t]);
  // z' of the neighbours: south / north rows (four cells each), west / east cells
  float zs[4] = {0.f, 0.f, 0.f, 0.f}, zn_[4] = {0.f, 0.f, 0.f, 0.f}, zw = 0.0f, ze = 0.0f;
  if (j > 0)
  {
    const size_t q = g - ld;
    const uint32_t cc = *reinterpret_cast<const uint32_t*>(code + q);
    if (cc != 0u)
    {
      const float4 a = *reinterpret_cast<const float4*>(r_old + q);
      const float4 b = *reinterpret_cast<const float4*>(s_old + q);
      const float4 c = *reinterpret_cast<const float4*>(w_old + q);
      float d0, d1;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        cg1_cell(f4_get(a, t), f4_get(b, t), f4_get(c, t), (cc >> (8 * t)) & 0xffu, alpha, beta, k, &d0,
                 &d1, &zs[t]);
    }
  }
  if (j + 1 < ny)
  {
    const size_t q = g + ld;
    const uint32_t cc = *reinterpret_cast<const uint32_t*>(code + q);
    if (cc != 0u)
    {
      const float4 a = *reinterpret_cast<const float4*>(r_old + q);
      const float4 b = *reinterpret_cast<const float4*>(s_old + q);
      const float4 c = *reinterpret_cast<const float4*>(w_old + q);
      float d0, d1;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        cg1_cell(f4_get(a, t), f4_get(b, t), f4_get(c, t), (cc >> (8 * t)) & 0xffu, alpha, beta, k, &d0,
                 &d1, &zn_[t]);
    }
  }
  {
    const uint32_t cw = cg1_code_at(code, ld, nx, ny, i0 - 1, j);
    if (cw != 0u)
    {
      float d0, d1;
      cg1_cell(r_old[g - 1], s_old[g - 1], w_old[g - 1], cw, alpha, beta, k, &d0, &d1, &zw);
    }
    const uint32_t ce = cg1_code_at(code, ld, nx, ny, i0 + 4, j);
    if (ce != 0u)
    {
      float d0, d1;
      cg1_cell(r_old[g + 4], s_old[g + 4], w_old[g + 4], ce, alpha, beta, k, &d0, &d1, &ze);
    }
  }
  float4 r4, s4, w4;
#pragma unroll
  for (int t = 0; t < 4; ++t)
  {
    const uint32_t cd = (c4 >> (8 * t)) & 0xffu;
    float wv = 0.0f;
    if (cd != 0u)
    {
      const float west = (t == 0) ? zw : zn[t - 1];
      const float east = (t == 3) ? ze : zn[t + 1];
      const float sum = (west + east) + (zs[t] + zn_[t]);
      wv = fmaf(cg1_diag(k, cd), zn[t], k.off * sum);
      // p <- z_old + beta p, x <- x + alpha p
      const float z_old = cg1_invdiag(k, cd) * f4_get(ro, t);
      const float pn = fmaf(beta, f4_get(pc, t), z_old);
      f4_set(pc, t, pn);
      f4_set(xc, t, fmaf(alpha, pn, f4_get(xc, t)));
      *acc_gamma += (double)rn[t] * (double)zn[t];
      *acc_delta += (double)zn[t] * (double)wv;
      *acc_r2 += (double)rn[t] * (double)rn[t];
    }
    else
    {
      sn[t] = 0.0f;
      rn[t] = 0.0f;
    }
    f4_set(r4, t, rn[t]);
    f4_set(s4, t, sn[t]);
    f4_set(w4, t, wv);
  }
  *reinterpret_cast<float4*>(r_new + g) = r4;
  *reinterpret_cast<float4*>(s_new + g) = s4;
  *reinterpret_cast<float4*>(w_new + g) = w4;
  *reinterpret_cast<float4*>(p + g) = pc;
  *reinterpret_cast<float4*>(x + g) = xc;
}

// scalar recurrences at the single reduction point (thread 0 of the last CTA / the host emulation)
struct Cg1Scalars
{
  double gamma, delta, r2, rhs2;
  float alpha, beta, thr;
  int iter, done, max_iters, init;
  unsigned int ticket;
};

__host__ __device__ inline void cg1_advance(Cg1Scalars* s, double gamma_new, double delta_new, double r2)
{
  s->r2 = r2;
  if (s->init)
  {
    // the set-up sweep (alpha = beta = 0): gamma_0 = r.z, delta_0 = z.Az
    s->init = 0;
    s->gamma = gamma_new;
    s->delta = delta_new;
    s->beta = 0.0f;
    s->alpha = (float)(gamma_new / delta_new);
    return;
  }
  if ((float)r2 < s->thr)
  {
    s->done = 1; // converged: Eigen breaks before i++
    return;
  }
  const float beta = (float)(gamma_new / s->gamma);
  const float alpha = (float)(gamma_new / (delta_new - (double)beta * gamma_new / (double)s->alpha));
  s->gamma = gamma_new;
  s->delta = delta_new;
  s->beta = beta;
  s->alpha = alpha;
  s->iter += 1;
  if (s->iter >= s->max_iters) s->done = 1;
}

} // namespace
