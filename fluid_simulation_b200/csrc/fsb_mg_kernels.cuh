// Level kernels of the opt-in multigrid preconditioner (fsb_mg.cu): smoothing, residual,
// restriction, prolongation, label coarsening, stencil codes.  All but the one-CTA coarse solve
// use no shared memory, shuffles or barriers, so tests/cpu_emul compiles this file for the host and
// checks a whole V-cycle against the numpy prototype (tools/studies/mgpcg_prototype.py) without a GPU.
// Product code: includes nothing of oracle/.
#pragma once

#include "fsb.h"
#include "fsb_device.cuh"

namespace {

constexpr int kMgCoarsest = 32; // the coarse-solve CTA handles up to 32 x 32 cells
constexpr int kMgStopDefault = 4; // coarsening continues until both sides are <= this (FSB_MG_STOP)
constexpr float kMgOmega = 2.0f / 3.0f;

struct MgCoef
{
  float inv_h2;
  float wdinv[5]; // omega / diagonal for 0..4 non-SOLID neighbours (0 for an isolated cell)
};


// A x on four cells of one row: inv_h2 * (W + E + S + N - cnt * x), zero where code == 0.
// x is exactly zero outside LIQUID cells and in the pad columns, so no neighbour masks are needed.
__device__ __forceinline__ float4 mg_apply4(const float* __restrict__ x, int ld, int ny, int i0,
                                            int j, const float4 xc, uint32_t code4, float inv_h2)
{
  const size_t k = i0 + (size_t)j * ld;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 xs = (j > 0) ? *reinterpret_cast<const float4*>(x + k - ld) : zero;
  const float4 xn = (j + 1 < ny) ? *reinterpret_cast<const float4*>(x + k + ld) : zero;
  const float w = (i0 > 0) ? x[k - 1] : 0.0f;
  const float e = (i0 + 4 < ld) ? x[k + 4] : 0.0f;
  float4 a;
  a.x = (w + xc.y) + (xs.x + xn.x);
  a.y = (xc.x + xc.z) + (xs.y + xn.y);
  a.z = (xc.y + xc.w) + (xs.z + xn.z);
  a.w = (xc.z + e) + (xs.w + xn.w);
  float4 q;
  const uint32_t c0 = code4 & 0xff, c1 = (code4 >> 8) & 0xff, c2 = (code4 >> 16) & 0xff, c3 = code4 >> 24;
  q.x = c0 ? inv_h2 * (a.x - (float)(c0 - 1) * xc.x) : 0.0f;
  q.y = c1 ? inv_h2 * (a.y - (float)(c1 - 1) * xc.y) : 0.0f;
  q.z = c2 ? inv_h2 * (a.z - (float)(c2 - 1) * xc.z) : 0.0f;
  q.w = c3 ? inv_h2 * (a.w - (float)(c3 - 1) * xc.w) : 0.0f;
  return q;
}

__device__ __forceinline__ float mg_wdinv(const MgCoef& k, uint32_t code)
{
  // code = 1 + n for a liquid cell: select instead of an indexed constant load
  return code == 5 ? k.wdinv[4] : code == 4 ? k.wdinv[3] : code == 3 ? k.wdinv[2]
       : code == 2 ? k.wdinv[1] : 0.0f;
}

// one damped-Jacobi sweep; FIRST: the input iterate is zero (x_out = omega D^-1 b)
template <bool FIRST>
__global__ void __launch_bounds__(256)
k_mg_smooth(const float* __restrict__ xin, const float* __restrict__ b,
            const uint8_t* __restrict__ code, float* __restrict__ xout, int nx, int ny, int ld,
            const MgCoef kf)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i0 >= ld) return;
  const size_t k = i0 + (size_t)j * ld;
  const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + k);
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 != 0u)
  {
    const float4 b4 = *reinterpret_cast<const float4*>(b + k);
    float4 res = b4, xc = out;
    if (!FIRST)
    {
      xc = *reinterpret_cast<const float4*>(xin + k);
      const float4 q = mg_apply4(xin, ld, ny, i0, j, xc, c4, kf.inv_h2);
      res = make_float4(b4.x - q.x, b4.y - q.y, b4.z - q.z, b4.w - q.w);
    }
    out.x = (c4 & 0xff) ? xc.x + mg_wdinv(kf, c4 & 0xff) * res.x : 0.0f;
    out.y = ((c4 >> 8) & 0xff) ? xc.y + mg_wdinv(kf, (c4 >> 8) & 0xff) * res.y : 0.0f;
    out.z = ((c4 >> 16) & 0xff) ? xc.z + mg_wdinv(kf, (c4 >> 16) & 0xff) * res.z : 0.0f;
    out.w = (c4 >> 24) ? xc.w + mg_wdinv(kf, c4 >> 24) * res.w : 0.0f;
  }
  *reinterpret_cast<float4*>(xout + k) = out;
}

// r = b - A x on LIQUID cells, 0 elsewhere
__global__ void __launch_bounds__(256)
k_mg_residual(const float* __restrict__ x, const float* __restrict__ b,
              const uint8_t* __restrict__ code, float* __restrict__ r, int nx, int ny, int ld,
              float inv_h2)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i0 >= ld) return;
  const size_t k = i0 + (size_t)j * ld;
  const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + k);
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 != 0u)
  {
    const float4 b4 = *reinterpret_cast<const float4*>(b + k);
    const float4 xc = *reinterpret_cast<const float4*>(x + k);
    const float4 q = mg_apply4(x, ld, ny, i0, j, xc, c4, inv_h2);
    out.x = (c4 & 0xff) ? b4.x - q.x : 0.0f;
    out.y = ((c4 >> 8) & 0xff) ? b4.y - q.y : 0.0f;
    out.z = ((c4 >> 16) & 0xff) ? b4.z - q.z : 0.0f;
    out.w = (c4 >> 24) ? b4.w - q.w : 0.0f;
  }
  *reinterpret_cast<float4*>(r + k) = out;
}

// Transfer weights next to walls.  The plain operators treat a coarse cell that is SOLID (or lies outside
// the grid) as a zero: the bilinear prolongation then under-weights the fine cells along a wall, and -- worse --
// its transpose, the restriction, loses the share of their residuals that would go to the missing coarse
// cell.  A zero-mean high-frequency residual next to a wall so acquires a net mass, which the coarse levels
// answer with a smooth correction of size O(n) times the local error: the largest eigenvalue of M A grows
// like n (1.2 / 1.7 / 3.3 / 6.1 at n = 64 .. 512 on the tank scene, tools/studies/mg_transfer_study.py) and
// the PCG iteration count with it.  Fix: every fine cell's four bilinear weights are renormalised over its
// non-SOLID coarse parents (constant extension across a wall; AIR parents stay in the sum -- their value is
// the Dirichlet zero), P = D^-1 P_bilinear, R = P^T / 4 = R_full-weighting D^-1: conservative at walls and
// still symmetric.  With it (and the hierarchy continued to <= 4 x 4) the count is 7 - 8 from 512^2 to 4096^2.
//
// mg_parent_norm: D for the fine cell whose own parent has the flag `oo`, the parent's neighbour on the
// child's side in x `on`, in y `no`, diagonal `nn` (1.0f = non-SOLID and inside, else 0.0f).
__device__ __forceinline__ float mg_parent_norm(float oo, float on, float no, float nn)
{
  return 0.75f * (0.75f * oo + 0.25f * on) + 0.25f * (0.75f * no + 0.25f * nn);
}

// b_coarse(I,J) = sum_{a,c} W[a] W[c] r_fine(2I-1+c, 2J-1+a) / D_fine, W = (1 3 3 1)/8; one thread per
// coarse cell; fine cells outside the grid contribute 0; 0 on non-LIQUID coarse cells.  clab: labels of the
// COARSE level (nullptr: plain full weighting, D = 1).
__global__ void k_mg_restrict(const float* __restrict__ rf, int fnx, int fny, int fld,
                              const uint8_t* __restrict__ ccode, float* __restrict__ bc, int cnx,
                              int cny, int cld, const uint8_t* __restrict__ clab)
{
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y;
  if (I >= cld) return;
  float out = 0.0f;
  if (I < cnx && ccode[I + (size_t)J * cld] != 0)
  {
    const float W[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    // non-SOLID flags of the 3 x 3 coarse neighbourhood; s[1][1] is this (LIQUID) cell
    float s[3][3];
    bool all = true;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 3; ++q)
      {
        const int II = I - 1 + q, JJ = J - 1 + a;
        const bool ok = !clab || (II >= 0 && II < cnx && JJ >= 0 && JJ < cny && clab[II + (size_t)JJ * cld] != FSB_SOLID);
        s[a][q] = ok ? 1.0f : 0.0f;
        all = all && ok;
      }
    // fine offset 0..3 (row 2J-1+a / column 2I-1+q): own parent and the neighbour on the child's side,
    // as indices into the 3 x 3 window
    const int own[4] = {0, 1, 1, 2}, nb[4] = {1, 0, 2, 1};
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      const int fj = 2 * J - 1 + a;
      if (fj < 0 || fj >= fny) continue;
      float row = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
      {
        const int fi = 2 * I - 1 + q;
        if (fi >= 0 && fi < fnx)
        {
          float v = rf[fi + (size_t)fj * fld];
          if (!all)
          {
            const float d = mg_parent_norm(s[own[a]][own[q]], s[own[a]][nb[q]], s[nb[a]][own[q]], s[nb[a]][nb[q]]);
            v = (d > 0.0f) ? v / d : 0.0f; // d = 0: no non-SOLID parent, the fine cell is SOLID and v is 0
          }
          row += W[q] * v;
        }
      }
      out += W[a] * row;
    }
  }
  bc[I + (size_t)J * cld] = out;
}

// x_fine += D^-1 P e_coarse on LIQUID fine cells (P = 4 R^T: per dimension 3/4 of the parent and 1/4 of
// the parent's neighbour on the child's side); four fine cells per thread.  clab as in k_mg_restrict.
__global__ void __launch_bounds__(256)
k_mg_prolong_add(float* __restrict__ xf, const uint8_t* __restrict__ fcode, int fnx, int fny,
                 int fld, const float* __restrict__ ec, int cnx, int cny, int cld,
                 const uint8_t* __restrict__ clab)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i0 >= fld) return;
  const size_t k = i0 + (size_t)j * fld;
  const uint32_t c4 = *reinterpret_cast<const uint32_t*>(fcode + k);
  if (c4 == 0u) return; // x stays 0 on the four cells
  const int J = j >> 1;
  const int Jn = (j & 1) ? J + 1 : J - 1; // the neighbour row on this child's side
  const int I0 = i0 >> 1;                 // parents of columns i0..i0+3: I0, I0, I0+1, I0+1
  bool all = true;
  auto row_vals = [&](int JJ, float* v, float* f) { // coarse values / non-SOLID flags at columns I0-1 .. I0+2 of row JJ
#pragma unroll
    for (int q = 0; q < 4; ++q)
    {
      const int II = I0 - 1 + q;
      const bool in = JJ >= 0 && JJ < cny && II >= 0 && II < cnx;
      v[q] = in ? ec[II + (size_t)JJ * cld] : 0.0f;
      const bool ok = !clab || (in && clab[II + (size_t)JJ * cld] != FSB_SOLID);
      f[q] = ok ? 1.0f : 0.0f;
      all = all && ok;
    }
  };
  float p[4], n[4], fp[4], fn[4];
  row_vals(J, p, fp);
  row_vals(Jn, n, fn);
  float m[4]; // blended in y
#pragma unroll
  for (int q = 0; q < 4; ++q) m[q] = 0.75f * p[q] + 0.25f * n[q];
  // columns: i0 (even child of I0: neighbour I0-1), i0+1 (odd child of I0: neighbour I0+1),
  //          i0+2 (even child of I0+1: neighbour I0), i0+3 (odd child of I0+1: neighbour I0+2)
  float a0 = 0.75f * m[1] + 0.25f * m[0], a1 = 0.75f * m[1] + 0.25f * m[2];
  float a2 = 0.75f * m[2] + 0.25f * m[1], a3 = 0.75f * m[2] + 0.25f * m[3];
  if (!all)
  {
    // a LIQUID fine cell's own parent is never SOLID, so every D below is at least 9/16
    // (pad columns beyond the grid may have none: their value is not stored)
    auto scaled = [](float a, float d) { return d > 0.0f ? a / d : 0.0f; };
    a0 = scaled(a0, mg_parent_norm(fp[1], fp[0], fn[1], fn[0]));
    a1 = scaled(a1, mg_parent_norm(fp[1], fp[2], fn[1], fn[2]));
    a2 = scaled(a2, mg_parent_norm(fp[2], fp[1], fn[2], fn[1]));
    a3 = scaled(a3, mg_parent_norm(fp[2], fp[3], fn[2], fn[3]));
  }
  float4 x4 = *reinterpret_cast<const float4*>(xf + k);
  if (c4 & 0xff) x4.x += a0;
  if ((c4 >> 8) & 0xff) x4.y += a1;
  if ((c4 >> 16) & 0xff) x4.z += a2;
  if (c4 >> 24) x4.w += a3;
  *reinterpret_cast<float4*>(xf + k) = x4;
}

#ifdef __CUDACC__
// coarsest level: kMgCoarseSweeps damped-Jacobi sweeps from a zero iterate, one CTA, shared memory
__global__ void __launch_bounds__(1024)
k_mg_coarse_solve(const float* __restrict__ b, const uint8_t* __restrict__ code,
                  float* __restrict__ xout, int nx, int ny, int ld, const MgCoef kf, int sweeps)
{
  constexpr int S = kMgCoarsest + 2;
  __shared__ float xs[2][S * S];
  const int t = threadIdx.x;
  const int i = t % kMgCoarsest, j = t / kMgCoarsest;
  for (int q = t; q < S * S; q += blockDim.x) xs[0][q] = xs[1][q] = 0.0f;
  const bool in = i < nx && j < ny;
  const uint32_t cd = in ? code[i + (size_t)j * ld] : 0u;
  const float bb = in ? b[i + (size_t)j * ld] : 0.0f;
  const float wd = mg_wdinv(kf, cd);
  const int o = (i + 1) + (j + 1) * S;
  __syncthreads();
  int cur = 0;
  for (int s = 0; s < sweeps; ++s)
  {
    const float* xi = xs[cur];
    float v = 0.0f;
    if (cd)
    {
      const float xc = xi[o];
      const float ax = kf.inv_h2 * ((xi[o - 1] + xi[o + 1]) + (xi[o - S] + xi[o + S]) - (float)(cd - 1) * xc);
      v = xc + wd * (bb - ax);
    }
    xs[cur ^ 1][o] = v;
    cur ^= 1;
    __syncthreads();
  }
  if (in) xout[i + (size_t)j * ld] = xs[cur][o];
}
#endif // __CUDACC__ (the host emulation runs the same sweeps with k_mg_smooth)

// labels of the next level (any child AIR -> AIR, else any LIQUID -> LIQUID, else SOLID; children
// outside the fine grid are SOLID)
__global__ void k_mg_coarsen_labels(const uint8_t* __restrict__ flab, int fnx, int fny, int fld,
                                    uint8_t* __restrict__ clab, int cnx, int cny, int cld)
{
  const int I = blockIdx.x * blockDim.x + threadIdx.x;
  const int J = blockIdx.y;
  if (I >= cld) return;
  uint8_t out = FSB_SOLID;
  if (I < cnx)
  {
    bool air = false, liq = false;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int q = 0; q < 2; ++q)
      {
        const int fi = 2 * I + q, fj = 2 * J + a;
        if (fi < fnx && fj < fny)
        {
          const uint8_t l = flab[fi + (size_t)fj * fld];
          air |= l == FSB_AIR;
          liq |= l == FSB_LIQUID;
        }
      }
    out = air ? FSB_AIR : (liq ? FSB_LIQUID : FSB_SOLID);
  }
  clab[I + (size_t)J * cld] = out;
}

// stencil code of a coarse level: 0 = not LIQUID, 1 + number of non-SOLID neighbours (outside = SOLID)
__global__ void k_mg_codes(const uint8_t* __restrict__ lab, uint8_t* __restrict__ code, int nx,
                           int ny, int ld)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= ld) return;
  uint8_t cd = 0;
  if (i < nx && lab[i + (size_t)j * ld] == FSB_LIQUID)
  {
    int n = 0;
    n += i > 0 && lab[i - 1 + (size_t)j * ld] != FSB_SOLID;
    n += i + 1 < nx && lab[i + 1 + (size_t)j * ld] != FSB_SOLID;
    n += j > 0 && lab[i + (size_t)(j - 1) * ld] != FSB_SOLID;
    n += j + 1 < ny && lab[i + (size_t)(j + 1) * ld] != FSB_SOLID;
    cd = (uint8_t)(1 + n);
  }
  code[i + (size_t)j * ld] = cd;
}


} // namespace
