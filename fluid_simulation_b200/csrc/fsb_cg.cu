// Pressure projection: src/FluidSolver.cpp:323-483 with Eigen's
// ConjugateGradient (Jacobi preconditioner, lower-triangular self-adjoint
// product) replaced by a matrix-free solver on full-grid
// masked vectors.
//
// Layout: x, r, p, q are fp32 grids with pitch ld, exactly zero outside LIQUID
// cells; `code` is one byte per cell: 0 = not liquid, 1 + n for a liquid cell
// with n non-SOLID neighbours (src/FluidSolver.cpp:378-410).  Because p is
// zero on non-liquid cells, sum_{LIQUID nbrs} p equals the plain 4-neighbour
// sum, so the operator needs no neighbour bits.
//
// One CG iteration = two kernels (Eigen's statement order is kept):
//   k_cg_direction : p = z + beta p (z = invdiag r; p = z on the first pass) on a
//                    tile and its halo, q = A p in registers, partial p.q
//                    -> 9 B read + 4 B written per cell (q is never stored)
//   k_cg_update    : alpha = absNew / p.q; q = A p recomputed from the staged
//                    p tile; x += alpha p; r -= alpha q; partial |r|^2, r.z
//                    -> 13 B read + 8 B written per cell
// 34 B of HBM traffic per cell and iteration against the 45 B of the textbook
// formulation.  The last block to finish each kernel folds the
// per-block partials in a fixed order (deterministic) and advances the
// device-resident scalars, so there is no host round trip inside the loop;
// the loop is a CUDA graph of kCheckEvery iterations and the host polls
// `done` one chunk behind the GPU.  Once `done` is set every later launch
// returns immediately, so x and the iteration count are exactly those of the
// converging iteration.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"

namespace {

constexpr int kCheckEvery = 32;

// ---------------------------------------------------------------- helpers --
__device__ __forceinline__ bool last_block_done(unsigned int* ticket, unsigned int nblocks)
{
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == nblocks - 1);
  }
  __syncthreads();
  return s_last;
}

// deterministic fold of `n` partials (stride `stride` doubles apart) by one block
__device__ __forceinline__ double fold_partials(const volatile double* part, int n)
{
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += part[k];
  return block_sum(s);
}

// --------------------------------------------------------- system set-up --
// src/FluidSolver.cpp:329-346,368-416: stencil code, right-hand side
// b = divVelX + divVelY (include/MacGrid.h:98-111) on LIQUID cells, x = 0,
// r = b, and the first dot products (|b|^2, b.z).
__global__ void k_cg_build(const float* __restrict__ uf, const float* __restrict__ vf,
                           const uint8_t* __restrict__ cell, uint8_t* __restrict__ code,
                           float* __restrict__ x, float* __restrict__ r, const GridDims d,
                           const CgCoef coef, CgScalars* __restrict__ s,
                           double* __restrict__ partials, float tol, int max_iters)
{
  const int64_t total = (int64_t)d.ld * d.ny;
  double acc_b2 = 0.0, acc_bz = 0.0, acc_n = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x)
  {
    const int i = (int)(t % d.ld);
    const int j = (int)(t / d.ld);
    uint8_t cd = 0;
    float b = 0.0f;
    if (i < d.nx && j < d.ny && cell[t] == FSB_LIQUID)
    {
      int n = 0;
      n += cell_type(cell, d, i - 1, j) != FSB_SOLID;
      n += cell_type(cell, d, i + 1, j) != FSB_SOLID;
      n += cell_type(cell, d, i, j - 1) != FSB_SOLID;
      n += cell_type(cell, d, i, j + 1) != FSB_SOLID;
      cd = (uint8_t)(1 + n);
      // a LIQUID cell is never on the border in a classified grid; clamp for safety
      const int ie = min(i + 1, d.nx - 1), jn = min(j + 1, d.ny - 1);
      b = (uf[ie + (size_t)j * d.ld] - uf[t]) / d.dx + (vf[i + (size_t)jn * d.ld] - vf[t]) / d.dy;
      const float z = coef.invdiag[n] * b;
      acc_b2 += (double)b * (double)b;
      acc_bz += (double)b * (double)z;
      acc_n += 1.0;
    }
    code[t] = cd;
    x[t] = 0.0f;
    r[t] = b;
  }
  const double b2 = block_sum(acc_b2);
  const double bz = block_sum(acc_bz);
  const double nn = block_sum(acc_n);
  if (threadIdx.x == 0)
  {
    partials[blockIdx.x] = b2;
    partials[gridDim.x + blockIdx.x] = bz;
    partials[2 * gridDim.x + blockIdx.x] = nn;
  }
  if (last_block_done(&s->ticket[0], gridDim.x))
  {
    const double tb2 = fold_partials(partials, gridDim.x);
    const double tbz = fold_partials(partials + gridDim.x, gridDim.x);
    const double tn = fold_partials(partials + 2 * gridDim.x, gridDim.x);
    if (threadIdx.x == 0)
    {
      const float rhs2 = (float)tb2;
      s->rhs2 = tb2;
      s->r2 = tb2;
      s->rz = tbz;
      s->pq = 0.0;
      s->n_liquid = (int)tn;
      s->tol = tol;
      s->max_iters = max_iters < 0 ? 2 * (int)tn : max_iters;
      // Eigen: threshold = max(tol*tol*rhsNorm2, FLT_MIN)
      float thr = tol * tol * rhs2;
      if (thr < FLT_MIN) thr = FLT_MIN;
      s->thr = thr;
      s->abs_new = (float)tbz;
      s->abs_old = 1.0f;
      s->beta = 0.0f;
      s->iter = 0;
      // rhsNorm2 == 0 -> x = 0, 0 iterations; |r|^2 < threshold -> 0 iterations;
      // maxIters == 0 -> the while loop never runs
      s->done = (rhs2 == 0.0f || rhs2 < thr || s->max_iters <= 0) ? 1 : 0;
      s->ticket[0] = 0;
    }
  }
}

// ------------------------------------------------------------ tile frame --
// Both iteration kernels work on tiles of kTileW columns x TH rows.  A CTA of
// 256 threads = 8 warps; lane l owns the four columns 4l..4l+3 of the tile
// (one 16-byte access per row), warp w owns rows w, w+8, ... of the tile.
// The search direction of the tile and of its one-cell halo is staged in
// shared memory (row pitch kPitch floats, interior starts at word 4 so the
// float4 stores stay 16-byte aligned); north/south neighbours are read from
// there, west/east neighbours come from the adjacent lanes (shuffle) and from
// the halo columns for lanes 0 and 31.  CTAs walk the tile list with a grid
// stride, so the number of per-CTA partial sums is bounded by the grid size.
constexpr int kTileW = 128;
constexpr int kPitch = kTileW + 8;
constexpr int kCgThreads = 256;
constexpr int kCgWarps = kCgThreads / 32;

// Per-cell coefficients come from a 2 x 8 shared-memory table indexed by the
// stencil code (0 = not liquid -> coefficient 0, 1 + n -> n non-SOLID
// neighbours): [0][code] = Jacobi inverse diagonal, [1][code] = diagonal.  A
// dynamically indexed kernel-parameter array would compile to indexed LDC,
// which issues on the slow XU pipe (measured: 78 % XU-bound, profiles/r01b).
struct CgLut
{
  float inv[8];
  float diag[8];
};

__device__ __forceinline__ void load_lut(CgLut* lut, const CgCoef& coef)
{
  if (threadIdx.x < 8)
  {
    const int t = threadIdx.x;
    float iv = 0.0f, dg = 0.0f;
    if (t >= 1 && t <= 5)
    {
      iv = coef.invdiag[t - 1];
      dg = coef.diag[t - 1];
    }
    lut->inv[t] = iv;
    lut->diag[t] = dg;
  }
}

// r and p_old are exactly zero on non-liquid cells and inv[0] = 0, so the
// result is exactly zero there without a branch.
__device__ __forceinline__ float dir_value(float r, float p_old, uint32_t cd, const CgLut* lut,
                                           float beta, bool first)
{
  const float z = lut->inv[cd] * r;
  return first ? z : z + beta * p_old;
}

__device__ __forceinline__ float4 dir_value4(const float4 r4, const float4 p4, uint32_t c4,
                                             const CgLut* lut, float beta, bool first)
{
  float4 o;
  o.x = dir_value(r4.x, p4.x, c4 & 0xff, lut, beta, first);
  o.y = dir_value(r4.y, p4.y, (c4 >> 8) & 0xff, lut, beta, first);
  o.z = dir_value(r4.z, p4.z, (c4 >> 16) & 0xff, lut, beta, first);
  o.w = dir_value(r4.w, p4.w, c4 >> 24, lut, beta, first);
  return o;
}

__device__ __forceinline__ float apply_a(float c, float w, float e, float s, float n, uint32_t cd,
                                         float off, const CgLut* lut)
{
  // each coefficient multiplies its own operand, as a sparse product does
  float acc = off * w;
  acc += off * e;
  acc += off * s;
  acc += off * n;
  acc += lut->diag[cd] * c;
  return cd ? acc : 0.0f;
}

// q = A p for the four cells of one lane in tile row `tr` (0-based inside the
// tile); `pc` is the lane's own direction, sp the staged tile.
__device__ __forceinline__ float4 apply_a4(const float* __restrict__ sp, int tr, unsigned lane,
                                           const float4 pc, uint32_t c4, float off,
                                           const CgLut* lut)
{
  const float* row = sp + (tr + 1) * kPitch + 4 + lane * 4;
  const float4 s4 = *reinterpret_cast<const float4*>(row - kPitch);
  const float4 n4 = *reinterpret_cast<const float4*>(row + kPitch);
  float w = __shfl_up_sync(0xffffffffu, pc.w, 1);
  float e = __shfl_down_sync(0xffffffffu, pc.x, 1);
  if (lane == 0) w = row[-1];
  if (lane == 31) e = row[4];
  float4 q;
  q.x = apply_a(pc.x, w, pc.y, s4.x, n4.x, c4 & 0xff, off, lut);
  q.y = apply_a(pc.y, pc.x, pc.z, s4.y, n4.y, (c4 >> 8) & 0xff, off, lut);
  q.z = apply_a(pc.z, pc.y, pc.w, s4.z, n4.z, (c4 >> 16) & 0xff, off, lut);
  q.w = apply_a(pc.w, pc.z, e, s4.w, n4.w, c4 >> 24, off, lut);
  return q;
}

// ------------------------------------------------- direction + p.Ap dot --
// p_new = z + beta p_old on the tile and its halo (the halo is RE-COMPUTED
// from r, p_old and code, so no other CTA's output is needed and p_new can be
// produced and consumed in the same kernel); q = A p_new stays in registers
// and only feeds the dot product.  13 B of HBM traffic per cell: r, p_old,
// code in, p_new out (ping-pong with p_old because neighbours still need the
// old halo).
template <int TH>
__global__ void __launch_bounds__(kCgThreads, 4)
k_cg_direction(const float* __restrict__ p_old, float* __restrict__ p_new,
               const float* __restrict__ r, const uint8_t* __restrict__ code, int ld, int ny,
               int tiles_x, int n_tiles, const CgCoef coef, CgScalars* __restrict__ s,
               double* __restrict__ partials)
{
  if (s->done) return;
  __shared__ __align__(16) float sp[(TH + 2) * kPitch];
  __shared__ CgLut lut;
  load_lut(&lut, coef);
  __syncthreads();
  const float off = coef.off;
  constexpr int RPW = TH / kCgWarps; // rows per warp
  const bool first = (s->iter == 0);
  const float beta = s->beta;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  double acc = 0.0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const int c0 = (tile % tiles_x) * kTileW;
    const int j0 = (tile / tiles_x) * TH;
    const int ci = c0 + (int)lane * 4;
    const bool col_ok = ci < ld;
    float4 pn[RPW];
    uint32_t cd[RPW];
    // owned rows
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      const int tr = (int)warp + k * kCgWarps;
      const int j = j0 + tr;
      pn[k] = zero4;
      cd[k] = 0;
      if (col_ok && j < ny)
      {
        const size_t o = (size_t)j * ld + ci;
        const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + o);
        const float4 r4 = *reinterpret_cast<const float4*>(r + o);
        float4 p4 = zero4;
        if (!first) p4 = *reinterpret_cast<const float4*>(p_old + o);
        cd[k] = c4;
        pn[k] = dir_value4(r4, p4, c4, &lut, beta, first);
      }
      *reinterpret_cast<float4*>(sp + (tr + 1) * kPitch + 4 + lane * 4) = pn[k];
    }
    // halo rows j0-1 (warp 0) and j0+TH (warp 1)
    if (warp < 2)
    {
      const int j = (warp == 0) ? j0 - 1 : j0 + TH;
      float4 h = zero4;
      if (col_ok && j >= 0 && j < ny)
      {
        const size_t o = (size_t)j * ld + ci;
        const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + o);
        if (c4)
        {
          const float4 r4 = *reinterpret_cast<const float4*>(r + o);
          float4 p4 = zero4;
          if (!first) p4 = *reinterpret_cast<const float4*>(p_old + o);
          h = dir_value4(r4, p4, c4, &lut, beta, first);
        }
      }
      *reinterpret_cast<float4*>(sp + ((warp == 0) ? 0 : (TH + 1)) * kPitch + 4 + lane * 4) = h;
    }
    // halo columns c0-1 and c0+kTileW of the TH tile rows: threads 64 .. 64+2*TH-1
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * TH)
    {
      const int t = threadIdx.x - 64;
      const int tr = t >> 1, side = t & 1;
      const int j = j0 + tr;
      const int cc = side ? c0 + kTileW : c0 - 1;
      float h = 0.0f;
      if (j < ny && cc >= 0 && cc < ld)
      {
        const size_t o = (size_t)j * ld + cc;
        const uint32_t c1 = code[o];
        if (c1) h = dir_value(r[o], first ? 0.0f : p_old[o], c1, &lut, beta, first);
      }
      sp[(tr + 1) * kPitch + (side ? 4 + kTileW : 3)] = h;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      const int tr = (int)warp + k * kCgWarps;
      const int j = j0 + tr;
      const float4 q = apply_a4(sp, tr, lane, pn[k], cd[k], off, &lut);
      if (col_ok && j < ny)
      {
        *reinterpret_cast<float4*>(p_new + (size_t)j * ld + ci) = pn[k];
        acc += (double)((pn[k].x * q.x + pn[k].y * q.y) + (pn[k].z * q.z + pn[k].w * q.w));
      }
    }
    __syncthreads(); // the tile buffer is reused
  }

  const double tot = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  if (last_block_done(&s->ticket[1], gridDim.x))
  {
    const double pq = fold_partials(partials, (int)gridDim.x);
    if (threadIdx.x == 0)
    {
      s->pq = pq;
      s->ticket[1] = 0;
    }
  }
}

// ------------------------------------------------------------ the update --
// alpha = absNew / p.Ap; x += alpha p; r -= alpha A p with A p RE-COMPUTED
// from the staged p tile (q is never stored); partial |r|^2 and r.z.
// 21 B of HBM traffic per cell: p, code, x, r in; x, r out.
template <int TH>
__global__ void __launch_bounds__(kCgThreads, TH >= 32 ? 3 : 4)
k_cg_update(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p,
            const uint8_t* __restrict__ code, int ld, int ny, int tiles_x, int n_tiles,
            const CgCoef coef, CgScalars* __restrict__ s, double* __restrict__ partials)
{
  if (s->done) return;
  __shared__ __align__(16) float sp[(TH + 2) * kPitch];
  __shared__ CgLut lut;
  load_lut(&lut, coef);
  __syncthreads();
  const float off = coef.off;
  constexpr int RPW = TH / kCgWarps;
  const float alpha = s->abs_new / (float)s->pq; // Eigen: alpha = absNew / p.dot(tmp)
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  double acc_r2 = 0.0, acc_rz = 0.0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    const int c0 = (tile % tiles_x) * kTileW;
    const int j0 = (tile / tiles_x) * TH;
    const int ci = c0 + (int)lane * 4;
    const bool col_ok = ci < ld;
    float4 pc[RPW], xv[RPW], rv[RPW];
    uint32_t cd[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      const int tr = (int)warp + k * kCgWarps;
      const int j = j0 + tr;
      pc[k] = zero4;
      xv[k] = zero4;
      rv[k] = zero4;
      cd[k] = 0;
      if (col_ok && j < ny)
      {
        const size_t o = (size_t)j * ld + ci;
        cd[k] = *reinterpret_cast<const uint32_t*>(code + o);
        pc[k] = *reinterpret_cast<const float4*>(p + o);
        if (cd[k])
        {
          xv[k] = *reinterpret_cast<const float4*>(x + o);
          rv[k] = *reinterpret_cast<const float4*>(r + o);
        }
      }
      *reinterpret_cast<float4*>(sp + (tr + 1) * kPitch + 4 + lane * 4) = pc[k];
    }
    if (warp < 2)
    {
      const int j = (warp == 0) ? j0 - 1 : j0 + TH;
      float4 h = zero4;
      if (col_ok && j >= 0 && j < ny) h = *reinterpret_cast<const float4*>(p + (size_t)j * ld + ci);
      *reinterpret_cast<float4*>(sp + ((warp == 0) ? 0 : (TH + 1)) * kPitch + 4 + lane * 4) = h;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * TH)
    {
      const int t = threadIdx.x - 64;
      const int tr = t >> 1, side = t & 1;
      const int j = j0 + tr;
      const int cc = side ? c0 + kTileW : c0 - 1;
      float h = 0.0f;
      if (j < ny && cc >= 0 && cc < ld) h = p[(size_t)j * ld + cc];
      sp[(tr + 1) * kPitch + (side ? 4 + kTileW : 3)] = h;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      const int tr = (int)warp + k * kCgWarps;
      const int j = j0 + tr;
      const uint32_t c4 = cd[k];
      const float4 q = apply_a4(sp, tr, lane, pc[k], c4, off, &lut); // shuffles: whole warp takes part
      if (c4 != 0 && col_ok && j < ny) // four non-liquid cells: x, r stay exactly zero
      {
        float4 xo = xv[k], ro = rv[k];
        xo.x = xo.x + alpha * pc[k].x; ro.x = ro.x - alpha * q.x;
        xo.y = xo.y + alpha * pc[k].y; ro.y = ro.y - alpha * q.y;
        xo.z = xo.z + alpha * pc[k].z; ro.z = ro.z - alpha * q.z;
        xo.w = xo.w + alpha * pc[k].w; ro.w = ro.w - alpha * q.w;
        const size_t o = (size_t)j * ld + ci;
        *reinterpret_cast<float4*>(x + o) = xo;
        *reinterpret_cast<float4*>(r + o) = ro;
        const float z0 = lut.inv[c4 & 0xff] * ro.x;
        const float z1 = lut.inv[(c4 >> 8) & 0xff] * ro.y;
        const float z2 = lut.inv[(c4 >> 16) & 0xff] * ro.z;
        const float z3 = lut.inv[c4 >> 24] * ro.w;
        acc_r2 += (double)((ro.x * ro.x + ro.y * ro.y) + (ro.z * ro.z + ro.w * ro.w));
        acc_rz += (double)((ro.x * z0 + ro.y * z1) + (ro.z * z2 + ro.w * z3));
      }
    }
    __syncthreads();
  }

  const double r2 = block_sum(acc_r2);
  const double rz = block_sum(acc_rz);
  if (threadIdx.x == 0)
  {
    partials[blockIdx.x] = r2;
    partials[gridDim.x + blockIdx.x] = rz;
  }
  if (last_block_done(&s->ticket[2], gridDim.x))
  {
    const double tr2 = fold_partials(partials, gridDim.x);
    const double trz = fold_partials(partials + gridDim.x, gridDim.x);
    if (threadIdx.x == 0)
    {
      s->r2 = tr2;
      s->rz = trz;
      if ((float)tr2 < s->thr)
      {
        s->done = 1; // converged: Eigen breaks before i++
      }
      else
      {
        s->abs_old = s->abs_new;
        s->abs_new = (float)trz;
        s->beta = s->abs_new / s->abs_old;
        s->iter = s->iter + 1;
        if (s->iter >= s->max_iters) s->done = 1;
      }
      s->ticket[2] = 0;
    }
  }
}

// ------------------------------------------------------- velocity patch --
// src/FluidSolver.cpp:428-482: faces touching a LIQUID cell get
// front - ((dt/density) * dp) / delta written into the BACK buffer; all other
// faces keep whatever the back buffer held.
__global__ void k_pressure_patch(const float* __restrict__ uf, const float* __restrict__ vf,
                                 float* __restrict__ ub, float* __restrict__ vb,
                                 const float* __restrict__ x, const uint8_t* __restrict__ code,
                                 const GridDims d, float dt, float density)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(t % d.ld);
  const int j = (int)(t / d.ld);
  if (i >= d.nx || j >= d.ny) return;
  const int im1 = clampi(i - 1, 0, d.nx - 1);
  const int jm1 = clampi(j - 1, 0, d.ny - 1);
  const size_t k = t, kw = im1 + (size_t)j * d.ld, ks = i + (size_t)jm1 * d.ld;
  const bool l = code[k] != 0, lw = code[kw] != 0, ls = code[ks] != 0;
  if (!(l || lw || ls)) return;
  // the particle-pressure terms are k * n with k = 0.0 (:443-453): exactly +0
  const float pc = l ? x[k] + 0.0f : 0.0f;
  const float pw = lw ? x[kw] + 0.0f : 0.0f;
  const float ps = ls ? x[ks] + 0.0f : 0.0f;
  const float ddx = pc - pw;
  const float ddy = pc - ps;
  ub[k] = uf[k] - ((dt / density) * ddx) / d.dx;
  vb[k] = vf[k] - ((dt / density) * ddy) / d.dy;
}

} // namespace

namespace {

CgCoef make_coef(const fsb_ctx* c)
{
  CgCoef coef;
  const double dx2 = std::pow((double)c->dx, 2);
  coef.off = (float)(1 / dx2);
  for (int n = 0; n < 5; ++n)
  {
    coef.diag[n] = (float)(-n / dx2);
    coef.invdiag[n] = (coef.diag[n] != 0.0f) ? 1.0f / coef.diag[n] : 1.0f;
  }
  return coef;
}

// Tile height: the tallest of 32/16/8 rows that still gives every SM several tiles.
int pick_tile_rows(const fsb_ctx* c)
{
  if (const char* e = getenv("FSB_CG_TILE_ROWS")) // tuning knob for profiling runs
  {
    const int th = atoi(e);
    if (th == 8 || th == 16 || th == 32) return th;
  }
  const int tiles_x = fsb_div_up(c->ld, kTileW);
  for (int th = 32; th > 8; th >>= 1)
    if ((int64_t)tiles_x * fsb_div_up(c->ny, th) >= (int64_t)4 * c->sm_count) return th;
  return 8;
}

// persistent-style grids: exactly the CTAs that are co-resident (one wave), each walking the
// tile list with a grid stride
template <int TH>
void resident_grids(fsb_ctx* c, int64_t n_tiles)
{
  int occ_dir = 4, occ_upd = 3;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, k_cg_direction<TH>, kCgThreads, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, k_cg_update<TH>, kCgThreads, 0);
  c->cg_grid_dir = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * std::max(occ_dir, 1));
  c->cg_grid_upd = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * std::max(occ_upd, 1));
}

// one CG iteration = two launches; `cur` selects the ping-pong direction buffer
int launch_iteration(fsb_ctx* c, const CgCoef& coef, int cur)
{
  const int th = c->cg_tile_rows;
  const int tiles_x = fsb_div_up(c->ld, kTileW);
  const int n_tiles = tiles_x * fsb_div_up(c->ny, th);
#define FSB_CG_LAUNCH(TH)                                                                          \
  k_cg_direction<TH><<<c->cg_grid_dir, kCgThreads, 0, c->stream>>>(                                \
      c->cg_p[cur], c->cg_p[cur ^ 1], c->cg_r, c->cg_code, c->ld, c->ny, tiles_x, n_tiles, coef,   \
      c->scal, c->partials);                                                                       \
  k_cg_update<TH><<<c->cg_grid_upd, kCgThreads, 0, c->stream>>>(                                   \
      c->cg_x, c->cg_r, c->cg_p[cur ^ 1], c->cg_code, c->ld, c->ny, tiles_x, n_tiles, coef,        \
      c->scal, c->partials)
  if (th == 32) { FSB_CG_LAUNCH(32); }
  else if (th == 16) { FSB_CG_LAUNCH(16); }
  else { FSB_CG_LAUNCH(8); }
#undef FSB_CG_LAUNCH
  FSB_CUDA(c, cudaGetLastError());
  return FSB_OK;
}

// kCheckEvery iterations captured once per context into a CUDA graph (all
// arguments are fixed for the life of the context; alpha, beta, the first-pass
// flag and `done` live in device memory).
int ensure_cg_graph(fsb_ctx* c, const CgCoef& coef)
{
  if (c->cg_graph_state != 0) return FSB_OK;
  c->cg_graph_state = -1; // direct launches unless the capture below succeeds
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
  {
    cudaGetLastError();
    return FSB_OK;
  }
  int rc = FSB_OK;
  for (int k = 0; k < kCheckEvery && rc == FSB_OK; ++k) rc = launch_iteration(c, coef, k & 1);
  const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (rc != FSB_OK || e != cudaSuccess || !graph)
  {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return FSB_OK;
  }
  if (cudaGraphInstantiate(&c->cg_graph, graph, 0) == cudaSuccess) c->cg_graph_state = 1;
  else cudaGetLastError();
  cudaGraphDestroy(graph);
  return FSB_OK;
}

int launch_chunk(fsb_ctx* c, const CgCoef& coef)
{
  static_assert(kCheckEvery % 2 == 0, "a chunk must leave the ping-pong where it started");
  if (c->cg_graph_state == 1)
  {
    FSB_CUDA(c, cudaGraphLaunch(c->cg_graph, c->stream));
  }
  else
  {
    for (int k = 0; k < kCheckEvery; ++k) FSB_TRY(launch_iteration(c, coef, k & 1));
  }
  c->launches += 2 * kCheckEvery;
  return FSB_OK;
}

} // namespace

int fsb_k_pressure_solve(fsb_ctx* c, float density, float dt)
{
  const GridDims d{c->nx, c->ny, c->ld, c->dx, c->dy};
  const CgCoef coef = make_coef(c);

  // ---- build
  fsb_prof_begin(c, FSB_PROF_RHS);
  const int64_t total = (int64_t)c->ld * c->ny;
  const int build_blocks = (int)std::min<int64_t>(fsb_div_up(total, 256), c->sm_count * 8);
  if (c->cg_tile_rows == 0)
  {
    c->cg_tile_rows = pick_tile_rows(c);
    const int64_t n_tiles =
        (int64_t)fsb_div_up(c->ld, kTileW) * fsb_div_up(c->ny, c->cg_tile_rows);
    if (c->cg_tile_rows == 32) resident_grids<32>(c, n_tiles);
    else if (c->cg_tile_rows == 16) resident_grids<16>(c, n_tiles);
    else resident_grids<8>(c, n_tiles);
  }
  const int need = std::max(3 * build_blocks, std::max(c->cg_grid_dir, 2 * c->cg_grid_upd));
  if (need > c->partials_cap)
  {
    if (c->partials) cudaFree(c->partials);
    c->partials = nullptr;
    FSB_CUDA(c, cudaMalloc(&c->partials, sizeof(double) * need));
    c->partials_cap = need;
  }
  k_cg_build<<<build_blocks, 256, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell, c->cg_code,
                                                  c->cg_x, c->cg_r, d, coef, c->scal, c->partials,
                                                  c->tol, c->max_iters);
  FSB_LAUNCHED(c);
  FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[0], c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                              c->stream));
  fsb_prof_end(c, FSB_PROF_RHS);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->scal_h[0].n_liquid == 0) return FSB_OK; // :347-350: nothing touched, no swap

  // ---- iterate: chunks of kCheckEvery iterations; the host reads the device
  // scalars of chunk k while chunk k+1 is already queued, so the GPU never
  // waits for the poll.  Launches after convergence return immediately.
  fsb_prof_begin(c, FSB_PROF_CG);
  CgScalars fin = c->scal_h[0];
  if (!fin.done)
  {
    FSB_TRY(ensure_cg_graph(c, coef));
    const int max_chunks = fsb_div_up(std::max(fin.max_iters, 1), kCheckEvery);
    int queued = 0, slot = 0;
    auto queue_chunk = [&]() -> int {
      FSB_TRY(launch_chunk(c, coef));
      FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[slot], c->scal, sizeof(CgScalars),
                                  cudaMemcpyDeviceToHost, c->stream));
      FSB_CUDA(c, cudaEventRecord(c->cg_ev[slot], c->stream));
      slot ^= 1;
      ++queued;
      return FSB_OK;
    };
    FSB_TRY(queue_chunk());
    for (;;)
    {
      const bool more = queued < max_chunks;
      if (more) FSB_TRY(queue_chunk());
      // the older of the (up to) two chunks in flight
      const int wait_slot = more ? slot : slot ^ 1;
      FSB_CUDA(c, cudaEventSynchronize(c->cg_ev[wait_slot]));
      fin = c->scal_h[wait_slot];
      if (fin.done) break;
      if (!more)
      {
        // cannot happen: max_chunks chunks always reach the iteration cap
        return fsb_fail(c, FSB_ERR_CUDA, "CG did not terminate after %d chunks", queued);
      }
    }
  }
  fsb_prof_end(c, FSB_PROF_CG);
  c->iters = fin.iter;
  c->err = (fin.rhs2 == 0.0 || (float)fin.rhs2 == 0.0f)
               ? 0.0f
               : std::sqrt((float)fin.r2 / (float)fin.rhs2);
  c->pressure_valid = true;

  // ---- patch + swap
  fsb_prof_begin(c, FSB_PROF_PATCH);
  k_pressure_patch<<<fsb_div_up(total, 256), 256, 0, c->stream>>>(
      fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), c->cg_x, c->cg_code, d, dt, density);
  FSB_LAUNCHED(c);
  c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:482
  fsb_prof_end(c, FSB_PROF_PATCH);
  return FSB_OK;
}
