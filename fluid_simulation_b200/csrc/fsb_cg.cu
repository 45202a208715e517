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
//   k_cg_dir_spmv : p = z + beta p (z = invdiag r; p = z on the first pass),
//                   q = A p computed from p in registers (3-row sliding window),
//                   partial p.q           -> 9 B read + 8 B written per cell
//   k_cg_update   : alpha = absNew / p.q; x += alpha p; r -= alpha q;
//                   partial |r|^2 and r.z -> 17 B read + 8 B written per cell
// The last block to finish each kernel folds the per-block partials in a fixed
// order (deterministic) and advances the device-resident scalars, so there is
// no host round trip inside the loop; the host polls `done` every
// kCheckEvery iterations.  Once `done` is set every later launch returns
// immediately, so x and the iteration count are exactly those of the
// converging iteration.
#include <cfloat>
#include <cmath>

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

// --------------------------------------------------- direction + product --
// Strip-marching stencil: each thread owns 4 consecutive columns and walks
// down kRows rows keeping the new search direction of three rows in registers,
// so p and r are read once per band (plus two halo rows) and q never needs p
// from memory.  West/east neighbours come from the adjacent lanes; the two
// edge lanes of a warp recompute their outer neighbour from r, p and code.
// The old direction is read from `p` and the new one written to `p_out`
// (ping-pong): halo rows and edge columns belong to other blocks, which may
// already have produced their new values.
constexpr int kSpmvThreads = 128;
constexpr int kRows = 32;

struct Row4
{
  float4 p; // new direction
  float w, e; // west neighbour of p.x, east neighbour of p.w
  uint32_t code;
};

__device__ __forceinline__ float dir_value(float r, float p_old, uint32_t cd, const CgCoef& coef,
                                           float beta, bool first)
{
  if (cd == 0) return 0.0f;
  const float z = coef.invdiag[cd - 1] * r;
  return first ? z : z + beta * p_old;
}

__device__ __forceinline__ void load_row(Row4& row, const float* __restrict__ r,
                                         const float* __restrict__ p,
                                         const uint8_t* __restrict__ code, int ci, int jj, int ld,
                                         int ny, const CgCoef& coef, float beta, bool first,
                                         unsigned lane)
{
  row.p = make_float4(0.f, 0.f, 0.f, 0.f);
  row.w = 0.f;
  row.e = 0.f;
  row.code = 0;
  const bool row_ok = (jj >= 0 && jj < ny);
  const bool col_ok = ci < ld;
  float edge = 0.0f;
  if (row_ok)
  {
    const size_t base = (size_t)jj * ld;
    if (col_ok)
    {
      const float4 r4 = *reinterpret_cast<const float4*>(r + base + ci);
      float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!first) p4 = *reinterpret_cast<const float4*>(p + base + ci);
      const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + base + ci);
      row.code = c4;
      row.p.x = dir_value(r4.x, p4.x, c4 & 0xff, coef, beta, first);
      row.p.y = dir_value(r4.y, p4.y, (c4 >> 8) & 0xff, coef, beta, first);
      row.p.z = dir_value(r4.z, p4.z, (c4 >> 16) & 0xff, coef, beta, first);
      row.p.w = dir_value(r4.w, p4.w, (c4 >> 24) & 0xff, coef, beta, first);
    }
    // outer neighbour of the warp's 128-column span: lane 0 -> column ci-1,
    // lane 31 -> column ci+4
    if (lane == 0 || lane == 31)
    {
      const int ce = (lane == 0) ? ci - 1 : ci + 4;
      if (ce >= 0 && ce < ld)
      {
        const uint32_t cd = code[base + ce];
        if (cd) edge = dir_value(r[base + ce], first ? 0.0f : p[base + ce], cd, coef, beta, first);
      }
    }
  }
  float w = __shfl_up_sync(0xffffffffu, row.p.w, 1);
  float e = __shfl_down_sync(0xffffffffu, row.p.x, 1);
  if (lane == 0) w = edge;
  if (lane == 31) e = edge;
  row.w = w;
  row.e = e;
}

__device__ __forceinline__ float apply_a(float c, float w, float e, float s, float n, uint32_t cd,
                                         const CgCoef& coef)
{
  if (cd == 0) return 0.0f;
  // each coefficient multiplies its own operand, as a sparse product does
  float acc = coef.off * w;
  acc += coef.off * e;
  acc += coef.off * s;
  acc += coef.off * n;
  acc += coef.diag[cd - 1] * c;
  return acc;
}

__global__ void __launch_bounds__(kSpmvThreads)
k_cg_dir_spmv(const float* __restrict__ p, float* __restrict__ p_out, float* __restrict__ q,
              const float* __restrict__ r, const uint8_t* __restrict__ code, int ld, int ny,
              const CgCoef coef, CgScalars* __restrict__ s, double* __restrict__ partials)
{
  if (s->done) return;
  const bool first = (s->iter == 0);
  const float beta = s->beta;
  const unsigned lane = threadIdx.x & 31;
  const int ci = (blockIdx.x * kSpmvThreads + threadIdx.x) * 4;
  const int j0 = blockIdx.y * kRows;
  const int j1 = min(j0 + kRows, ny);
  const bool col_ok = ci < ld;

  Row4 a, b, c; // rows jj-2, jj-1, jj
  load_row(a, r, p, code, ci, j0 - 1, ld, ny, coef, beta, first, lane);
  load_row(b, r, p, code, ci, j0, ld, ny, coef, beta, first, lane);
  double acc = 0.0;
  for (int jj = j0 + 1; jj <= j1; ++jj)
  {
    load_row(c, r, p, code, ci, jj, ld, ny, coef, beta, first, lane);
    // finish row jj-1 (= b): its south is a, its north is c
    if (col_ok)
    {
      const size_t o = (size_t)(jj - 1) * ld + ci;
      float4 qv;
      qv.x = apply_a(b.p.x, b.w, b.p.y, a.p.x, c.p.x, b.code & 0xff, coef);
      qv.y = apply_a(b.p.y, b.p.x, b.p.z, a.p.y, c.p.y, (b.code >> 8) & 0xff, coef);
      qv.z = apply_a(b.p.z, b.p.y, b.p.w, a.p.z, c.p.z, (b.code >> 16) & 0xff, coef);
      qv.w = apply_a(b.p.w, b.p.z, b.e, a.p.w, c.p.w, (b.code >> 24) & 0xff, coef);
      *reinterpret_cast<float4*>(p_out + o) = b.p;
      *reinterpret_cast<float4*>(q + o) = qv;
      acc += (double)b.p.x * (double)qv.x + (double)b.p.y * (double)qv.y +
             (double)b.p.z * (double)qv.z + (double)b.p.w * (double)qv.w;
    }
    a = b;
    b = c;
  }

  const unsigned int nblocks = gridDim.x * gridDim.y;
  const unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) partials[bid] = tot;
  if (last_block_done(&s->ticket[1], nblocks))
  {
    const double pq = fold_partials(partials, (int)nblocks);
    if (threadIdx.x == 0)
    {
      s->pq = pq;
      s->ticket[1] = 0;
    }
  }
}

// ------------------------------------------------------------ the update --
constexpr int kUpdThreads = 256;

__global__ void __launch_bounds__(kUpdThreads)
k_cg_update(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p,
            const float* __restrict__ q, const uint8_t* __restrict__ code, int64_t n4,
            const CgCoef coef, CgScalars* __restrict__ s, double* __restrict__ partials)
{
  if (s->done) return;
  const float alpha = s->abs_new / (float)s->pq; // Eigen: alpha = absNew / p.dot(tmp)
  double acc_r2 = 0.0, acc_rz = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4;
       t += (int64_t)gridDim.x * blockDim.x)
  {
    const uint32_t c4 = reinterpret_cast<const uint32_t*>(code)[t];
    if (c4 == 0) continue; // four non-liquid cells: x, r stay exactly zero
    float4 xv = reinterpret_cast<float4*>(x)[t];
    float4 rv = reinterpret_cast<float4*>(r)[t];
    const float4 pv = reinterpret_cast<const float4*>(p)[t];
    const float4 qv = reinterpret_cast<const float4*>(q)[t];
    xv.x = xv.x + alpha * pv.x; rv.x = rv.x - alpha * qv.x;
    xv.y = xv.y + alpha * pv.y; rv.y = rv.y - alpha * qv.y;
    xv.z = xv.z + alpha * pv.z; rv.z = rv.z - alpha * qv.z;
    xv.w = xv.w + alpha * pv.w; rv.w = rv.w - alpha * qv.w;
    reinterpret_cast<float4*>(x)[t] = xv;
    reinterpret_cast<float4*>(r)[t] = rv;
    const uint32_t c0 = c4 & 0xff, c1 = (c4 >> 8) & 0xff, c2 = (c4 >> 16) & 0xff, c3 = c4 >> 24;
    const float z0 = c0 ? coef.invdiag[c0 - 1] * rv.x : 0.0f;
    const float z1 = c1 ? coef.invdiag[c1 - 1] * rv.y : 0.0f;
    const float z2 = c2 ? coef.invdiag[c2 - 1] * rv.z : 0.0f;
    const float z3 = c3 ? coef.invdiag[c3 - 1] * rv.w : 0.0f;
    acc_r2 += (double)rv.x * rv.x + (double)rv.y * rv.y + (double)rv.z * rv.z + (double)rv.w * rv.w;
    acc_rz += (double)rv.x * z0 + (double)rv.y * z1 + (double)rv.z * z2 + (double)rv.w * z3;
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

int fsb_k_pressure_solve(fsb_ctx* c, float density, float dt)
{
  const GridDims d{c->nx, c->ny, c->ld, c->dx, c->dy};
  CgCoef coef;
  const double dx2 = std::pow((double)c->dx, 2);
  coef.off = (float)(1 / dx2);
  for (int n = 0; n < 5; ++n)
  {
    coef.diag[n] = (float)(-n / dx2);
    coef.invdiag[n] = (coef.diag[n] != 0.0f) ? 1.0f / coef.diag[n] : 1.0f;
  }

  // ---- build
  fsb_prof_begin(c, FSB_PROF_RHS);
  const int64_t total = (int64_t)c->ld * c->ny;
  const int build_blocks = (int)std::min<int64_t>(fsb_div_up(total, 256), c->sm_count * 8);
  const dim3 spmv_grid(fsb_div_up(c->ld, kSpmvThreads * 4), fsb_div_up(c->ny, kRows));
  const int64_t n4 = total / 4;
  const int upd_blocks = (int)std::min<int64_t>(fsb_div_up(n4, kUpdThreads), c->sm_count * 8);
  const int need = std::max(std::max(3 * build_blocks, (int)(spmv_grid.x * spmv_grid.y)),
                            2 * upd_blocks);
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
  FSB_CUDA(c, cudaMemcpyAsync(c->scal_h, c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                              c->stream));
  fsb_prof_end(c, FSB_PROF_RHS);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->scal_h->n_liquid == 0) return FSB_OK; // :347-350: nothing touched, no swap

  // ---- iterate
  fsb_prof_begin(c, FSB_PROF_CG);
  int cur = 0; // cg_p[cur] holds the previous direction
  while (!c->scal_h->done)
  {
    for (int k = 0; k < kCheckEvery; ++k)
    {
      k_cg_dir_spmv<<<spmv_grid, kSpmvThreads, 0, c->stream>>>(
          c->cg_p[cur], c->cg_p[cur ^ 1], c->cg_q, c->cg_r, c->cg_code, c->ld, c->ny, coef, c->scal,
          c->partials);
      FSB_LAUNCHED(c);
      k_cg_update<<<upd_blocks, kUpdThreads, 0, c->stream>>>(c->cg_x, c->cg_r, c->cg_p[cur ^ 1],
                                                             c->cg_q, c->cg_code, n4, coef, c->scal,
                                                             c->partials);
      FSB_LAUNCHED(c);
      cur ^= 1;
    }
    FSB_CUDA(c, cudaMemcpyAsync(c->scal_h, c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                                c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  fsb_prof_end(c, FSB_PROF_CG);
  c->iters = c->scal_h->iter;
  c->err = (c->scal_h->rhs2 == 0.0 || (float)c->scal_h->rhs2 == 0.0f)
               ? 0.0f
               : std::sqrt((float)c->scal_h->r2 / (float)c->scal_h->rhs2);
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
