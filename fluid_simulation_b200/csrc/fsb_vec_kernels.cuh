// Vectorised stage kernels: four cells per thread, 16-byte accesses on the fp32 grids, labels as
// bit masks (LabWin, fsb_device.cuh).  None of them uses shared memory, shuffles or barriers, so
// the same source is also compiled for the host by tests/cpu_emul (one loop iteration per
// thread) and checked against the oracle without a GPU.  Product code: includes nothing of oracle/.
// The including translation unit selects the groups it launches: FSB_VEC_WANT_GRID (labels,
// walls, extension) and / or FSB_VEC_WANT_CG (pressure set-up group, velocity patch).
#pragma once

#include "fsb.h"
#include "fsb_device.cuh"

namespace {

#ifdef FSB_VEC_WANT_GRID
// 4-cells-per-thread forms.  One thread per float4 group: 16-byte accesses on the fp32 grids,
// 4-byte accesses on the label rows, labels handled as bit masks (fsb_device.cuh, LabWin).  The
// one-cell-per-thread kernels above launch 65 536 CTAs of byte-wide loads at 4096^2 and run at
// 8-38 % of the DRAM bandwidth (profiles/r01g_stage_kernels_ncu_full.md); these forms are the
// ones the steps use.  Every face's arithmetic is the reference's, so results are bit-identical.
// src/MacGrid.cpp:32-50 + src/FluidDomain.cpp:169-179, 16 labels per thread; pad columns SOLID
__global__ void k_fill_labels16(uint8_t* __restrict__ cell, const GridDims d)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int j = blockIdx.y;
  if (i0 >= d.ld) return;
  uint32_t w[4];
  const bool border_row = (j == 0 || j == d.ny - 1);
#pragma unroll
  for (int q = 0; q < 4; ++q)
  {
    uint32_t word = 0u;
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
      const int i = i0 + 4 * q + t;
      const bool solid = border_row || i == 0 || i >= d.nx - 1;
      word |= (uint32_t)(solid ? FSB_SOLID : FSB_AIR) << (8 * t);
    }
    w[q] = word;
  }
  *reinterpret_cast<uint4*>(cell + i0 + (size_t)j * d.ld) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256)
k_enforce_dirichlet4(float* __restrict__ uf, float* __restrict__ vf,
                     const uint8_t* __restrict__ cell, const GridDims d)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i0 >= d.nx) return;
  const size_t k = i0 + (size_t)j * d.ld;
  const uint32_t sd_c = window_mask(lab_window(cell, d, i0, j), FSB_SOLID);
  const uint32_t sd_s = center_mask(cell, d, i0, j - 1, FSB_SOLID);
  if (((sd_c | (sd_c << 1) | sd_s) & kOwnBits) == 0u) return; // no wall next to these faces
  const float4 u0 = *reinterpret_cast<const float4*>(uf + k);
  const float4 v0 = *reinterpret_cast<const float4*>(vf + k);
  float4 u = u0, v = v0;
  dirichlet4(u, v, sd_c, sd_s);
  if (u.x != u0.x || u.y != u0.y || u.z != u0.z || u.w != u0.w)
    *reinterpret_cast<float4*>(uf + k) = u;
  if (v.x != v0.x || v.y != v0.y || v.z != v0.z || v.w != v0.w)
    *reinterpret_cast<float4*>(vf + k) = v;
}

// ---- velocity extension with two sweeps (the only count any step uses) in two passes -------
// src/FluidSolver.cpp:485-622.  The validity masks of the init pass are pure functions of the
// labels (u face valid <=> LIQUID at (i,j) or (i-1,j); v face: (i,j) or (i,j-1)), so they are
// recomputed as bit masks instead of being stored and re-read:
//   pass A = init + sweep 1: reads front u, v and labels, writes back u, v and ONE packed mask
//            byte per cell (bit 0: u face valid after sweep 1, bit 1: v face);
//   pass B = sweep 2 (in place on the back buffers; it writes only faces whose mask bit is 0 and
//            reads only faces whose bit is 1) + the zeroing of the front u of the init pass
//            (:509,527 -- including the typo that zeroes U for an invalid V face), which has to
//            wait until no thread reads the front buffer any more.
// 18 + 2 B per cell instead of 21 + 2 x 3 B, two launches instead of three.
__device__ __forceinline__ void extend_faces(float4& out, uint32_t& valid_c, uint32_t need,
                                             uint32_t m_s, uint32_t m_c, uint32_t m_n,
                                             const float4& row_c, const float4& row_s,
                                             const float4& row_n, float west, float east)
{
#pragma unroll
  for (int t = 0; t < 4; ++t)
  {
    if (!((need >> (t + 4)) & 1u)) continue;
    float nv = 0.0f;
    int n = 0;
    // the reference's order: (i-1,j), (i,j-1), (i,j+1), (i+1,j)
    if ((m_c >> (t + 3)) & 1u) { nv += (t == 0) ? west : f4_get(row_c, t - 1); n++; }
    if ((m_s >> (t + 4)) & 1u) { nv += f4_get(row_s, t); n++; }
    if ((m_n >> (t + 4)) & 1u) { nv += f4_get(row_n, t); n++; }
    if ((m_c >> (t + 5)) & 1u) { nv += (t == 3) ? east : f4_get(row_c, t + 1); n++; }
    if (n > 0)
    {
      f4_set(out, t, nv / (float)n);
      valid_c |= 1u << (t + 4);
    }
  }
}

constexpr int kExtendRows = 4; // rows per thread: fewer, fatter CTAs and independent loads in flight

__device__ __forceinline__ void extend2_a_group(const float* __restrict__ uf,
                                                const float* __restrict__ vf,
                                                float* __restrict__ ub, float* __restrict__ vb,
                                                uint8_t* __restrict__ m1,
                                                const uint8_t* __restrict__ cell, const GridDims& d,
                                                int i0, int j, const float4 u4, const float4 v4,
                                                const uint32_t lab4)
{
  const size_t k = i0 + (size_t)j * d.ld;
  if (i0 + 4 <= d.nx && lab4 == FSB_LIQUID * 0x01010101u)
  {
    // four LIQUID cells (the bulk of a full tank): every face is valid, the pass is a copy
    *reinterpret_cast<float4*>(ub + k) = u4;
    *reinterpret_cast<float4*>(vb + k) = v4;
    *reinterpret_cast<uint32_t*>(m1 + k) = 0x03030303u;
    return;
  }
  const LabWin w_c = lab_window(cell, d, i0, j), w_s = lab_window(cell, d, i0, j - 1);
  const uint32_t lq_c = window_mask(w_c, FSB_LIQUID), lq_s = window_mask(w_s, FSB_LIQUID);
  uint32_t mu_c = lq_c | (lq_c << 1), mv_c = lq_c | lq_s; // validity after the init pass
  float4 ou = u4, ov = v4;
#pragma unroll
  for (int t = 0; t < 4; ++t)
  {
    if (!((mu_c >> (t + 4)) & 1u)) f4_set(ou, t, 0.0f);
    if (!((mv_c >> (t + 4)) & 1u)) f4_set(ov, t, 0.0f);
  }
  const uint32_t inner = interior_columns(d, i0, j);
  if (inner != 0u && ((~mu_c | ~mv_c) & inner) != 0u)
  {
    const uint32_t sd_c = window_mask(w_c, FSB_SOLID), sd_s = window_mask(w_s, FSB_SOLID);
    const uint32_t need_u = ~mu_c & ~sd_c & ~(sd_c << 1) & inner;
    const uint32_t need_v = ~mv_c & ~sd_c & ~sd_s & inner;
    if ((need_u | need_v) != 0u)
    {
      // interior rows: j-1 >= 0 and j+1 <= ny-1 exist, and so do columns i-1 and i+1 of a face in need
      const uint32_t lq_n = window_mask(lab_window(cell, d, i0, j + 1), FSB_LIQUID);
      if (need_u != 0u)
      {
        const uint32_t mu_s = lq_s | (lq_s << 1), mu_n = lq_n | (lq_n << 1);
        const float4 rs = *reinterpret_cast<const float4*>(uf + k - d.ld);
        const float4 rn = *reinterpret_cast<const float4*>(uf + k + d.ld);
        const float west = ((need_u >> 4) & 1u) ? uf[k - 1] : 0.0f;
        const float east = ((need_u >> 7) & 1u) ? uf[k + 4] : 0.0f;
        extend_faces(ou, mu_c, need_u, mu_s, mu_c, mu_n, u4, rs, rn, west, east);
      }
      if (need_v != 0u)
      {
        const uint32_t lq_s2 = center_mask(cell, d, i0, j - 2, FSB_LIQUID);
        const uint32_t mv_s = lq_s | lq_s2, mv_n = lq_n | lq_c;
        const float4 rs = *reinterpret_cast<const float4*>(vf + k - d.ld);
        const float4 rn = *reinterpret_cast<const float4*>(vf + k + d.ld);
        const float west = ((need_v >> 4) & 1u) ? vf[k - 1] : 0.0f;
        const float east = ((need_v >> 7) & 1u) ? vf[k + 4] : 0.0f;
        extend_faces(ov, mv_c, need_v, mv_s, mv_c, mv_n, v4, rs, rn, west, east);
      }
    }
  }
  *reinterpret_cast<float4*>(ub + k) = ou;
  *reinterpret_cast<float4*>(vb + k) = ov;
  uint32_t packed = 0u;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    packed |= (((mu_c >> (t + 4)) & 1u) | (((mv_c >> (t + 4)) & 1u) << 1)) << (8 * t);
  *reinterpret_cast<uint32_t*>(m1 + k) = packed;
}

__global__ void __launch_bounds__(256)
k_extend2_a(const float* __restrict__ uf, const float* __restrict__ vf, float* __restrict__ ub,
            float* __restrict__ vb, uint8_t* __restrict__ m1, const uint8_t* __restrict__ cell,
            const GridDims d)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= d.nx) return;
  // the loads of all rows first: the rows are independent, but each row's stores would otherwise keep the
  // next row's loads behind them (one DRAM round trip per row instead of one per thread)
  float4 u4[kExtendRows], v4[kExtendRows];
  uint32_t lab4[kExtendRows];
#pragma unroll
  for (int r = 0; r < kExtendRows; ++r)
  {
    const int j = min((int)blockIdx.y * kExtendRows + r, d.ny - 1);
    const size_t k = i0 + (size_t)j * d.ld;
    u4[r] = *reinterpret_cast<const float4*>(uf + k);
    v4[r] = *reinterpret_cast<const float4*>(vf + k);
    lab4[r] = *reinterpret_cast<const uint32_t*>(cell + k);
  }
#pragma unroll
  for (int r = 0; r < kExtendRows; ++r)
  {
    const int j = blockIdx.y * kExtendRows + r;
    if (j < d.ny) extend2_a_group(uf, vf, ub, vb, m1, cell, d, i0, j, u4[r], v4[r], lab4[r]);
  }
}

__device__ __forceinline__ void extend2_b_group(float* __restrict__ uf, float* __restrict__ ub,
                                                float* __restrict__ vb,
                                                const uint8_t* __restrict__ m1,
                                                const uint8_t* __restrict__ cell, const GridDims& d,
                                                int i0, int j)
{
  const size_t k = i0 + (size_t)j * d.ld;
  // four LIQUID cells: nothing to zero and every face was valid from the start
  if (i0 + 4 <= d.nx && *reinterpret_cast<const uint32_t*>(cell + k) == FSB_LIQUID * 0x01010101u)
    return;
  const LabWin w_c = lab_window(cell, d, i0, j), w_s = lab_window(cell, d, i0, j - 1);
  const uint32_t lq_c = window_mask(w_c, FSB_LIQUID), lq_s = window_mask(w_s, FSB_LIQUID);
  // init pass, front buffer: U is zeroed where the u face OR the v face is invalid (:509,527)
  const uint32_t zero = ~((lq_c | (lq_c << 1)) & (lq_c | lq_s)) & own_columns(d, i0);
  if (zero == kOwnBits) *reinterpret_cast<float4*>(uf + k) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  else if (zero != 0u)
  {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((zero >> (t + 4)) & 1u) uf[k + t] = 0.0f;
  }
  const uint32_t inner = interior_columns(d, i0, j);
  if (inner == 0u) return;
  const LabWin q_c = byte_window(m1 + (size_t)j * d.ld, d.ld, i0);
  const uint32_t mu_c = window_bit(q_c, 0), mv_c = window_bit(q_c, 1);
  if (((~mu_c | ~mv_c) & inner) == 0u) return;
  const uint32_t sd_c = window_mask(w_c, FSB_SOLID), sd_s = window_mask(w_s, FSB_SOLID);
  const uint32_t need_u = ~mu_c & ~sd_c & ~(sd_c << 1) & inner;
  const uint32_t need_v = ~mv_c & ~sd_c & ~sd_s & inner;
  if ((need_u | need_v) == 0u) return;
  const LabWin q_s = byte_window(m1 + (size_t)(j - 1) * d.ld, d.ld, i0);
  const LabWin q_n = byte_window(m1 + (size_t)(j + 1) * d.ld, d.ld, i0);
  if (need_u != 0u)
  {
    const float4 rc = *reinterpret_cast<const float4*>(ub + k);
    const float4 rs = *reinterpret_cast<const float4*>(ub + k - d.ld);
    const float4 rn = *reinterpret_cast<const float4*>(ub + k + d.ld);
    const float west = ((need_u >> 4) & 1u) ? ub[k - 1] : 0.0f;
    const float east = ((need_u >> 7) & 1u) ? ub[k + 4] : 0.0f;
    float4 out = rc;
    uint32_t valid = mu_c;
    extend_faces(out, valid, need_u, window_bit(q_s, 0), mu_c, window_bit(q_n, 0), rc, rs, rn, west,
                 east);
    const uint32_t fresh = valid & ~mu_c;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((fresh >> (t + 4)) & 1u) ub[k + t] = f4_get(out, t);
  }
  if (need_v != 0u)
  {
    const float4 rc = *reinterpret_cast<const float4*>(vb + k);
    const float4 rs = *reinterpret_cast<const float4*>(vb + k - d.ld);
    const float4 rn = *reinterpret_cast<const float4*>(vb + k + d.ld);
    const float west = ((need_v >> 4) & 1u) ? vb[k - 1] : 0.0f;
    const float east = ((need_v >> 7) & 1u) ? vb[k + 4] : 0.0f;
    float4 out = rc;
    uint32_t valid = mv_c;
    extend_faces(out, valid, need_v, window_bit(q_s, 1), mv_c, window_bit(q_n, 1), rc, rs, rn, west,
                 east);
    const uint32_t fresh = valid & ~mv_c;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if ((fresh >> (t + 4)) & 1u) vb[k + t] = f4_get(out, t);
  }
}


__global__ void __launch_bounds__(256)
k_extend2_b(float* __restrict__ uf, float* __restrict__ ub, float* __restrict__ vb,
            const uint8_t* __restrict__ m1, const uint8_t* __restrict__ cell, const GridDims d)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= d.nx) return;
  // Row groups from the top of the grid down: the rows with work in this pass (free surface, AIR above it)
  // are the upper ones in a tank, and their CTAs -- dependent label / mask / value loads, four rows in turn
  // -- are the long ones; started first they overlap the all-LIQUID rows instead of forming the tail.
  const int jy = (int)gridDim.y - 1 - (int)blockIdx.y;
  // the label words of all rows first (independent loads); an all-LIQUID group has nothing to do
  uint32_t lab4[kExtendRows];
#pragma unroll
  for (int r = 0; r < kExtendRows; ++r)
  {
    const int j = min(jy * kExtendRows + r, d.ny - 1);
    lab4[r] = *reinterpret_cast<const uint32_t*>(cell + i0 + (size_t)j * d.ld);
  }
#pragma unroll
  for (int r = 0; r < kExtendRows; ++r)
  {
    const int j = jy * kExtendRows + r;
    if (j >= d.ny) continue;
    if (i0 + 4 <= d.nx && lab4[r] == FSB_LIQUID * 0x01010101u) continue;
    extend2_b_group(uf, ub, vb, m1, cell, d, i0, j);
  }
}

#endif // FSB_VEC_WANT_GRID

#ifdef FSB_VEC_WANT_CG
// Four faces of each kind per thread; DIRICHLET folds the enforceDirichlet that follows the
// projection in every step (src/FluidSolver.cpp:297-321 after :482's swap) into the same pass:
// it acts on the patched value where the face was patched and on the stale back-buffer value
// elsewhere, exactly what the separate pass sees after the swap.  LIQUID <=> code != 0 (the
// labels have not changed since k_cg_build), so the labels are the only mask read.
template <bool DIRICHLET, class D>
__global__ void __launch_bounds__(256)
k_pressure_patch4(const float* __restrict__ uf, const float* __restrict__ vf, float* __restrict__ ub,
                  float* __restrict__ vb, const float* __restrict__ x,
                  const uint8_t* __restrict__ cell, const D d, float dt, float density)
{
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i0 >= d.nx) return;
  const size_t k = i0 + (size_t)j * d.ld;
  const LabWin w_c = lab_window(cell, d, i0, j);
  const uint32_t lq_c = window_mask(w_c, FSB_LIQUID), lq_s = center_mask(cell, d, i0, j - 1, FSB_LIQUID);
  const uint32_t own = own_columns(d, i0);
  const uint32_t touched = (lq_c | (lq_c << 1) | lq_s) & own; // :438
  const uint32_t sd_c = DIRICHLET ? window_mask(w_c, FSB_SOLID) : 0u;
  const uint32_t sd_s = DIRICHLET ? center_mask(cell, d, i0, j - 1, FSB_SOLID) : 0u;
  const uint32_t walls = (sd_c | (sd_c << 1) | sd_s) & own; // faces a wall condition can change
  if (touched == 0u && walls == 0u) return;
  float4 ou = make_float4(0.0f, 0.0f, 0.0f, 0.0f), ov = ou;
  if (touched != kOwnBits)
  {
    ou = *reinterpret_cast<const float4*>(ub + k); // stale values stay
    ov = *reinterpret_cast<const float4*>(vb + k);
  }
  if (touched != 0u)
  {
    const float4 u4 = *reinterpret_cast<const float4*>(uf + k);
    const float4 v4 = *reinterpret_cast<const float4*>(vf + k);
    const float4 xc = *reinterpret_cast<const float4*>(x + k);
    // index-clamped neighbours (:432-433): row -1 is row 0, column -1 is column 0
    const float4 xs = (j > 0) ? *reinterpret_cast<const float4*>(x + k - d.ld) : xc;
    const float xw = (i0 > 0) ? x[k - 1] : xc.x;
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
      if (!((touched >> (t + 4)) & 1u)) continue;
      const bool l = (lq_c >> (t + 4)) & 1u, lw = (lq_c >> (t + 3)) & 1u, ls = (lq_s >> (t + 4)) & 1u;
      // the particle-pressure terms are k * n with k = 0.0 (:443-453): exactly +0
      const float pc = l ? f4_get(xc, t) + 0.0f : 0.0f;
      const float pw = lw ? ((t == 0) ? xw : f4_get(xc, t - 1)) + 0.0f : 0.0f;
      const float ps = ls ? f4_get(xs, t) + 0.0f : 0.0f;
      const float ddx = pc - pw;
      const float ddy = pc - ps;
      // x / delta through div_dx / div_dy: one multiplication when delta is a power of two
      f4_set(ou, t, f4_get(u4, t) - div_dx(d, (dt / density) * ddx));
      f4_set(ov, t, f4_get(v4, t) - div_dy(d, (dt / density) * ddy));
    }
  }
  if (DIRICHLET && walls != 0u) dirichlet4(ou, ov, sd_c, sd_s);
  *reinterpret_cast<float4*>(ub + k) = ou;
  *reinterpret_cast<float4*>(vb + k) = ov;
}


// One float4 group of the pressure system set-up (src/FluidSolver.cpp:329-346,368-416; the
// arithmetic of k_cg_build): stencil codes, b = divergence on LIQUID cells, and the group's
// contributions to |b|^2, b.z and the liquid count.  `invdiag` = Eigen's Jacobi preconditioner.
template <class D>
__device__ __forceinline__ void cg_build_group(const float* __restrict__ uf,
                                               const float* __restrict__ vf,
                                               const uint8_t* __restrict__ cell, const D& d,
                                               const float* invdiag, int i0, int j, uint32_t* code4,
                                               float4* b4, double* acc_b2, double* acc_bz,
                                               double* acc_n)
{
  const size_t t0 = i0 + (size_t)j * d.ld;
  uint32_t cd = 0u;
  float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (i0 < d.nx)
  {
    const LabWin w_c = lab_window(cell, d, i0, j);
    const uint32_t liq = window_mask(w_c, FSB_LIQUID) & own_columns(d, i0);
    if (liq != 0u)
    {
      const uint32_t sd_c = window_mask(w_c, FSB_SOLID);
      const uint32_t sd_s = center_mask(cell, d, i0, j - 1, FSB_SOLID);
      const uint32_t sd_n = center_mask(cell, d, i0, j + 1, FSB_SOLID);
      const float4 u4 = *reinterpret_cast<const float4*>(uf + t0);
      const float4 v4 = *reinterpret_cast<const float4*>(vf + t0);
      // a LIQUID cell is never on the border in a classified grid; clamp for safety
      const int jn = min(j + 1, d.ny - 1);
      const float4 vn = *reinterpret_cast<const float4*>(vf + i0 + (size_t)jn * d.ld);
      const float ue = uf[min(i0 + 4, d.nx - 1) + (size_t)j * d.ld];
      // bulk of the liquid: four LIQUID cells without a SOLID neighbour
      const bool bulk = liq == kOwnBits && ((sd_c & 0x1f8u) | sd_s | sd_n) == 0u;
#pragma unroll
      for (int t = 0; t < 4; ++t)
      {
        if (!((liq >> (t + 4)) & 1u)) continue;
        const int n = bulk ? 4
                           : (int)(((~sd_c >> (t + 3)) & 1u) + ((~sd_c >> (t + 5)) & 1u) +
                                   ((~sd_s >> (t + 4)) & 1u) + ((~sd_n >> (t + 4)) & 1u));
        cd |= (uint32_t)(1 + n) << (8 * t);
        const int ie = min(i0 + t + 1, d.nx - 1) - i0; // east face: inside the group, after it, or clamped
        const float u_e =
            (ie >= 4) ? ue : ((ie == t + 1 && t < 3) ? f4_get(u4, t + 1) : f4_get(u4, t));
        const float bb = div_dx(d, u_e - f4_get(u4, t)) + div_dy(d, f4_get(vn, t) - f4_get(v4, t));
        f4_set(b, t, bb);
        const float z = invdiag[n] * bb;
        *acc_b2 += (double)bb * (double)bb;
        *acc_bz += (double)bb * (double)z;
        *acc_n += 1.0;
      }
    }
  }
  *code4 = cd;
  *b4 = b;
}

#endif // FSB_VEC_WANT_CG

} // namespace
