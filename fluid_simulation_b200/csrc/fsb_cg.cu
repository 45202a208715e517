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
// One CG iteration = two sweeps (Eigen's statement order is kept):
//   direction : p = z + beta p (z = invdiag r; p = z on the first pass) on a tile and its halo,
//               q = A p in registers, partial p.q        -> 9 B read + 4 B written per cell
//   update    : alpha = absNew / p.q; q = A p recomputed from the staged p tile (q is never
//               stored); r -= alpha q; partial |r|^2, r.z; x += alpha p -- in the persistent
//               kernel only on odd iterations, two updates back to back (bit-identical x)
//                                                        -> 9 + 4 B (+ 12 B every other iteration)
// 32 B of HBM traffic per cell and iteration against the 45 B of the textbook formulation
// (SURVEY.md 8d).  Launch modes: k_cg_solve, one persistent cooperative kernel for the whole solve
// with a software grid barrier that carries the reductions (default), or k_cg_direction +
// k_cg_update per iteration in a CUDA graph of kCheckEvery iterations with the host polling `done`
// one chunk behind the GPU.  Partials are folded in a fixed order (deterministic); once `done` is
// set nothing is modified any more, so x and the iteration count are exactly those of the
// converging iteration.  DESIGN.md section 5.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"
#define FSB_VEC_WANT_CG
#include "fsb_vec_kernels.cuh"
#include "fsb_cg_frame.cuh"

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
// first dot products -> device scalars (thread 0 of the last block of the build kernel)
__device__ __forceinline__ void cg_init_scalars(CgScalars* __restrict__ s, double tb2, double tbz,
                                                double tn, float tol, int max_iters)
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
  s->bar_count = 0;
  s->bar_release = 0;
  s->tile_list = nullptr; // set by k_cg_tile_compact when tile skipping is on
  s->n_active_tiles = 0;
  s->n_prefix_tiles = 0;
}

// src/FluidSolver.cpp:329-346,368-416: stencil code, right-hand side
// b = divVelX + divVelY (include/MacGrid.h:98-111) on LIQUID cells, x = 0,
// r = b, and the first dot products (|b|^2, b.z).
__global__ void k_cg_build(const float* __restrict__ uf, const float* __restrict__ vf,
                           const uint8_t* __restrict__ cell, uint8_t* __restrict__ code,
                           float* __restrict__ x, float* __restrict__ r, const GridDims d,
                           const CgCoef coef, CgScalars* __restrict__ s,
                           double* __restrict__ partials, float tol, int max_iters)
{
  // blocks walk (row, 256-column segment) pairs: no per-cell integer division
  const int segs = (d.ld + 255) / 256;
  const int n_work = segs * d.ny;
  double acc_b2 = 0.0, acc_bz = 0.0, acc_n = 0.0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x)
  {
    const int j = w / segs;
    const int i = (w - j * segs) * 256 + threadIdx.x;
    if (i >= d.ld) continue;
    const size_t t = i + (size_t)j * d.ld;
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
    if (threadIdx.x == 0) cg_init_scalars(s, tb2, tbz, tn, tol, max_iters);
  }
}


// The same set-up with four cells per thread (cg_build_group, fsb_vec_kernels.cuh)
template <class D>
__global__ void __launch_bounds__(256)
k_cg_build4(const float* __restrict__ uf, const float* __restrict__ vf,
            const uint8_t* __restrict__ cell, uint8_t* __restrict__ code, float* __restrict__ x,
            float* __restrict__ r, const D d, const CgCoef coef, CgScalars* __restrict__ s,
            double* __restrict__ partials, float tol, int max_iters)
{
  const int segs = (d.ld + 1023) / 1024;
  const int n_work = segs * d.ny;
  double acc_b2 = 0.0, acc_bz = 0.0, acc_n = 0.0;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x)
  {
    const int j = w / segs;
    const int i0 = ((w - j * segs) * 256 + threadIdx.x) * 4;
    if (i0 >= d.ld) continue;
    const size_t t0 = i0 + (size_t)j * d.ld;
    uint32_t cd;
    float4 b;
    cg_build_group(uf, vf, cell, d, coef.invdiag, i0, j, &cd, &b, &acc_b2, &acc_bz, &acc_n);
    *reinterpret_cast<uint32_t*>(code + t0) = cd;
    *reinterpret_cast<float4*>(x + t0) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    *reinterpret_cast<float4*>(r + t0) = b;
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
    if (threadIdx.x == 0) cg_init_scalars(s, tb2, tbz, tn, tol, max_iters);
  }
}


template <int TH>
struct DirStage // k_cg_direction: r, p_old (both with halo), code (with halo)
{
  static constexpr int kF32 = kHaloW * (TH + 2) * 4;
  static constexpr int kCode = kCodeW * (TH + 2);
  static constexpr int oR = 0, oP = align128(kF32), oC = 2 * align128(kF32);
  static constexpr int kBytes = align128(oC + kCode);
  static constexpr int kTx = 2 * kF32 + kCode;
};
template <int TH>
struct UpdStage // k_cg_update: p (with halo), x, r (interior), code (with halo)
{
  static constexpr int kHalo = kHaloW * (TH + 2) * 4;
  static constexpr int kInner = kTileW * TH * 4;
  static constexpr int kCode = kCodeW * (TH + 2);
  static constexpr int oP = 0, oX = align128(kHalo), oR = oX + kInner, oC = oR + kInner;
  static constexpr int kBytes = align128(oC + kCode);
  static constexpr int kTx = kHalo + 2 * kInner + kCode;
};

template <int TH>
struct UpdStageX : UpdStage<TH> // persistent solve: + the previous direction (deferred x update)
{
  static constexpr int oQ = UpdStage<TH>::kBytes;
  static constexpr int kBytes = align128(oQ + UpdStage<TH>::kInner);
};

struct CgMaps
{
  CUtensorMap halo_a; // fp32, box kHaloW x (TH+2)
  CUtensorMap halo_b; // fp32, box kHaloW x (TH+2)
  CUtensorMap inner_a; // fp32, box kTileW x TH
  CUtensorMap inner_b; // fp32, box kTileW x TH
  CUtensorMap code;    // u8,   box kCodeW x (TH+2)
};




// ---- active-tile list: tiles whose interior holds at least one LIQUID cell (code != 0).  All CG
// vectors are exactly zero on the other tiles and stay zero, so the sweeps skip them: in a tank
// that is the air cap, in a dam-break scene most of the grid.
__global__ void __launch_bounds__(128)
k_cg_tile_flags(const uint8_t* __restrict__ code, int ld, int tiles_x, int th, int row_lo, int row_hi,
                int* __restrict__ flags, int keep_edge_rows)
{
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  // sharded solves: the tiles of the slab's first and last tile row stay active whatever they
  // hold, so that their rows (zeros included) are stored into the neighbours' ghost rows in
  // every sweep -- the ghost rows are the peers' to write, this rank never clears them
  // (the one-sweep solve sends TWO rows of p per side: every tile holding one of the slab's last two rows)
  if (keep_edge_rows && (ty == 0 || row_lo + (ty + 1) * th >= row_hi - 1))
  {
    if (threadIdx.x == 0) flags[blockIdx.x] = 1;
    return;
  }
  const int col = tx * kTileW + (int)(threadIdx.x & 7) * 16;
  int any = 0;
  if (col < ld)
    for (int r = threadIdx.x >> 3; r < th; r += 16)
    {
      const int row = row_lo + ty * th + r;
      if (row < row_hi)
      {
        const uint4 v = *reinterpret_cast<const uint4*>(code + col + (size_t)row * ld);
        any |= (v.x | v.y | v.z | v.w) != 0u;
      }
    }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[blockIdx.x] = any ? 1 : 0;
}

// ordered compaction by one CTA (the list order fixes which CTA sums which tile: deterministic).
// boundary_first: the active tiles of the slab's first and last tile row go to the front of the
// list (TileWalk's prefix), then all other active tiles in tile order.
__global__ void __launch_bounds__(1024)
k_cg_tile_compact(const int* __restrict__ flags, int n_tiles, int tiles_x, int* __restrict__ list,
                  CgScalars* __restrict__ s, int boundary_first)
{
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int last_ty = n_tiles / tiles_x - 1;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  int n_prefix = 0;
  for (int pass = boundary_first ? 0 : 1; pass < 2; ++pass)
  {
    for (int c0 = 0; c0 < n_tiles; c0 += 4096)
    {
      // four consecutive tiles per thread: a third of the rounds (each costs four CTA barriers)
      const int t0 = c0 + (int)threadIdx.x * 4;
      int fq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
      {
        const int t = t0 + q;
        fq[q] = (t < n_tiles) ? (flags[t] != 0) : 0;
        if (boundary_first && t < n_tiles)
        {
          const int ty = t / tiles_x;
          const bool edge = (ty == 0 || ty == last_ty);
          if (edge != (pass == 0)) fq[q] = 0;
        }
      }
      const int f = fq[0] + fq[1] + fq[2] + fq[3];
      int incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[wid] = incl;
      __syncthreads();
      if (wid == 0)
      {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const int v = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w; // exclusive offset of warp `lane`
      }
      __syncthreads();
      const int base = s_base;
      int ex = base + s_warp[wid] + incl - f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (fq[q])
        {
          const int t = t0 + q;
          list[ex++] = ((t / tiles_x) << 16) | (t % tiles_x);
        }
      __syncthreads();
      if (threadIdx.x == 1023) s_base = ex;
      __syncthreads();
    }
    if (pass == 0) n_prefix = s_base;
  }
  if (threadIdx.x == 0)
  {
    s->n_active_tiles = s_base;
    s->n_prefix_tiles = n_prefix;
    s->tile_list = list;
  }
}


// ---- scalar updates shared by the single-GPU last-CTA path and the combine kernels
__device__ __forceinline__ void finalize_update(CgScalars* s, double tr2, double trz)
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
}


// Called by ONE thread after the values are final (and after a system-scope fence if peer rows
// were stored by this kernel): write them into every rank's mailbox.
__device__ __forceinline__ void mail_post(const ShardArgs& sh, int type, double v0, double v1,
                                          unsigned long long seq)
{
  const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0);
  const unsigned long long b1 = (unsigned long long)__double_as_longlong(v1);
  for (int q = 0; q < sh.world; ++q)
  {
    volatile unsigned long long* w =
        reinterpret_cast<volatile unsigned long long*>(sh.mail[q] + type * kMaxRanks + sh.rank);
    w[0] = mail_word((unsigned int)b0, seq);
    w[1] = mail_word((unsigned int)(b0 >> 32), seq);
    w[2] = mail_word((unsigned int)b1, seq);
    w[3] = mail_word((unsigned int)(b1 >> 32), seq);
  }
}

__device__ __forceinline__ unsigned long long global_ns();

// Called by ONE thread right after mail_post: wait until every rank's entry of `type` carries
// `seq`, then add the entries in rank order (the same order on every rank -> bit-identical
// totals).  Returns false after kMailTimeoutNs without an answer.
__device__ __forceinline__ bool mail_collect(const ShardArgs& sh, int type, unsigned long long seq,
                                             double* v0, double* v1);


__device__ __forceinline__ bool mail_collect(const ShardArgs& sh, int type, unsigned long long seq,
                                             double* v0, double* v1)
{
  const unsigned long long tag = seq & 0xffffffffull;
  const unsigned long long t0 = global_ns();
  double a = 0.0, b = 0.0;
  for (int q = 0; q < sh.world; ++q) // rank order: identical totals on every rank
  {
    volatile unsigned long long* w =
        reinterpret_cast<volatile unsigned long long*>(sh.mail[sh.rank] + type * kMaxRanks + q);
    unsigned long long w0, w1, w2, w3;
    for (;;)
    {
      w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
      if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
      if (global_ns() - t0 > kMailTimeoutNs) return false;
    }
    a += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    b += __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  }
  __threadfence_system(); // acquire: the peers' boundary rows were stored before their entries
  *v0 = a;
  *v1 = b;
  return true;
}

// The same exchange done by a whole warp (all 32 lanes call it with v0, v1 already broadcast):
// lane q posts into rank q's mailbox and polls this rank's slot q, so the NVLink stores and the
// polls of all peers are in flight together; the entries are then added in rank order through
// shuffles, which keeps the totals bit-identical on every rank.
__device__ __forceinline__ bool mail_allreduce_warp(const ShardArgs& sh, int type,
                                                    unsigned long long seq, double* v0, double* v1)
{
  const int lane = threadIdx.x & 31;
  const unsigned long long tag = seq & 0xffffffffull;
  double a = 0.0, b = 0.0;
  bool ok = true;
  if (lane < sh.world)
  {
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(*v0);
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(*v1);
    volatile unsigned long long* out =
        reinterpret_cast<volatile unsigned long long*>(sh.mail[lane] + type * kMaxRanks + sh.rank);
    out[0] = mail_word((unsigned int)b0, seq);
    out[1] = mail_word((unsigned int)(b0 >> 32), seq);
    out[2] = mail_word((unsigned int)b1, seq);
    out[3] = mail_word((unsigned int)(b1 >> 32), seq);
    volatile unsigned long long* w =
        reinterpret_cast<volatile unsigned long long*>(sh.mail[sh.rank] + type * kMaxRanks + lane);
    const unsigned long long t0 = global_ns();
    unsigned long long w0, w1, w2, w3;
    for (;;)
    {
      w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
      if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
      if (global_ns() - t0 > kMailTimeoutNs)
      {
        ok = false;
        break;
      }
    }
    a = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    b = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  }
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system(); // acquire: the peers' boundary rows were stored before their entries
  double s0 = 0.0, s1 = 0.0;
  for (int q = 0; q < sh.world; ++q)
  {
    s0 += __shfl_sync(0xffffffffu, a, q);
    s1 += __shfl_sync(0xffffffffu, b, q);
  }
  *v0 = s0;
  *v1 = s1;
  return ok;
}

// end-of-solve barrier (TYPE 2): one thread waits for every rank's entry
template <int TYPE>
__global__ void k_cg_combine(const ShardArgs sh, CgScalars* __restrict__ s)
{
  if (threadIdx.x != 0) return;
  const unsigned long long want = s->seq[TYPE] + 1;
  double v0, v1;
  if (!mail_collect(sh, TYPE, want, &v0, &v1))
  {
    s->comm_error = 1;
    s->done = 1;
    s->comm_diag[0] = 1000000000ull + TYPE; // a kernel-boundary wait (TYPE 2: the end-of-solve barrier)
    s->comm_diag[1] = 0;
    s->comm_diag[2] = want;
    s->comm_diag[3] = 0;
  }
  s->seq[TYPE] = want;
}

// end of a sharded solve: every rank stores its rows of x into every peer's copy
__global__ void k_shard_scatter_rows(const float* __restrict__ src, float* const* __restrict__ dst,
                                     int n_dst, int64_t first, int64_t count4)
{
  const float4* s4 = reinterpret_cast<const float4*>(src + first);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count4;
       t += (int64_t)gridDim.x * blockDim.x)
  {
    const float4 v = s4[t];
    for (int q = 0; q < n_dst; ++q) reinterpret_cast<float4*>(dst[q] + first)[t] = v;
  }
}

__global__ void k_shard_post_barrier(const ShardArgs sh, CgScalars* __restrict__ s)
{
  if (threadIdx.x == 0)
  {
    __threadfence_system();
    mail_post(sh, 2, 0.0, 0.0, s->seq[2] + 1);
  }
}

// block-level sum of NW consumer warps' doubles (N values each), then the
// grid-level fold by the last CTA to finish; returns true in the threads of
// that last CTA, with the totals in out[] (valid in thread 0).
template <int NW, int N>
__device__ __forceinline__ bool fold_consumers(double (&acc)[N], unsigned int* ticket,
                                               double* __restrict__ partials, double (&out)[N],
                                               bool pushed)
{
  __shared__ double s_part[N][32];
  __shared__ bool s_last;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < N; ++n)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_down_sync(0xffffffffu, acc[n], o);
    if (lane == 0) s_part[n][warp] = acc[n];
  }
  consumer_sync(NW * 32);
  if (warp == 0)
  {
#pragma unroll
    for (int n = 0; n < N; ++n)
    {
      double v = (lane < NW) ? s_part[n][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) partials[n * gridDim.x + blockIdx.x] = v;
    }
    if (lane == 0)
    {
      // a CTA that stored slab boundary rows into peer memory makes them visible at system
      // scope before it checks in; the mailbox entry follows the last CTA's fold
      if (pushed) __threadfence_system();
      else __threadfence();
      s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
  }
  consumer_sync(NW * 32);
  if (!s_last) return false;
  __threadfence();
  const volatile double* part = partials;
#pragma unroll
  for (int n = 0; n < N; ++n)
  {
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += NW * 32) v += part[n * gridDim.x + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[n][warp] = v;
  }
  consumer_sync(NW * 32);
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int n = 0; n < N; ++n)
    {
      double t = 0.0;
      for (int k = 0; k < NW; ++k) t += s_part[n][k];
      out[n] = t;
    }
    *ticket = 0;
  }
  return true;
}

// ------------------------------------------------- direction + p.Ap dot --
// p_new = z + beta p_old on the warp's RPW rows and on their one-cell halo
// (re-computed from the staged r, p_old, code: no other warp's or CTA's output
// is needed), q = A p_new from registers and shuffles, kept for the dot product
// only.  HBM traffic 13 B per cell: r, p_old, code in (TMA), p_new out
// (STG.128; ping-pong with p_old because other CTAs still read the old halo).
template <int NW, int RPW>
__global__ void __launch_bounds__((NW + 1) * 32)
k_cg_direction(const __grid_constant__ CgMaps maps, float* __restrict__ p_new, int ld,
               int tiles_x, int n_tiles, int stages, const CgCoef coef,
               CgScalars* __restrict__ s, double* __restrict__ partials,
               const __grid_constant__ ShardArgs sh, float* __restrict__ push_lo,
               float* __restrict__ push_hi, int reverse)
{
  constexpr int TH = NW * RPW;
  using St = DirStage<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  __shared__ float4 lut[8];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // set-up that does not depend on the previous kernel overlaps its tail (PDL)
  load_lut(lut, coef);
  if (threadIdx.x == 0)
  {
    for (int k = 0; k < stages; ++k)
    {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], NW);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  if (s->done) return;
  const bool first = (s->iter == 0);
  const float beta = first ? 0.0f : s->beta;
  const int* __restrict__ tile_list = s->tile_list;
  const int n_walk = tile_list ? s->n_active_tiles : n_tiles;
  const int n_prefix = tile_list ? s->n_prefix_tiles : 0;

  if (warp == NW)
  {
    // ---- producer
    if (lane == 0)
    {
      TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, reverse != 0, tile_list, n_prefix);
      int st = 0, round = 0;
      for (int tk = 0; tk < t.count; ++tk, t.next())
      {
        if (round > 0) mbar_wait(&empty[st], (round - 1) & 1);
        const int c0 = t.tx * kTileW, j0 = sh.row_lo + t.ty * TH;
        unsigned char* base = smem + st * St::kBytes;
        mbar_expect_tx(&full[st], first ? St::kTx - St::kF32 : St::kTx);
        tma_load_2d(base + St::oR, &maps.halo_a, c0 - 4, j0 - 1, &full[st]);
        if (!first) tma_load_2d(base + St::oP, &maps.halo_b, c0 - 4, j0 - 1, &full[st]);
        tma_load_2d(base + St::oC, &maps.code, c0 - 16, j0 - 1, &full[st]);
        if (++st == stages) { st = 0; ++round; }
      }
    }
    return;
  }

  // ---- consumers
  const float inv5 = coef.invdiag[4], diag5 = coef.diag[4], off = coef.off;
  // stage-relative offsets of this lane: fp32 element (row r0 of the halo box), code byte
  const int r0 = (int)warp * RPW; // first own row; halo-box rows r0 .. r0+RPW+1
  const int fo = r0 * kHaloW + 4 + (int)lane * 4;
  const int co = r0 * kCodeW + 16 + (int)lane * 4;
  // lanes 0 / 31 also own the west / east halo cell of every row
  const int hfo = r0 * kHaloW + (lane == 31 ? 4 + kTileW : 3);
  const int hco = r0 * kCodeW + (lane == 31 ? 16 + kTileW : 15);
  const bool edge = (lane == 0 || lane == 31);
  double acc[1] = {0.0};
  TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, reverse != 0, tile_list, n_prefix);
  int st = 0, round = 0;
  bool pushed = false; // CTA-uniform: one of this CTA's tiles holds a slab boundary row with a peer
  const int last_ty = n_tiles / tiles_x - 1;
  for (int tk = 0; tk < t.count; ++tk, t.next())
  {
    const unsigned char* base = smem + st * St::kBytes;
    const float* sr = reinterpret_cast<const float*>(base + St::oR);
    const float* sp = reinterpret_cast<const float*>(base + St::oP);
    const unsigned char* sc = base + St::oC;
    pushed |= (push_lo && t.ty == 0) || (push_hi && t.ty == last_ty);
    mbar_wait(&full[st], round & 1);

    float4 pn[RPW + 2];
    uint32_t cd[RPW + 2];
    float he[RPW + 2]; // lanes 0 / 31: new direction of the west / east halo cell
#pragma unroll
    for (int k = 0; k < RPW + 2; ++k)
    {
      cd[k] = *reinterpret_cast<const uint32_t*>(sc + co + k * kCodeW);
      const float4 r4 = *reinterpret_cast<const float4*>(sr + fo + k * kHaloW);
      float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!first) p4 = *reinterpret_cast<const float4*>(sp + fo + k * kHaloW);
      pn[k] = direction4(r4, p4, cd[k], lut, inv5, beta);
      he[k] = 0.0f;
      if (edge && k >= 1 && k <= RPW)
      {
        const float inv = lut[sc[hco + k * kCodeW]].x;
        const float hp = first ? 0.0f : sp[hfo + k * kHaloW];
        he[k] = fmaf(beta, hp, inv * sr[hfo + k * kHaloW]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]); // everything is in registers
    if (++st == stages) { st = 0; ++round; }

    const int ci = t.tx * kTileW + (int)lane * 4;
    const int jb = sh.row_lo + t.ty * TH + r0;
#pragma unroll
    for (int k = 1; k <= RPW; ++k)
    {
      float w = __shfl_up_sync(0xffffffffu, pn[k].w, 1);
      float e = __shfl_down_sync(0xffffffffu, pn[k].x, 1);
      if (lane == 0) w = he[k];
      if (lane == 31) e = he[k];
      const float4 q = apply_a4(pn[k], w, e, pn[k - 1], pn[k + 1], cd[k], lut, diag5, off);
      const int j = jb + k - 1;
      if (j < sh.row_hi && ci < ld)
      {
        acc[0] += (double)dot4(pn[k], q);
        *reinterpret_cast<float4*>(p_new + (size_t)j * ld + ci) = pn[k];
        // slab boundary rows also go straight into the neighbour's ghost row (NVLink store)
        if (j == sh.row_lo && push_lo) *reinterpret_cast<float4*>(push_lo + ci) = pn[k];
        if (j == sh.row_hi - 1 && push_hi) *reinterpret_cast<float4*>(push_hi + ci) = pn[k];
      }
    }
  }

  double tot[1] = {0.0};
  if (fold_consumers<NW, 1>(acc, &s->ticket[1], partials, tot, pushed) && warp == 0)
  {
    // last CTA, warp 0: (sharded) exchange the slab sums with all ranks, then publish p.Ap
    double pq = __shfl_sync(0xffffffffu, tot[0], 0), unused = 0.0;
    bool ok = true;
    unsigned long long seq = 0;
    if (sh.world > 1)
    {
      seq = s->seq[0] + 1;
      if (lane == 0) __threadfence_system(); // every CTA's check-in (and its peer rows) precede the entry
      __syncwarp();
      ok = mail_allreduce_warp(sh, 0, seq, &pq, &unused);
    }
    if (lane == 0)
    {
      if (sh.world > 1) s->seq[0] = seq;
      if (!ok)
      {
        s->comm_error = 1;
        s->done = 1;
      }
      s->pq = pq;
    }
  }
}

// ------------------------------------------------------------ the update --
// alpha = absNew / p.Ap; q = A p RE-COMPUTED from the staged p tile (q is
// never stored); x += alpha p; r -= alpha q; partial |r|^2 and r.z.
// HBM traffic 21 B per cell: p, code, x, r in (TMA), x, r out (STG.128).
template <int NW, int RPW>
__global__ void __launch_bounds__((NW + 1) * 32)
k_cg_update(const __grid_constant__ CgMaps maps, float* __restrict__ x, float* __restrict__ r,
            int ld, int tiles_x, int n_tiles, int stages, const CgCoef coef,
            CgScalars* __restrict__ s, double* __restrict__ partials,
            const __grid_constant__ ShardArgs sh, float* __restrict__ push_lo,
            float* __restrict__ push_hi, int reverse)
{
  constexpr int TH = NW * RPW;
  using St = UpdStage<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  __shared__ float4 lut[8];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  load_lut(lut, coef);
  if (threadIdx.x == 0)
  {
    for (int k = 0; k < stages; ++k)
    {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], NW);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  if (s->done) return;
  const float alpha = s->abs_new / (float)s->pq; // Eigen: alpha = absNew / p.dot(tmp)
  const float nalpha = -alpha;
  const int* __restrict__ tile_list = s->tile_list;
  const int n_walk = tile_list ? s->n_active_tiles : n_tiles;
  const int n_prefix = tile_list ? s->n_prefix_tiles : 0;

  if (warp == NW)
  {
    if (lane == 0)
    {
      TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, reverse != 0, tile_list, n_prefix);
      int st = 0, round = 0;
      for (int tk = 0; tk < t.count; ++tk, t.next())
      {
        if (round > 0) mbar_wait(&empty[st], (round - 1) & 1);
        const int c0 = t.tx * kTileW, j0 = sh.row_lo + t.ty * TH;
        unsigned char* base = smem + st * St::kBytes;
        mbar_expect_tx(&full[st], St::kTx);
        tma_load_2d(base + St::oP, &maps.halo_a, c0 - 4, j0 - 1, &full[st]);
        tma_load_2d(base + St::oX, &maps.inner_a, c0, j0, &full[st]);
        tma_load_2d(base + St::oR, &maps.inner_b, c0, j0, &full[st]);
        tma_load_2d(base + St::oC, &maps.code, c0 - 16, j0 - 1, &full[st]);
        if (++st == stages) { st = 0; ++round; }
      }
    }
    return;
  }

  const float inv5 = coef.invdiag[4], diag5 = coef.diag[4], off = coef.off;
  const int r0 = (int)warp * RPW;
  const int fo = r0 * kHaloW + 4 + (int)lane * 4;
  const int co = (r0 + 1) * kCodeW + 16 + (int)lane * 4;
  const int io = r0 * kTileW + (int)lane * 4;
  const int hfo = r0 * kHaloW + (lane == 31 ? 4 + kTileW : 3);
  const bool edge = (lane == 0 || lane == 31);
  double acc[2] = {0.0, 0.0};
  TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, reverse != 0, tile_list, n_prefix);
  int st = 0, round = 0;
  bool pushed = false;
  const int last_ty = n_tiles / tiles_x - 1;
  for (int tk = 0; tk < t.count; ++tk, t.next())
  {
    const unsigned char* base = smem + st * St::kBytes;
    const float* sp = reinterpret_cast<const float*>(base + St::oP);
    const float* sx = reinterpret_cast<const float*>(base + St::oX);
    const float* sr = reinterpret_cast<const float*>(base + St::oR);
    const unsigned char* sc = base + St::oC;
    pushed |= (push_lo && t.ty == 0) || (push_hi && t.ty == last_ty);
    mbar_wait(&full[st], round & 1);

    float4 pc[RPW + 2], xo[RPW], ro[RPW];
    uint32_t cd[RPW];
    float he[RPW];
#pragma unroll
    for (int k = 0; k < RPW + 2; ++k) pc[k] = *reinterpret_cast<const float4*>(sp + fo + k * kHaloW);
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      cd[k] = *reinterpret_cast<const uint32_t*>(sc + co + k * kCodeW);
      xo[k] = *reinterpret_cast<const float4*>(sx + io + k * kTileW);
      ro[k] = *reinterpret_cast<const float4*>(sr + io + k * kTileW);
      he[k] = edge ? sp[hfo + (k + 1) * kHaloW] : 0.0f;
    }
    // the slot is released after the loop below has CONSUMED the staged values: an arrive does not wait
    // for shared-memory loads still in flight (see fsb_cg_one.cu)
    const int slot_done = st;
    if (++st == stages) { st = 0; ++round; }

    const int ci = t.tx * kTileW + (int)lane * 4;
    const int jb = sh.row_lo + t.ty * TH + r0;
#pragma unroll
    for (int k = 0; k < RPW; ++k)
    {
      const float4 p4 = pc[k + 1];
      float w = __shfl_up_sync(0xffffffffu, p4.w, 1);
      float e = __shfl_down_sync(0xffffffffu, p4.x, 1);
      if (lane == 0) w = he[k];
      if (lane == 31) e = he[k];
      const uint32_t c4 = cd[k];
      const int j = jb + k;
      // four non-liquid cells: x and r stay exactly zero, nothing to write; rows past the
      // slab end belong to the neighbour
      if (c4 != 0 && j < sh.row_hi)
      {
        const float4 q = apply_a4(p4, w, e, pc[k], pc[k + 2], c4, lut, diag5, off);
        float4 xn, rn;
        xn.x = fmaf(alpha, p4.x, xo[k].x); rn.x = fmaf(nalpha, q.x, ro[k].x);
        xn.y = fmaf(alpha, p4.y, xo[k].y); rn.y = fmaf(nalpha, q.y, ro[k].y);
        xn.z = fmaf(alpha, p4.z, xo[k].z); rn.z = fmaf(nalpha, q.z, ro[k].z);
        xn.w = fmaf(alpha, p4.w, xo[k].w); rn.w = fmaf(nalpha, q.w, ro[k].w);
        const size_t o = (size_t)j * ld + ci; // c4 != 0 implies inside the grid
        *reinterpret_cast<float4*>(x + o) = xn;
        *reinterpret_cast<float4*>(r + o) = rn;
        if (j == sh.row_lo && push_lo) *reinterpret_cast<float4*>(push_lo + ci) = rn;
        if (j == sh.row_hi - 1 && push_hi) *reinterpret_cast<float4*>(push_hi + ci) = rn;
        float4 z;
        if (c4 == kInterior4) z = make_float4(inv5 * rn.x, inv5 * rn.y, inv5 * rn.z, inv5 * rn.w);
        else
          z = make_float4(lut[c4 & 0xff].x * rn.x, lut[(c4 >> 8) & 0xff].x * rn.y,
                          lut[(c4 >> 16) & 0xff].x * rn.z, lut[c4 >> 24].x * rn.w);
        acc[0] += (double)dot4(rn, rn);
        acc[1] += (double)dot4(rn, z);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot_done]);
  }

  double tot[2] = {0.0, 0.0};
  if (fold_consumers<NW, 2>(acc, &s->ticket[2], partials, tot, pushed) && warp == 0)
  {
    double tr2 = __shfl_sync(0xffffffffu, tot[0], 0), trz = __shfl_sync(0xffffffffu, tot[1], 0);
    bool ok = true;
    unsigned long long seq = 0;
    if (sh.world > 1)
    {
      seq = s->seq[1] + 1;
      if (lane == 0) __threadfence_system();
      __syncwarp();
      ok = mail_allreduce_warp(sh, 1, seq, &tr2, &trz);
    }
    if (lane == 0)
    {
      if (sh.world > 1) s->seq[1] = seq;
      if (ok) finalize_update(s, tr2, trz);
      else
      {
        s->comm_error = 1;
        s->done = 1;
      }
    }
  }
}

// ------------------------------------------------ persistent fused solve --
// The whole CG loop as ONE cooperative kernel: every CTA stays resident and runs the
// direction phase and the update phase of every iteration over its tile list (same TMA
// ring, same per-tile arithmetic as the two kernels above), with a software grid barrier
// after each phase that carries the reduction.
//
// Barrier = reduction, with no serial hop through a "last" CTA:
//   * a CTA stores its partial sums, fences, and adds 1 to a monotone arrival counter;
//   * single GPU: warp 0 of EVERY CTA spins on the counter, then folds all the partials
//     itself in a fixed order (strided per lane + xor butterfly: fp addition commutes, so
//     all lanes and all CTAs get the same bits) and advances ITS OWN copy of the CG scalars
//     (alpha, beta, iteration count, done) in shared memory -- nothing is broadcast;
//   * sharded: only CTA 0 waits for the counter and folds; its warp posts the slab sum into
//     every rank's mailbox (lane q -> rank q, NVLink stores in flight together) and warp 0 of
//     every CTA of every rank polls its OWN rank's mailbox (local HBM) for the `world` tagged
//     entries and adds them in rank order -> bit-identical scalars everywhere.  A rank's entry
//     implies all of its CTAs arrived (and fenced their peer-row stores at system scope).
//   * the producer thread meanwhile has already issued the next phase's loads that do not
//     depend on the barrier (phase A: p_old, code; phase B: x, r, code -- written by this same
//     CTA or before the previous barrier) for a full ring of tiles; only the one dependent
//     array (A: r halo, B: p_new halo) is requested after the release.
struct SolveMaps
{
  CUtensorMap halo_r, halo_p[2], inner_x, inner_r, code;
  CUtensorMap inner_p[2]; // the deferred x update reads the previous direction without halo
};
struct SolvePush // peer rows receiving this rank's slab boundary rows (null: no neighbour)
{
  float *p_lo[2], *p_hi[2], *r_lo, *r_hi;
};


// per-CTA copy of the CG scalars: every CTA derives the same values from the same totals
struct SolveState
{
  double pq, r2, rz;
  unsigned long long seq[2];
  float abs_new, abs_old, beta, alpha, thr;
  int iter, done, max_iters, comm_error;
  volatile unsigned int released; // last phase whose reduction this CTA has completed
};

// Barrier + reduction for the NW consumer warps of every CTA (see above).  TYPE 0: p.Ap,
// TYPE 1: (|r|^2, r.z).  `partials` is this TYPE's own region (N * gridDim.x doubles): a region
// is rewritten two barriers later, after every reader has arrived at the barrier in between.
template <int NW, int N, int TYPE>
__device__ __forceinline__ void grid_reduce(double (&acc)[N], CgScalars* s,
                                            double* __restrict__ partials, unsigned phase_id,
                                            const ShardArgs& sh, SolveState* ss, bool pushed)
{
  __shared__ double s_part[N][32];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // every thread orders its own generic-proxy stores against the async proxy (a proxy fence is not
  // cumulative, see one_reduce in fsb_cg_one.cu)
  fence_proxy_async_all();
#pragma unroll
  for (int n = 0; n < N; ++n)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_down_sync(0xffffffffu, acc[n], o);
    if (lane == 0) s_part[n][warp] = acc[n];
  }
  consumer_sync(NW * 32);
  if (warp == 0)
  {
    const int G = (int)gridDim.x;
    double tot[N];
#pragma unroll
    for (int n = 0; n < N; ++n)
    {
      double v = (lane < NW) ? s_part[n][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      tot[n] = v;
    }
    if (lane == 0)
    {
#pragma unroll
      for (int n = 0; n < N; ++n) partials[n * G + blockIdx.x] = tot[n];
      // the CTA's global stores (p / x / r rows, peer rows) must be visible to the TMA loads of
      // the next phase on every SM (and GPU) before the arrival is
      fence_proxy_async_all();
      if (pushed) __threadfence_system();
      else __threadfence();
      atomicAdd(&s->bar_count, 1u);
    }
    __syncwarp();
    if (sh.world == 1 || blockIdx.x == 0)
    {
      const volatile unsigned int* cnt = &s->bar_count;
      const unsigned int target = phase_id * (unsigned int)G;
      {
        SpinGuard g;
        while (*cnt < target) g.tick();
      }
      __threadfence();
      const volatile double* part = partials;
#pragma unroll
      for (int n = 0; n < N; ++n)
      {
        double v = 0.0;
        for (int k = (int)lane; k < G; k += 32) v += part[n * G + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        tot[n] = v;
      }
    }
    bool ok = true;
    const unsigned long long seq = ss->seq[TYPE] + 1;
    if (sh.world > 1)
    {
      const unsigned long long tag = seq & 0xffffffffull;
      if (blockIdx.x == 0)
      {
        // every local CTA's arrival (and its peer rows, fenced at system scope) precedes the entry
        __threadfence_system();
        if ((int)lane < sh.world)
        {
          const unsigned long long b0 = (unsigned long long)__double_as_longlong(tot[0]);
          const unsigned long long b1 = (unsigned long long)__double_as_longlong(tot[N - 1]);
          volatile unsigned long long* out = reinterpret_cast<volatile unsigned long long*>(
              sh.mail[lane] + TYPE * kMaxRanks + sh.rank);
          out[0] = mail_word((unsigned int)b0, seq);
          out[1] = mail_word((unsigned int)(b0 >> 32), seq);
          out[2] = mail_word((unsigned int)b1, seq);
          out[3] = mail_word((unsigned int)(b1 >> 32), seq);
        }
      }
      double a = 0.0, b = 0.0;
      if ((int)lane < sh.world)
      {
        volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(
            sh.mail[sh.rank] + TYPE * kMaxRanks + lane);
        const unsigned long long t0 = global_ns();
        unsigned long long w0, w1, w2, w3;
        unsigned int spins = 0;
        for (;;)
        {
          w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
          if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
          if ((++spins & 1023u) == 0 && global_ns() - t0 > kMailTimeoutNs)
          {
            ok = false;
            break;
          }
        }
        a = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        b = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
      }
      ok = __all_sync(0xffffffffu, ok);
      __threadfence_system(); // acquire: the peers' boundary rows were stored before their entries
      double s0 = 0.0, s1 = 0.0;
      for (int q = 0; q < sh.world; ++q) // rank order: identical totals on every rank
      {
        s0 += __shfl_sync(0xffffffffu, a, q);
        s1 += __shfl_sync(0xffffffffu, b, q);
      }
      tot[0] = s0;
      if (N > 1) tot[N - 1] = s1;
    }
    if (lane == 0)
    {
      ss->seq[TYPE] = seq;
      if (!ok)
      {
        ss->comm_error = 1;
        ss->done = 1;
      }
      else if (TYPE == 0)
      {
        ss->pq = tot[0];
        ss->alpha = ss->abs_new / (float)tot[0]; // Eigen: alpha = absNew / p.dot(tmp)
      }
      else
      {
        // finalize_update on the CTA's own copy
        ss->r2 = tot[0];
        ss->rz = tot[N - 1];
        if ((float)tot[0] < ss->thr) ss->done = 1; // converged: Eigen breaks before i++
        else
        {
          ss->abs_old = ss->abs_new;
          ss->abs_new = (float)tot[N - 1];
          ss->beta = ss->abs_new / ss->abs_old;
          ss->iter = ss->iter + 1;
          if (ss->iter >= ss->max_iters) ss->done = 1;
        }
      }
      __threadfence_block();
      ss->released = phase_id;
    }
  }
  consumer_sync(NW * 32);
}


template <int NW, int RPW>
__global__ void __launch_bounds__((NW + 1) * 32)
k_cg_solve(const __grid_constant__ SolveMaps maps, float* __restrict__ x, float* __restrict__ r,
           float* __restrict__ p0, float* __restrict__ p1, int ld, int tiles_x, int n_tiles,
           int stages, int stage_bytes, const CgCoef coef, CgScalars* __restrict__ s,
           double* __restrict__ partials, const __grid_constant__ ShardArgs sh,
           const __grid_constant__ SolvePush push, int flags)
{
  constexpr int TH = NW * RPW;
  using Sd = DirStage<TH>;
  using Su = UpdStageX<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  __shared__ float4 lut[8];
  __shared__ SolveState ss;
  // Deferred x update: x is touched only in the update phase of ODD iterations, where
  // x += alpha_{k-1} p_{k-1} and x += alpha_k p_k are applied back to back (the previous direction
  // is still intact in the other ping-pong buffer) -- the same two roundings in the same order,
  // so x is bit-identical, for 12 instead of 16 B per cell and iteration pair.  A solve that ends
  // on an even iteration applies the pending update in a final sweep.
  const bool xdefer = (flags & 128) != 0;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool serp = (flags & 1) != 0;     // phase B walks the tile list backwards
  const bool xhint = (flags & 2) != 0;    // x is pure streaming: evict-first loads and stores
  const bool prefetch = (flags & 4) != 0; // barrier-independent loads issued before the barrier
  // L2 residency: r and the stencil codes are touched by BOTH sweeps of every iteration and are the
  // only data worth keeping when the vectors do not all fit (4096^2: r 67 MB + code 17 MB of the
  // 126 MB L2) -> evict-last on a fraction keep/4 of their loads and stores; the old direction is
  // dead once phase A has read it -> evict-first
  const bool phint = (flags & 8) != 0;
  const int keep = (flags >> 4) & 7;
  double* part_a = partials;              // phase A region: gridDim.x doubles
  double* part_b = partials + gridDim.x;  // phase B region: 2 * gridDim.x doubles

  load_lut(lut, coef);
  if (threadIdx.x == 0)
  {
    for (int k = 0; k < stages; ++k)
    {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], NW);
    }
    fence_barrier_init();
    volatile CgScalars* vs = s;
    ss.pq = 0.0; ss.r2 = vs->r2; ss.rz = vs->rz;
    ss.seq[0] = vs->seq[0]; ss.seq[1] = vs->seq[1];
    ss.abs_new = vs->abs_new; ss.abs_old = vs->abs_old; ss.beta = vs->beta; ss.alpha = 0.0f;
    ss.thr = vs->thr;
    ss.iter = vs->iter; ss.done = vs->done; ss.max_iters = vs->max_iters; ss.comm_error = 0;
    ss.released = 0;
  }
  __syncthreads();
  unsigned phase_id = 0;
  const int* __restrict__ tile_list = s->tile_list; // fixed for the whole solve
  const int n_walk = tile_list ? s->n_active_tiles : n_tiles;
  const int n_prefix = tile_list ? s->n_prefix_tiles : 0;

  if (warp == NW)
  {
    // ---- producer: one elected lane
    if (lane != 0) return;
    const uint64_t pol_x = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last(keep);
    // a load with or without an L2 policy
    auto load = [&](void* dst, const CUtensorMap* map, int c0, int j0, uint64_t* bar, int hint) {
      if (hint == 1) tma_load_2d_hint(dst, map, c0, j0, bar, pol_x);
      else if (hint == 2) tma_load_2d_hint(dst, map, c0, j0, bar, pol_keep);
      else tma_load_2d(dst, map, c0, j0, bar);
    };
    const int h_keep = keep ? 2 : 0, h_x = xhint ? 1 : 0, h_pold = phint ? 1 : 0;
    RingPos rp = {0, 0};  // next slot to allocate
    RingPos pre = {0, 0}; // first slot of the tiles whose independent loads are already out
    int npre = 0;
    int cur = 0;
    bool first = ss.iter == 0;
    int kind = 0; // 0: phase A (direction), 1: phase B (update)
    if (ss.done) return;
    // part: 1 = loads that do not depend on the barrier, 2 = the dependent one, 3 = both
    int it_p = 0; // iteration the producer is issuing loads for (a solve starts at iteration 0)
    auto issue = [&](int knd, int part, bool frst, int cr, int slot, int c0, int j0) {
      unsigned char* base = smem + slot * stage_bytes;
      const bool with_x = !xdefer || (it_p & 1);
      if (knd == 0)
      {
        if (part & 1)
        {
          mbar_expect_tx(&full[slot], frst ? Sd::kTx - Sd::kF32 : Sd::kTx);
          if (!frst) load(base + Sd::oP, &maps.halo_p[cr], c0 - 4, j0 - 1, &full[slot], h_pold);
          load(base + Sd::oC, &maps.code, c0 - 16, j0 - 1, &full[slot], h_keep);
        }
        if (part & 2) load(base + Sd::oR, &maps.halo_r, c0 - 4, j0 - 1, &full[slot], h_keep);
      }
      else
      {
        if (part & 1)
        {
          mbar_expect_tx(&full[slot], Su::kTx - (with_x ? 0 : Su::kInner) +
                                          (with_x && xdefer ? Su::kInner : 0));
          if (with_x) load(base + Su::oX, &maps.inner_x, c0, j0, &full[slot], h_x);
          if (with_x && xdefer) load(base + Su::oQ, &maps.inner_p[cr], c0, j0, &full[slot], h_pold);
          load(base + Su::oR, &maps.inner_r, c0, j0, &full[slot], h_keep);
          load(base + Su::oC, &maps.code, c0 - 16, j0 - 1, &full[slot], h_keep);
        }
        if (part & 2) tma_load_2d(base + Su::oP, &maps.halo_p[cr ^ 1], c0 - 4, j0 - 1, &full[slot]);
      }
    };
    for (;;)
    {
      TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, serp && kind == 1, tile_list, n_prefix);
      // 1. the dependent load of the tiles that were started before the barrier
      RingPos q = pre;
      int k = 0;
      for (; k < npre; ++k, t.next())
      {
        issue(kind, 2, first, cur, q.st, t.tx * kTileW, sh.row_lo + t.ty * TH);
        q.advance(stages);
      }
      // 2. the rest of this phase's tiles
      for (; k < t.count; ++k, t.next())
      {
        if (rp.round > 0) mbar_wait_guarded(&empty[rp.st], (rp.round - 1) & 1);
        issue(kind, 3, first, cur, rp.st, t.tx * kTileW, sh.row_lo + t.ty * TH);
        rp.advance(stages);
      }
      // 3. start the next phase's tiles: everything that does not depend on this phase's
      //    reduction or on other CTAs' stores of this phase
      const int nkind = kind ^ 1;
      const int ncur = kind == 1 ? cur ^ 1 : cur;
      npre = 0;
      pre = rp;
      if (prefetch)
      {
        TileWalk tn(blockIdx.x, gridDim.x, tiles_x, n_walk, serp && nkind == 1, tile_list, n_prefix);
        const int want = min(stages, tn.count);
        for (; npre < want; ++npre, tn.next())
        {
          if (rp.round > 0) mbar_wait_guarded(&empty[rp.st], (rp.round - 1) & 1);
          issue(nkind, 1, false, ncur, rp.st, tn.tx * kTileW, sh.row_lo + tn.ty * TH);
          rp.advance(stages);
        }
      }
      // 4. the barrier
      ++phase_id;
      {
        SpinGuard g;
        while (ss.released < phase_id) g.tick();
      }
      fence_proxy_async_all();
      const bool done = ss.done != 0;
      if (kind == 1) ++it_p;
      kind = nkind;
      cur = ncur;
      first = false;
      if (done)
      {
        // nobody will consume the started tiles: complete them before the CTA may exit
        TileWalk tn(blockIdx.x, gridDim.x, tiles_x, n_walk, serp && kind == 1, tile_list, n_prefix);
        RingPos w = pre;
        for (int m = 0; m < npre; ++m, tn.next())
        {
          issue(kind, 2, false, cur, w.st, tn.tx * kTileW, sh.row_lo + tn.ty * TH);
          mbar_wait_guarded(&full[w.st], w.round & 1);
          w.advance(stages);
        }
        return;
      }
    }
  }

  // ---- consumers
  const uint64_t pol_x = l2_policy_evict_first();
  const uint64_t pol_keep = l2_policy_evict_last(keep);
  const float inv5 = coef.invdiag[4], diag5 = coef.diag[4], off = coef.off;
  const int r0 = (int)warp * RPW;
  const int fo = r0 * kHaloW + 4 + (int)lane * 4;
  const int hfo = r0 * kHaloW + (lane == 31 ? 4 + kTileW : 3);
  const int hco = r0 * kCodeW + (lane == 31 ? 16 + kTileW : 15);
  const int co = r0 * kCodeW + 16 + (int)lane * 4;
  const int io = r0 * kTileW + (int)lane * 4;
  const bool edge = (lane == 0 || lane == 31);
  const int last_ty = n_tiles / tiles_x - 1;
  RingPos rp = {0, 0};
  int cur = 0;
  float alpha_prev = 0.0f;       // alpha of the previous iteration (pending x update)
  bool pending = false;          // x lacks the update of the last iteration
  const float* last_p = nullptr; // direction of the last iteration
  while (!ss.done)
  {
    const bool with_x = !xdefer || (ss.iter & 1);
    // ================= phase A: direction + p.Ap =================
    {
      const bool first = ss.iter == 0;
      const float beta = first ? 0.0f : ss.beta;
      float* __restrict__ p_new = cur ? p0 : p1;
      last_p = p_new;
      float* push_lo = push.p_lo[cur ^ 1];
      float* push_hi = push.p_hi[cur ^ 1];
      double acc[1] = {0.0};
      bool pushed = false;
      TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, false, tile_list, n_prefix);
      for (int tk = 0; tk < t.count; ++tk, t.next())
      {
        const unsigned char* base = smem + rp.st * stage_bytes;
        const float* sr = reinterpret_cast<const float*>(base + Sd::oR);
        const float* sp = reinterpret_cast<const float*>(base + Sd::oP);
        const unsigned char* sc = base + Sd::oC;
        pushed |= (push_lo && t.ty == 0) || (push_hi && t.ty == last_ty);
        mbar_wait_guarded(&full[rp.st], rp.round & 1);
        float4 pn[RPW + 2];
        uint32_t cd[RPW + 2];
        float he[RPW + 2];
#pragma unroll
        for (int k = 0; k < RPW + 2; ++k)
        {
          cd[k] = *reinterpret_cast<const uint32_t*>(sc + co + k * kCodeW);
          const float4 r4 = *reinterpret_cast<const float4*>(sr + fo + k * kHaloW);
          float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!first) p4 = *reinterpret_cast<const float4*>(sp + fo + k * kHaloW);
          pn[k] = direction4(r4, p4, cd[k], lut, inv5, beta);
          he[k] = 0.0f;
          if (edge && k >= 1 && k <= RPW)
          {
            const float inv = lut[sc[hco + k * kCodeW]].x;
            const float hp = first ? 0.0f : sp[hfo + k * kHaloW];
            he[k] = fmaf(beta, hp, inv * sr[hfo + k * kHaloW]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[rp.st]);
        rp.advance(stages);
        const int ci = t.tx * kTileW + (int)lane * 4;
        const int jb = sh.row_lo + t.ty * TH + r0;
#pragma unroll
        for (int k = 1; k <= RPW; ++k)
        {
          float w = __shfl_up_sync(0xffffffffu, pn[k].w, 1);
          float e = __shfl_down_sync(0xffffffffu, pn[k].x, 1);
          if (lane == 0) w = he[k];
          if (lane == 31) e = he[k];
          const float4 q = apply_a4(pn[k], w, e, pn[k - 1], pn[k + 1], cd[k], lut, diag5, off);
          const int j = jb + k - 1;
          if (j < sh.row_hi && ci < ld)
          {
            acc[0] += (double)dot4(pn[k], q);
            *reinterpret_cast<float4*>(p_new + (size_t)j * ld + ci) = pn[k];
            if (j == sh.row_lo && push_lo) *reinterpret_cast<float4*>(push_lo + ci) = pn[k];
            if (j == sh.row_hi - 1 && push_hi) *reinterpret_cast<float4*>(push_hi + ci) = pn[k];
          }
        }
      }
      ++phase_id;
      grid_reduce<NW, 1, 0>(acc, s, part_a, phase_id, sh, &ss, pushed);
      if (ss.done) break; // only a communication failure ends the solve here
    }
    // ================= phase B: update + |r|^2, r.z =================
    {
      const float alpha = ss.alpha, nalpha = -alpha;
      double acc[2] = {0.0, 0.0};
      bool pushed = false;
      TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, serp, tile_list, n_prefix);
      for (int tk = 0; tk < t.count; ++tk, t.next())
      {
        const unsigned char* base = smem + rp.st * stage_bytes;
        const float* sp = reinterpret_cast<const float*>(base + Su::oP);
        const float* sx = reinterpret_cast<const float*>(base + Su::oX);
        const float* sq = reinterpret_cast<const float*>(base + Su::oQ);
        const float* sr = reinterpret_cast<const float*>(base + Su::oR);
        const unsigned char* sc = base + Su::oC;
        pushed |= (push.r_lo && t.ty == 0) || (push.r_hi && t.ty == last_ty);
        mbar_wait_guarded(&full[rp.st], rp.round & 1);
        float4 pc[RPW + 2], xo[RPW], ro[RPW];
        uint32_t cd[RPW];
        float he[RPW];
#pragma unroll
        for (int k = 0; k < RPW + 2; ++k)
          pc[k] = *reinterpret_cast<const float4*>(sp + fo + k * kHaloW);
#pragma unroll
        for (int k = 0; k < RPW; ++k)
        {
          cd[k] = *reinterpret_cast<const uint32_t*>(sc + co + (k + 1) * kCodeW);
          xo[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (with_x)
          {
            xo[k] = *reinterpret_cast<const float4*>(sx + io + k * kTileW);
            if (xdefer)
            {
              // the pending update of the previous iteration first (same rounding order as before)
              const float4 pp = *reinterpret_cast<const float4*>(sq + io + k * kTileW);
              xo[k].x = fmaf(alpha_prev, pp.x, xo[k].x);
              xo[k].y = fmaf(alpha_prev, pp.y, xo[k].y);
              xo[k].z = fmaf(alpha_prev, pp.z, xo[k].z);
              xo[k].w = fmaf(alpha_prev, pp.w, xo[k].w);
            }
          }
          ro[k] = *reinterpret_cast<const float4*>(sr + io + k * kTileW);
          he[k] = edge ? sp[hfo + (k + 1) * kHaloW] : 0.0f;
        }
        const int slot_done = rp.st; // released after the loop below has consumed the staged values
        rp.advance(stages);
        const int ci = t.tx * kTileW + (int)lane * 4;
        const int jb = sh.row_lo + t.ty * TH + r0;
#pragma unroll
        for (int k = 0; k < RPW; ++k)
        {
          const float4 p4 = pc[k + 1];
          float w = __shfl_up_sync(0xffffffffu, p4.w, 1);
          float e = __shfl_down_sync(0xffffffffu, p4.x, 1);
          if (lane == 0) w = he[k];
          if (lane == 31) e = he[k];
          const uint32_t c4 = cd[k];
          const int j = jb + k;
          if (c4 != 0 && j < sh.row_hi)
          {
            const float4 q = apply_a4(p4, w, e, pc[k], pc[k + 2], c4, lut, diag5, off);
            float4 xn, rn;
            xn.x = fmaf(alpha, p4.x, xo[k].x); rn.x = fmaf(nalpha, q.x, ro[k].x);
            xn.y = fmaf(alpha, p4.y, xo[k].y); rn.y = fmaf(nalpha, q.y, ro[k].y);
            xn.z = fmaf(alpha, p4.z, xo[k].z); rn.z = fmaf(nalpha, q.z, ro[k].z);
            xn.w = fmaf(alpha, p4.w, xo[k].w); rn.w = fmaf(nalpha, q.w, ro[k].w);
            const size_t o = (size_t)j * ld + ci;
            if (with_x)
            {
              if (xhint) st_f4_hint(x + o, xn, pol_x);
              else *reinterpret_cast<float4*>(x + o) = xn;
            }
            if (keep) st_f4_hint(r + o, rn, pol_keep);
            else *reinterpret_cast<float4*>(r + o) = rn;
            if (j == sh.row_lo && push.r_lo) *reinterpret_cast<float4*>(push.r_lo + ci) = rn;
            if (j == sh.row_hi - 1 && push.r_hi) *reinterpret_cast<float4*>(push.r_hi + ci) = rn;
            float4 z;
            if (c4 == kInterior4) z = make_float4(inv5 * rn.x, inv5 * rn.y, inv5 * rn.z, inv5 * rn.w);
            else
              z = make_float4(lut[c4 & 0xff].x * rn.x, lut[(c4 >> 8) & 0xff].x * rn.y,
                              lut[(c4 >> 16) & 0xff].x * rn.z, lut[c4 >> 24].x * rn.w);
            acc[0] += (double)dot4(rn, rn);
            acc[1] += (double)dot4(rn, z);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot_done]);
      }
      ++phase_id;
      grid_reduce<NW, 2, 1>(acc, s, part_b, phase_id, sh, &ss, pushed);
      alpha_prev = alpha;
      pending = !with_x;
    }
    cur ^= 1;
  }
  if (pending && !ss.comm_error)
  {
    // the solve ended on an even iteration: x += alpha p of that iteration.  Same tile -> thread
    // mapping as the phases (this thread wrote these p and x elements itself); p is exactly zero
    // outside LIQUID cells, so masked cells keep x = 0.
    TileWalk t(blockIdx.x, gridDim.x, tiles_x, n_walk, false, tile_list, n_prefix);
    for (int tk = 0; tk < t.count; ++tk, t.next())
    {
      const int ci = t.tx * kTileW + (int)lane * 4;
      const int jb = sh.row_lo + t.ty * TH + r0;
#pragma unroll
      for (int k = 0; k < RPW; ++k)
      {
        const int j = jb + k;
        if (j < sh.row_hi && ci < ld)
        {
          const size_t o = (size_t)j * ld + ci;
          const float4 p4 = *reinterpret_cast<const float4*>(last_p + o);
          float4 x4 = *reinterpret_cast<const float4*>(x + o);
          x4.x = fmaf(alpha_prev, p4.x, x4.x);
          x4.y = fmaf(alpha_prev, p4.y, x4.y);
          x4.z = fmaf(alpha_prev, p4.z, x4.z);
          x4.w = fmaf(alpha_prev, p4.w, x4.w);
          *reinterpret_cast<float4*>(x + o) = x4;
        }
      }
    }
  }
  // every CTA holds the same final scalars; CTA 0 publishes them for the host
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    s->pq = ss.pq; s->r2 = ss.r2; s->rz = ss.rz;
    s->abs_new = ss.abs_new; s->abs_old = ss.abs_old; s->beta = ss.beta;
    s->iter = ss.iter;
    s->done = ss.done;
    s->seq[0] = ss.seq[0]; s->seq[1] = ss.seq[1];
    if (ss.comm_error) s->comm_error = 1;
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
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= d.nx || j >= d.ny) return;
  const int im1 = clampi(i - 1, 0, d.nx - 1);
  const int jm1 = clampi(j - 1, 0, d.ny - 1);
  const size_t k = i + (size_t)j * d.ld, kw = im1 + (size_t)j * d.ld, ks = i + (size_t)jm1 * d.ld;
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



int pick_tile_rows(const fsb_ctx* c)
{
  if (const char* e = getenv("FSB_CG_TILE_ROWS")) // tuning knob for profiling runs
  {
    const int th = atoi(e);
    if (th == 8 || th == 16 || th == 32) return th;
  }
  const int tiles_x = fsb_div_up(c->ld, kTileW);
  const int rows = c->shard.row_hi - c->shard.row_lo;
  if ((int64_t)tiles_x * fsb_div_up(rows, 16) >= (int64_t)4 * c->sm_count) return 16;
  return 8;
}


template <int RPW>
int configure_kernels(fsb_ctx* c, int64_t n_tiles)
{
  constexpr int TH = kNW * RPW;
  const int threads = (kNW + 1) * 32;
  // ring depth: as many stages as two resident CTAs per SM allow (at least 2, at most kMaxStages)
  const int budget = (227 * 1024 - 2 * 2048) / 2;
  c->cg_stages_dir = std::max(2, std::min(kMaxStages, budget / DirStage<TH>::kBytes));
  c->cg_stages_upd = std::max(2, std::min(kMaxStages, budget / UpdStage<TH>::kBytes));
  if (const char* e = getenv("FSB_CG_STAGES")) // tuning knob for profiling runs
  {
    const int v = atoi(e);
    if (v >= 2 && v <= kMaxStages) c->cg_stages_dir = c->cg_stages_upd = v;
  }
  const int smem_dir = c->cg_stages_dir * DirStage<TH>::kBytes;
  const int smem_upd = c->cg_stages_upd * UpdStage<TH>::kBytes;
  if (smem_dir > 227 * 1024 - 2048 || smem_upd > 227 * 1024 - 2048)
    return fsb_fail(c, FSB_ERR_INVALID, "CG ring of %d/%d stages does not fit shared memory",
                    c->cg_stages_dir, c->cg_stages_upd);
  auto kd = k_cg_direction<kNW, RPW>;
  auto ku = k_cg_update<kNW, RPW>;
  FSB_CUDA(c, cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dir));
  FSB_CUDA(c, cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_upd));
  int occ_dir = 1, occ_upd = 1;
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, kd, threads, smem_dir));
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, ku, threads, smem_upd));
  int cap = 4; // CTAs per SM worth keeping resident (each already keeps STAGES-1 tiles in flight)
  if (const char* e = getenv("FSB_CG_CTAS_PER_SM")) cap = std::max(1, atoi(e));
  occ_dir = std::max(1, std::min(occ_dir, cap));
  occ_upd = std::max(1, std::min(occ_upd, cap));
  // persistent-style grids: exactly one resident wave, each CTA walks the tile list
  c->cg_grid_dir = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ_dir);
  c->cg_grid_upd = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ_upd);
  return FSB_OK;
}

template <int RPW>
int configure_fused(fsb_ctx* c, int64_t n_tiles)
{
  constexpr int TH = kNW * RPW;
  const int threads = (kNW + 1) * 32;
  c->cg_fused_stage_bytes = std::max(DirStage<TH>::kBytes, UpdStageX<TH>::kBytes);
  const int budget = (227 * 1024 - 2 * 2048) / 2; // two resident CTAs per SM
  int stages = std::max(2, std::min(kMaxStages, budget / c->cg_fused_stage_bytes));
  if (const char* e = getenv("FSB_CG_STAGES"))
  {
    const int v = atoi(e);
    if (v >= 2 && v <= kMaxStages) stages = v;
  }
  const int smem = stages * c->cg_fused_stage_bytes;
  if (smem > 227 * 1024 - 2048)
    return fsb_fail(c, FSB_ERR_INVALID, "CG ring of %d stages does not fit shared memory", stages);
  auto kf = k_cg_solve<kNW, RPW>;
  FSB_CUDA(c, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 1;
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, threads, smem));
  int cap = 4;
  if (const char* e = getenv("FSB_CG_CTAS_PER_SM")) cap = std::max(1, atoi(e));
  occ = std::max(1, std::min(occ, cap));
  c->cg_fused_stages = stages;
  c->cg_grid_fused = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ);
  return FSB_OK;
}

int configure_cg(fsb_ctx* c)
{
  if (c->cg_tile_rows != 0) return FSB_OK;
  if (c->shard.world == 1)
  {
    c->shard.row_lo = 0;
    c->shard.row_hi = c->ny;
  }
  int th = pick_tile_rows(c);
  // the persistent kernels need every CTA resident at once: cooperative launch
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  {
    const char* mode = getenv("FSB_CG_MODE");
    c->cg_one = coop != 0 && (!mode || mode[0] == 'o');
  }
  if (c->cg_one)
  {
    // the one-sweep kernel runs kNWOne consumer warps per CTA: same rows per warp, shorter tiles
    th = th / 8 * kNWOne;
    c->cg_grid_dir = c->cg_grid_upd = c->cg_grid_fused = 0;
  }
  else
  {
    const int64_t n_tiles = (int64_t)fsb_div_up(c->ld, kTileW) *
                            fsb_div_up(c->shard.row_hi - c->shard.row_lo, th);
    if (th == 32) FSB_TRY(configure_kernels<4>(c, n_tiles));
    else if (th == 16) FSB_TRY(configure_kernels<2>(c, n_tiles));
    else FSB_TRY(configure_kernels<1>(c, n_tiles));
    if (th == 32) FSB_TRY(configure_fused<4>(c, n_tiles));
    else if (th == 16) FSB_TRY(configure_fused<2>(c, n_tiles));
    else FSB_TRY(configure_fused<1>(c, n_tiles));
  }
  {
    // Default: the persistent single-kernel solve (k_cg_solve) -- fastest at every size measured
    // (profiles/r01g).  FSB_CG_MODE=graph selects two launches per iteration in a CUDA graph.
    // Default: the one-sweep solve (fsb_cg_one.cu: one sweep and one reduction point per iteration,
    // 23 B per cell).  FSB_CG_MODE=fused selects the two-sweep persistent kernel of this file
    // (Eigen's two reduction points, 32 B per cell), FSB_CG_MODE=graph its two-kernel form.
    const char* mode = getenv("FSB_CG_MODE");
    c->cg_fused = coop != 0 && !(mode && mode[0] == 'g');
    // Sharded solves on short slabs are bound by the two cross-GPU reductions per iteration, and the
    // kernel-boundary form of that handshake (last CTA + one-warp combine) is lighter than the
    // in-kernel one where every CTA polls the mailbox: measured on 2 B200, 8.4 M cells per rank
    // 64.8 us (graph) vs 68.5 us (persistent), 33.5 M cells per rank 209.5 vs 194.0 us
    // (profiles/r01h_2gpu.md).  FSB_CG_MODE=fused / graph overrides.
    if (!mode && !c->cg_one && c->shard.world > 1 &&
        (int64_t)(c->shard.row_hi - c->shard.row_lo) * c->ld < (int64_t)12 * 1000 * 1000)
      c->cg_fused = false;
    const char* pdl = getenv("FSB_CG_PDL"); // profiling knob: 0 disables dependent launch
    c->cg_pdl = !(pdl && pdl[0] == '0');
    // bit 0: serpentine sweeps (the update walks the tile list backwards), bit 1: x loads / stores
    // carry an L2 evict-first hint, bit 2: the fused kernel issues barrier-independent loads early
    auto knob = [](const char* name, int dflt) {
      const char* e = getenv(name);
      return e ? atoi(e) != 0 : dflt != 0;
    };
    // bit 3: the old direction is loaded evict-first; bits 4-6: r and the stencil codes are kept
    // evict-last on a fraction keep/4 of their accesses (FSB_CG_KEEP = 0..4)
    int keep = 0;
    if (const char* e = getenv("FSB_CG_KEEP")) keep = std::max(0, std::min(4, atoi(e)));
    c->cg_persist_mb = 0;
    if (const char* e = getenv("FSB_CG_PERSIST_MB")) c->cg_persist_mb = std::max(0, atoi(e));
    c->cg_persist_miss_normal = knob("FSB_CG_PERSIST_MISS_NORMAL", 0);
    c->cg_skip_tiles = knob("FSB_CG_SKIP_TILES", 1);
    // measured on 2 B200 (profiles/r01i_notes.md): no gain (68.4 vs 68.1 us at 4096^2, 202.2 vs
    // 201.1 us at 8192^2) -- the system-scope fence after the boundary stores is not what the
    // sharded iteration waits for; kept as a knob
    c->cg_edge_first = knob("FSB_CG_EDGE_FIRST", 0);
    c->cg_flags = (knob("FSB_CG_SERP", 1) ? 1 : 0) | (knob("FSB_CG_XHINT", 0) ? 2 : 0) |
                  (knob("FSB_CG_PREFETCH", 1) ? 4 : 0) | (knob("FSB_CG_PHINT", 0) ? 8 : 0) |
                  (keep << 4) | (knob("FSB_CG_XDEFER", 1) ? 128 : 0) |
                  (knob("FSB_CG_DEBUG_NOTILES", 0) ? 256 : 0) | (knob("FSB_CG_DEBUG_NOFAST", 0) ? 512 : 0) |
                  // sharded one-sweep solve: bit 11 -- GPU-scope instead of system-scope fence after the
                  // mailbox poll, bit 12 -- the same for the fence in front of the post.  Default: GPU
                  // scope.  Measured on 2 B200 (bare reduction, sweeps without tiles): 27.8 us per
                  // reduction with both at system scope, 14.8 with the poll fence at GPU scope, 13.6 with
                  // both (one GPU: 7.8).  Why it is enough: the boundary rows and the mailbox entry are
                  // stored into THIS GPU's memory by the peer, which fences at system scope between the
                  // two (every CTA that stored peer rows does, before it arrives at its own barrier); what
                  // reads them here are TMA loads issued after the entry was seen, through the L2 that
                  // received them.  FSB_CG_POLL_FENCE_SYS=1 / FSB_CG_POST_FENCE_SYS=1 restore the
                  // system-scope fences.
                  (knob("FSB_CG_POLL_FENCE_SYS", 0) ? 0 : 2048) | (knob("FSB_CG_POST_FENCE_SYS", 0) ? 0 : 4096) |
                  (knob("FSB_CG_DEBUG_TIMES", 0) ? 8192 : 0) | (knob("FSB_CG_ROTATE", 0) ? 16384 : 0);
  }

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FSB_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fsb_fail(c, FSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  if (c->cg_one)
  {
    c->cg_tile_rows = th; // its tensor maps are made by fsb_cg_one.cu
    return FSB_OK;
  }
  CUtensorMap halo_r, halo_p[2], inner_x, inner_r, code, inner_p[2];
  FSB_TRY(make_map(c, encode, &inner_p[0], c->cg_p[0], true, kTileW, th));
  FSB_TRY(make_map(c, encode, &inner_p[1], c->cg_p[1], true, kTileW, th));
  FSB_TRY(make_map(c, encode, &halo_r, c->cg_r, true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &halo_p[0], c->cg_p[0], true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &halo_p[1], c->cg_p[1], true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &inner_x, c->cg_x, true, kTileW, th));
  FSB_TRY(make_map(c, encode, &inner_r, c->cg_r, true, kTileW, th));
  FSB_TRY(make_map(c, encode, &code, c->cg_code, false, kCodeW, th + 2));
  {
    SolveMaps* m = reinterpret_cast<SolveMaps*>(c->cg_maps_fused);
    memset(m, 0, sizeof(SolveMaps));
    m->halo_r = halo_r;
    m->halo_p[0] = halo_p[0];
    m->halo_p[1] = halo_p[1];
    m->inner_x = inner_x;
    m->inner_r = inner_r;
    m->code = code;
    m->inner_p[0] = inner_p[0];
    m->inner_p[1] = inner_p[1];
  }
  for (int cur = 0; cur < 2; ++cur)
  {
    CgMaps* d = reinterpret_cast<CgMaps*>(c->cg_maps_dir[cur]);
    CgMaps* u = reinterpret_cast<CgMaps*>(c->cg_maps_upd[cur]);
    memset(d, 0, sizeof(CgMaps));
    memset(u, 0, sizeof(CgMaps));
    d->halo_a = halo_r;          // direction: r and the OLD direction p[cur]
    d->halo_b = halo_p[cur];
    d->code = code;
    u->halo_a = halo_p[cur ^ 1]; // update: the NEW direction p[cur^1], x, r
    u->inner_a = inner_x;
    u->inner_b = inner_r;
    u->code = code;
  }
  c->cg_tile_rows = th;
  return FSB_OK;
}

// one CG iteration = two launches; `cur` selects the ping-pong direction buffer
int launch_iteration(fsb_ctx* c, const CgCoef& coef, int cur)
{
  const int th = c->cg_tile_rows;
  const int tiles_x = fsb_div_up(c->ld, kTileW);
  const ShardArgs& sh = c->shard;
  const int n_tiles = tiles_x * fsb_div_up(sh.row_hi - sh.row_lo, th);
  const CgMaps& md = *reinterpret_cast<const CgMaps*>(c->cg_maps_dir[cur]);
  const CgMaps& mu = *reinterpret_cast<const CgMaps*>(c->cg_maps_upd[cur]);
  // slab boundary rows are stored into the neighbours' copies of the same rows
  const bool south = sh.world > 1 && sh.rank > 0, north = sh.world > 1 && sh.rank < sh.world - 1;
  const size_t lo_off = (size_t)sh.row_lo * c->ld, hi_off = (size_t)(sh.row_hi - 1) * c->ld;
  float* p_lo = south ? c->peer_p[cur ^ 1][sh.rank - 1] + lo_off : nullptr;
  float* p_hi = north ? c->peer_p[cur ^ 1][sh.rank + 1] + hi_off : nullptr;
  float* r_lo = south ? c->peer_r[sh.rank - 1] + lo_off : nullptr;
  float* r_hi = north ? c->peer_r[sh.rank + 1] + hi_off : nullptr;
  // both kernels allow programmatic dependent launch: the next kernel's CTAs are scheduled and
  // run their set-up while this kernel's reduction tail finishes (griddepcontrol in the kernels)
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = c->cg_pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3((kNW + 1) * 32);
  cfg.stream = c->stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float* p_new = c->cg_p[cur ^ 1];
  cudaError_t e1, e2;
#define FSB_CG_LAUNCH(RPW)                                                                         \
  cfg.gridDim = dim3(c->cg_grid_dir);                                                              \
  cfg.dynamicSmemBytes = (size_t)c->cg_stages_dir * DirStage<kNW * RPW>::kBytes;                   \
  e1 = cudaLaunchKernelEx(&cfg, k_cg_direction<kNW, RPW>, md, p_new, c->ld, tiles_x, n_tiles,      \
                          c->cg_stages_dir, coef, c->scal, c->partials, sh, p_lo, p_hi, 0);        \
  cfg.gridDim = dim3(c->cg_grid_upd);                                                              \
  cfg.dynamicSmemBytes = (size_t)c->cg_stages_upd * UpdStage<kNW * RPW>::kBytes;                   \
  e2 = cudaLaunchKernelEx(&cfg, k_cg_update<kNW, RPW>, mu, c->cg_x, c->cg_r, c->ld, tiles_x,       \
                          n_tiles, c->cg_stages_upd, coef, c->scal, c->partials, sh, r_lo, r_hi,   \
                          c->cg_flags & 1)
  if (th == 32) { FSB_CG_LAUNCH(4); }
  else if (th == 16) { FSB_CG_LAUNCH(2); }
  else { FSB_CG_LAUNCH(1); }
#undef FSB_CG_LAUNCH
  FSB_CUDA(c, e1);
  FSB_CUDA(c, e2);
  return FSB_OK;
}

// L2 persistence for the residual: r is touched three times per iteration (read by both sweeps,
// written by the update), more than any other vector.  FSB_CG_PERSIST_MB > 0 sets that much of the
// L2 aside (cudaLimitPersistingL2CacheSize) and puts an access-policy window over this rank's rows
// of r on the context's stream for the duration of the solve.
int set_l2_window(fsb_ctx* c, bool on)
{
  if (c->cg_persist_mb <= 0) return FSB_OK;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof attr);
  if (on)
  {
    cudaDeviceProp prop;
    FSB_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    size_t carve = std::min((size_t)c->cg_persist_mb << 20, (size_t)prop.persistingL2CacheMaxSize);
    if (carve == 0) return FSB_OK;
    FSB_CUDA(c, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    const ShardArgs& sh = c->shard;
    size_t bytes = (size_t)(sh.row_hi - sh.row_lo) * c->ld * sizeof(float);
    bytes = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
    attr.accessPolicyWindow.base_ptr = c->cg_r + (size_t)sh.row_lo * c->ld;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp =
        c->cg_persist_miss_normal ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
    if (getenv("FSB_CG_VERBOSE"))
      fprintf(stderr, "[fsb] L2 %d MB, persisting max %d MB, window max %d MB; carve %zu MB over %zu MB of r, hit ratio %.3f\n",
              prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20,
              prop.accessPolicyMaxWindowSize >> 20, carve >> 20, bytes >> 20,
              attr.accessPolicyWindow.hitRatio);
  }
  FSB_CUDA(c, cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  if (!on)
  {
    // give the set-aside back: it shrinks the L2 every other kernel sees
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  }
  return FSB_OK;
}

// the whole solve as one cooperative launch (k_cg_solve)
int launch_fused(fsb_ctx* c, const CgCoef& coef)
{
  const int th = c->cg_tile_rows;
  int tiles_x = fsb_div_up(c->ld, kTileW);
  const ShardArgs& sh = c->shard;
  int n_tiles = tiles_x * fsb_div_up(sh.row_hi - sh.row_lo, th);
  const SolveMaps& maps = *reinterpret_cast<const SolveMaps*>(c->cg_maps_fused);
  const bool south = sh.world > 1 && sh.rank > 0, north = sh.world > 1 && sh.rank < sh.world - 1;
  const size_t lo_off = (size_t)sh.row_lo * c->ld, hi_off = (size_t)(sh.row_hi - 1) * c->ld;
  SolvePush push;
  for (int k = 0; k < 2; ++k)
  {
    push.p_lo[k] = south ? c->peer_p[k][sh.rank - 1] + lo_off : nullptr;
    push.p_hi[k] = north ? c->peer_p[k][sh.rank + 1] + hi_off : nullptr;
  }
  push.r_lo = south ? c->peer_r[sh.rank - 1] + lo_off : nullptr;
  push.r_hi = north ? c->peer_r[sh.rank + 1] + hi_off : nullptr;
  int ld = c->ld, stages = c->cg_fused_stages, stage_bytes = c->cg_fused_stage_bytes;
  CgCoef cf = coef;
  int flags = c->cg_flags;
  void* args[] = {(void*)&maps, &c->cg_x, &c->cg_r, &c->cg_p[0], &c->cg_p[1], &ld, &tiles_x, &n_tiles,
                  &stages, &stage_bytes, &cf, &c->scal, &c->partials, (void*)&sh, &push, &flags};
  const dim3 grid(c->cg_grid_fused), block((kNW + 1) * 32);
  const size_t smem = (size_t)stages * stage_bytes;
  cudaError_t e;
  if (th == 32) e = cudaLaunchCooperativeKernel((void*)k_cg_solve<kNW, 4>, grid, block, args, smem, c->stream);
  else if (th == 16) e = cudaLaunchCooperativeKernel((void*)k_cg_solve<kNW, 2>, grid, block, args, smem, c->stream);
  else e = cudaLaunchCooperativeKernel((void*)k_cg_solve<kNW, 1>, grid, block, args, smem, c->stream);
  if (e != cudaSuccess)
    return fsb_fail(c, FSB_ERR_CUDA, "cooperative launch of the CG solve failed: %s",
                    cudaGetErrorString(e));
  c->launches += 1;
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

void fsb_cg_reconfigure(fsb_ctx* c)
{
  c->cg_tile_rows = 0;
  c->cg_one_th = 0;
  if (c->cg_graph) cudaGraphExecDestroy(c->cg_graph);
  c->cg_graph = nullptr;
  c->cg_graph_state = 0;
}

int fsb_k_pressure_solve(fsb_ctx* c, float density, float dt, bool fuse_dirichlet)
{
  const GridDims d = make_grid_dims(c->nx, c->ny, c->ld, c->dx, c->dy);
  const CgCoef coef = make_coef(c);

  // ---- build
  fsb_prof_begin(c, FSB_PROF_RHS);
  const int64_t total = (int64_t)c->ld * c->ny;
  // one resident wave of the set-up kernel (5 CTAs per SM at 48 registers): its grid-stride loop then has no
  // second, partly filled wave -- 0.114 -> 0.104 ms at 4096^2 against the former 8 per SM (4: 0.113, 6: 0.106,
  // 16: 0.113, 32: 0.128; tools/stage_knobs.py).  FSB_BUILD_BLOCKS_PER_SM overrides.  (Requesting the next
  // group's label rows early with prefetch.global.L1 made the pass slower: 0.106 -> 0.111 ms.)
  if (c->build_blocks_per_sm == 0)
  {
    int per_sm = 0;
    const cudaError_t e = (d.pow2 == 3)
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_build4<GridDimsP2>, 256, 0)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_build4<GridDims>, 256, 0);
    c->build_blocks_per_sm = (e == cudaSuccess && per_sm > 0) ? per_sm : 5;
  }
  const int build_blocks = (int)std::min<int64_t>(fsb_div_up(total, 256), c->sm_count * c->build_blocks_per_sm);
  FSB_TRY(configure_cg(c));
  const int need = std::max(std::max(std::max(3 * build_blocks, 3 * c->cg_grid_fused),
                                     std::max(c->cg_grid_dir, 2 * c->cg_grid_upd)),
                            fsb_cg_one_partials(c));
  if (need > c->partials_cap)
  {
    if (c->partials) cudaFree(c->partials);
    c->partials = nullptr;
    FSB_CUDA(c, cudaMalloc(&c->partials, sizeof(double) * need));
    c->partials_cap = need;
  }
  auto build = [&]() -> int {
    // active-tile list of this rank's slab (tiles of the configured height)
    const int th = c->cg_tile_rows;
    const int tiles_x = fsb_div_up(c->ld, kTileW);
    const int n_t = tiles_x * fsb_div_up(c->shard.row_hi - c->shard.row_lo, th);
    if (c->cg_skip_tiles && n_t > c->cg_tile_cap)
    {
      if (c->cg_tile_flags) cudaFree(c->cg_tile_flags);
      if (c->cg_tile_list) cudaFree(c->cg_tile_list);
      c->cg_tile_flags = c->cg_tile_list = nullptr;
      FSB_CUDA(c, cudaMalloc(&c->cg_tile_flags, sizeof(int) * n_t));
      FSB_CUDA(c, cudaMalloc(&c->cg_tile_list, sizeof(int) * n_t));
      c->cg_tile_cap = n_t;
    }
    if (c->stage_v1)
      k_cg_build<<<build_blocks, 256, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell, c->cg_code,
                                                      c->cg_x, c->cg_r, d, coef, c->scal, c->partials,
                                                      c->tol, c->max_iters);
    else if (d.pow2 == 3)
      k_cg_build4<GridDimsP2><<<build_blocks, 256, 0, c->stream>>>(
          fsb_uf(c), fsb_vf(c), c->cell, c->cg_code, c->cg_x, c->cg_r, as_pow2(d), coef, c->scal,
          c->partials, c->tol, c->max_iters);
    else
      k_cg_build4<GridDims><<<build_blocks, 256, 0, c->stream>>>(
          fsb_uf(c), fsb_vf(c), c->cell, c->cg_code, c->cg_x, c->cg_r, d, coef, c->scal, c->partials,
          c->tol, c->max_iters);
    FSB_LAUNCHED(c);
    if (c->cg_skip_tiles)
    {
      k_cg_tile_flags<<<n_t, 128, 0, c->stream>>>(c->cg_code, c->ld, tiles_x, th, c->shard.row_lo,
                                                  c->shard.row_hi, c->cg_tile_flags,
                                                  c->shard.world > 1 ? 1 : 0);
      FSB_LAUNCHED(c);
      k_cg_tile_compact<<<1, 1024, 0, c->stream>>>(c->cg_tile_flags, n_t, tiles_x, c->cg_tile_list,
                                                   c->scal, (c->shard.world > 1 && c->cg_edge_first) ? 1 : 0);
      FSB_LAUNCHED(c);
    }
    FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[0], c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                                c->stream));
    return FSB_OK;
  };
  FSB_TRY(build());
  fsb_prof_end(c, FSB_PROF_RHS);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->scal_h[0].n_liquid == 0) // :347-350: nothing touched, no swap
    return fuse_dirichlet ? fsb_k_enforce_dirichlet(c) : FSB_OK;

  // ---- iterate: chunks of kCheckEvery iterations; the host reads the device
  // scalars of chunk k while chunk k+1 is already queued, so the GPU never
  // waits for the poll.  Launches after convergence return immediately.
  fsb_prof_begin(c, FSB_PROF_CG);
  CgScalars fin = c->scal_h[0];
  c->cg_last_active_tiles = fin.tile_list ? fin.n_active_tiles : 0;
  if (getenv("FSB_CG_VERBOSE"))
    fprintf(stderr, "[fsb] solve: %d liquid cells, active tiles %d (list %s), tile rows %d, mode %s\n",
            fin.n_liquid, fin.n_active_tiles, fin.tile_list ? "on" : "off", c->cg_tile_rows,
            c->cg_one ? "one-sweep" : c->cg_fused ? "persistent" : "graph");
  c->last_solve_mg = false;
  if (!fin.done && c->precond == FSB_PRECOND_MULTIGRID && c->shard.world == 1)
  {
    // opt-in: multigrid-preconditioned CG (fsb_mg.cu); on breakdown the set-up is repeated
    // (x = 0, r = b, scalars) and the reference's Jacobi-preconditioned iteration runs instead
    int converged = 0;
    FSB_TRY(fsb_k_mg_solve(c, &converged));
    if (converged)
    {
      c->last_solve_mg = true;
      fin.done = 1;
    }
    else
    {
      // the abandoned iteration may have left non-finite values in the direction buffers,
      // which the sweeps expect to be zero outside the active tiles
      const size_t bytes = sizeof(float) * (size_t)c->ld * c->ny;
      FSB_CUDA(c, cudaMemsetAsync(c->cg_p[0], 0, bytes, c->stream));
      FSB_CUDA(c, cudaMemsetAsync(c->cg_p[1], 0, bytes, c->stream));
      FSB_TRY(build());
      FSB_CUDA(c, cudaStreamSynchronize(c->stream));
      fin = c->scal_h[0];
    }
  }
  c->last_solve_one = false;
  if (!fin.done && c->cg_one)
  {
    // one persistent kernel, one sweep per iteration
    FSB_TRY(fsb_k_cg_one_solve(c, coef));
    FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[0], c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                                c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    fin = c->scal_h[0];
    c->last_solve_one = true;
  }
  else if (!fin.done && c->cg_fused)
  {
    // one persistent kernel runs the loop to completion; no polling
    FSB_TRY(set_l2_window(c, true));
    FSB_TRY(launch_fused(c, coef));
    FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[0], c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                                c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    FSB_TRY(set_l2_window(c, false));
    fin = c->scal_h[0];
  }
  else if (!fin.done)
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
  if (c->shard.world > 1)
  {
    // every rank needs the whole pressure field for the (replicated) velocity patch
    const ShardArgs& sh = c->shard;
    const int64_t first = (int64_t)sh.row_lo * c->ld;
    const int64_t count4 = (int64_t)(sh.row_hi - sh.row_lo) * c->ld / 4;
    if (count4 > 0)
    {
      k_shard_scatter_rows<<<c->sm_count * 2, 256, 0, c->stream>>>(c->cg_x, c->peer_x_dev,
                                                                   sh.world - 1, first, count4);
      FSB_LAUNCHED(c);
    }
    k_shard_post_barrier<<<1, 32, 0, c->stream>>>(sh, c->scal);
    FSB_LAUNCHED(c);
    k_cg_combine<2><<<1, 32, 0, c->stream>>>(sh, c->scal);
    FSB_LAUNCHED(c);
    FSB_CUDA(c, cudaMemcpyAsync(&c->scal_h[0], c->scal, sizeof(CgScalars), cudaMemcpyDeviceToHost,
                                c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->scal_h[0].comm_error || fin.comm_error)
    {
      FSB_CUDA(c, cudaMemsetAsync(&c->scal->comm_error, 0, sizeof(int), c->stream));
      const CgScalars& q = c->scal_h[0];
      return fsb_fail(c, FSB_ERR_COMM,
                      "a peer rank did not answer within the mailbox time-out (rank %d of %d, iteration %d; "
                      "gave up in reduction %llu on rank/word %llu.%llu, expected tag %llu, saw word %llx)",
                      c->shard.rank, c->shard.world, fin.iter, q.comm_diag[0], q.comm_diag[1] / 16,
                      q.comm_diag[1] % 16, q.comm_diag[2], q.comm_diag[3]);
    }
  }
  fsb_prof_end(c, FSB_PROF_CG);
  if (c->cg_skip_tiles && !c->cg_one)
  {
    // skipped tiles rely on the direction buffers being zero there: leave this rank's OWN rows zero
    // for the next solve (whose liquid region differs).  The ghost rows next to the slab belong to
    // the neighbours (their boundary tiles are always active and rewrite them in every sweep), so
    // they are not touched here: a peer may already be storing into them for its next solve.
    const int lo = c->shard.world > 1 ? c->shard.row_lo : 0;
    const int hi = c->shard.world > 1 ? c->shard.row_hi : c->ny;
    const size_t off = (size_t)lo * c->ld, bytes = sizeof(float) * (size_t)(hi - lo) * c->ld;
    FSB_CUDA(c, cudaMemsetAsync(c->cg_p[0] + off, 0, bytes, c->stream));
    FSB_CUDA(c, cudaMemsetAsync(c->cg_p[1] + off, 0, bytes, c->stream));
  }
  if (!c->last_solve_mg)
  {
    c->iters = fin.iter;
    c->err = (fin.rhs2 == 0.0 || (float)fin.rhs2 == 0.0f)
                 ? 0.0f
                 : std::sqrt((float)fin.r2 / (float)fin.rhs2);
  }
  c->pressure_valid = true;

  // ---- patch + swap
  fsb_prof_begin(c, FSB_PROF_PATCH);
  const dim3 grid4(fsb_div_up(c->ld, 1024), c->ny);
  if (c->stage_v1)
    k_pressure_patch<<<dim3(fsb_div_up(c->ld, 256), c->ny), 256, 0, c->stream>>>(
        fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), c->cg_x, c->cg_code, d, dt, density);
  else
  {
#define FSB_PATCH(DIR, D, d_)                                                                       \
  k_pressure_patch4<DIR, D><<<grid4, 256, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), \
                                                          c->cg_x, c->cell, d_, dt, density)
    if (d.pow2 == 3) { if (fuse_dirichlet) FSB_PATCH(true, GridDimsP2, as_pow2(d)); else FSB_PATCH(false, GridDimsP2, as_pow2(d)); }
    else { if (fuse_dirichlet) FSB_PATCH(true, GridDims, d); else FSB_PATCH(false, GridDims, d); }
#undef FSB_PATCH
  }
  FSB_LAUNCHED(c);
  c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:482
  fsb_prof_end(c, FSB_PROF_PATCH);
  if (fuse_dirichlet && c->stage_v1) return fsb_k_enforce_dirichlet(c);
  return FSB_OK;
}
