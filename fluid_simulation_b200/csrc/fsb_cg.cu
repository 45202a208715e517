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
#include <cstring>

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
// Both iteration kernels are TMA-fed, warp-specialised, persistent kernels:
//   * one CTA per SM; TH consumer warps + 1 producer warp; tiles of kTileW = 128
//     columns x TH rows; consumer warp w owns tile row w, lane l its columns
//     4l..4l+3 (one 16-byte shared-memory access).
//   * the producer's elected lane walks the CTA's tile list and issues
//     cp.async.bulk.tensor (TMA) box loads into a kStages-deep shared-memory
//     ring, completion on full[stage] (mbarrier, expect_tx); consumers release a
//     stage by arriving on empty[stage].  Boxes include the one-cell halo
//     (fp32: 136 x (TH+2) starting at (c0-4, j0-1); code bytes: 160 x (TH+2)
//     starting at (c0-16, j0-1)); TMA zero-fills everything outside the grid,
//     so the kernels have no bounds logic on the load side and the pipeline
//     keeps (kStages-1) tiles of loads in flight per SM regardless of what the
//     consumer warps are doing.
//   * per-cell coefficients come from a shared-memory table of float4
//     {inverse diagonal, diagonal, off-diagonal, 0} indexed by the stencil code
//     (code 0 -> all zero, so masked cells come out exactly 0 without branches);
//     one LDS.128 per cell.  (An indexed kernel-parameter array compiles to
//     indexed LDC on the XU pipe: measured 78 % XU-bound, profiles/r01b.)
//   * arithmetic uses explicit FMA: the CG is held to the solver tolerance and
//     comparable iteration counts, not to Eigen's rounding.
constexpr int kTileW = 128;
constexpr int kHaloW = kTileW + 8;  // fp32 halo box width: columns c0-4 .. c0+131
constexpr int kCodeW = kTileW + 32; // code halo box width: columns c0-16 .. c0+143
constexpr int kStages = 4;

__host__ __device__ constexpr int align128(int x) { return (x + 127) / 128 * 128; }

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

struct CgMaps
{
  CUtensorMap halo_a; // fp32, box kHaloW x (TH+2)
  CUtensorMap halo_b; // fp32, box kHaloW x (TH+2)
  CUtensorMap inner_a; // fp32, box kTileW x TH
  CUtensorMap inner_b; // fp32, box kTileW x TH
  CUtensorMap code;    // u8,   box kCodeW x (TH+2)
};

// ---- PTX wrappers: mbarrier + TMA (sm_90+ forms, SASS: SYNCS / UTMALDG)
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void consumer_sync(int n_threads) // named barrier 1: consumer warps only
{
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// coefficient table: [code] -> {inv diag, diag, off, 0}; code 0 (not liquid) -> zeros
__device__ __forceinline__ void load_lut(float4* lut, const CgCoef& coef)
{
  if (threadIdx.x < 8)
  {
    const int t = threadIdx.x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 1 && t <= 5) v = make_float4(coef.invdiag[t - 1], coef.diag[t - 1], coef.off, 0.f);
    lut[t] = v;
  }
}

// q = A p for the four cells of one lane.  `prow` points at the lane's first
// cell in the staged halo tile of p (row pitch kHaloW); pc is that float4.
__device__ __forceinline__ float4 apply_a4(const float* __restrict__ prow, unsigned lane,
                                           const float4 pc, const float4 k0, const float4 k1,
                                           const float4 k2, const float4 k3)
{
  const float4 s4 = *reinterpret_cast<const float4*>(prow - kHaloW);
  const float4 n4 = *reinterpret_cast<const float4*>(prow + kHaloW);
  float w = __shfl_up_sync(0xffffffffu, pc.w, 1);
  float e = __shfl_down_sync(0xffffffffu, pc.x, 1);
  if (lane == 0) w = prow[-1];
  if (lane == 31) e = prow[4];
  float4 q;
  q.x = fmaf(k0.y, pc.x, k0.z * ((w + pc.y) + (s4.x + n4.x)));
  q.y = fmaf(k1.y, pc.y, k1.z * ((pc.x + pc.z) + (s4.y + n4.y)));
  q.z = fmaf(k2.y, pc.z, k2.z * ((pc.y + pc.w) + (s4.z + n4.z)));
  q.w = fmaf(k3.y, pc.w, k3.z * ((pc.z + e) + (s4.w + n4.w)));
  return q;
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b)
{
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// ------------------------------------------------- direction + p.Ap dot --
// p_new = z + beta p_old on the tile AND on its one-cell halo ring (recomputed
// from the staged r, p_old, code, so no other CTA's output is needed), written
// over p_old in the stage buffer; q = A p_new from shared memory / shuffles,
// kept in registers for the dot product only.  HBM traffic 13 B per cell:
// r, p_old, code in (TMA), p_new out (STG.128; ping-pong with p_old because
// other CTAs still read the old halo).
template <int TH>
__global__ void __launch_bounds__((TH + 1) * 32, 1)
k_cg_direction(const __grid_constant__ CgMaps maps, float* __restrict__ p_new, int ld, int ny,
               int tiles_x, int n_tiles, const CgCoef coef, CgScalars* __restrict__ s,
               double* __restrict__ partials)
{
  if (s->done) return;
  using St = DirStage<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  __shared__ float4 lut[8];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool first = (s->iter == 0);
  const float beta = first ? 0.0f : s->beta;

  load_lut(lut, coef);
  if (threadIdx.x == 0)
  {
    for (int k = 0; k < kStages; ++k)
    {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], TH);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == TH)
  {
    // ---- producer
    if (lane == 0)
    {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it)
      {
        const int st = it % kStages;
        if (it >= kStages) mbar_wait(&empty[st], ((it / kStages) - 1) & 1);
        const int c0 = (tile % tiles_x) * kTileW;
        const int j0 = (tile / tiles_x) * TH;
        unsigned char* base = smem + st * St::kBytes;
        mbar_expect_tx(&full[st], first ? St::kTx - St::kF32 : St::kTx);
        tma_load_2d(base + St::oR, &maps.halo_a, c0 - 4, j0 - 1, &full[st]);
        if (!first) tma_load_2d(base + St::oP, &maps.halo_b, c0 - 4, j0 - 1, &full[st]);
        tma_load_2d(base + St::oC, &maps.code, c0 - 16, j0 - 1, &full[st]);
      }
    }
    return;
  }

  // ---- consumers
  double acc = 0.0;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it)
  {
    const int st = it % kStages;
    const int c0 = (tile % tiles_x) * kTileW;
    const int j0 = (tile / tiles_x) * TH;
    unsigned char* base = smem + st * St::kBytes;
    const float* sr = reinterpret_cast<const float*>(base + St::oR);
    float* sp = reinterpret_cast<float*>(base + St::oP);
    const unsigned char* sc = base + St::oC;
    mbar_wait(&full[st], (it / kStages) & 1);

    // direction on the owned row (stage row warp+1) ...
    const int row = (int)warp + 1;
    const int fo = row * kHaloW + 4 + (int)lane * 4;
    const uint32_t c4 = *reinterpret_cast<const uint32_t*>(sc + row * kCodeW + 16 + lane * 4);
    const float4 k0 = lut[c4 & 0xff], k1 = lut[(c4 >> 8) & 0xff], k2 = lut[(c4 >> 16) & 0xff],
                 k3 = lut[c4 >> 24];
    const float4 r4 = *reinterpret_cast<const float4*>(sr + fo);
    float4 pn;
    if (first)
    {
      pn = make_float4(k0.x * r4.x, k1.x * r4.y, k2.x * r4.z, k3.x * r4.w);
    }
    else
    {
      const float4 p4 = *reinterpret_cast<const float4*>(sp + fo);
      pn.x = fmaf(beta, p4.x, k0.x * r4.x);
      pn.y = fmaf(beta, p4.y, k1.x * r4.y);
      pn.z = fmaf(beta, p4.z, k2.x * r4.z);
      pn.w = fmaf(beta, p4.w, k3.x * r4.w);
    }
    *reinterpret_cast<float4*>(sp + fo) = pn;
    // ... on its west / east halo cells (lanes 0 and 31) ...
    if (lane == 0 || lane == 31)
    {
      const int ho = row * kHaloW + (lane == 0 ? 3 : 4 + kTileW);
      const float inv = lut[sc[row * kCodeW + (lane == 0 ? 15 : 16 + kTileW)]].x;
      sp[ho] = first ? inv * sr[ho] : fmaf(beta, sp[ho], inv * sr[ho]);
    }
    // ... and on the south / north halo rows (warps 0 and 1)
    if (warp < 2)
    {
      const int hrow = (warp == 0) ? 0 : TH + 1;
      const int ho = hrow * kHaloW + 4 + (int)lane * 4;
      const uint32_t h4 = *reinterpret_cast<const uint32_t*>(sc + hrow * kCodeW + 16 + lane * 4);
      const float4 hr = *reinterpret_cast<const float4*>(sr + ho);
      float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!first) hp = *reinterpret_cast<const float4*>(sp + ho);
      float4 hn;
      hn.x = fmaf(beta, hp.x, lut[h4 & 0xff].x * hr.x);
      hn.y = fmaf(beta, hp.y, lut[(h4 >> 8) & 0xff].x * hr.y);
      hn.z = fmaf(beta, hp.z, lut[(h4 >> 16) & 0xff].x * hr.z);
      hn.w = fmaf(beta, hp.w, lut[h4 >> 24].x * hr.w);
      *reinterpret_cast<float4*>(sp + ho) = hn;
    }
    consumer_sync(TH * 32);

    const float4 q = apply_a4(sp + fo, lane, pn, k0, k1, k2, k3);
    acc += (double)dot4(pn, q);
    const int j = j0 + (int)warp, ci = c0 + (int)lane * 4;
    if (j < ny && ci < ld) *reinterpret_cast<float4*>(p_new + (size_t)j * ld + ci) = pn;

    // the stage was written through the generic proxy; the next TMA load into
    // it goes through the async proxy
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

  // ---- reduction over the consumer warps, then over the CTAs (last one folds)
  __shared__ double s_part[32];
  __shared__ bool s_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) s_part[warp] = acc;
  consumer_sync(TH * 32);
  if (warp == 0)
  {
    double v = (lane < TH) ? s_part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0)
    {
      partials[blockIdx.x] = v;
      __threadfence();
      s_last = (atomicAdd(&s->ticket[1], 1u) == gridDim.x - 1);
    }
  }
  consumer_sync(TH * 32);
  if (s_last)
  {
    __threadfence();
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += TH * 32)
      v += reinterpret_cast<const volatile double*>(partials)[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp] = v;
    consumer_sync(TH * 32);
    if (threadIdx.x == 0)
    {
      double t = 0.0;
      for (int k = 0; k < TH; ++k) t += s_part[k];
      s->pq = t;
      s->ticket[1] = 0;
    }
  }
}

// ------------------------------------------------------------ the update --
// alpha = absNew / p.Ap; q = A p RE-COMPUTED from the staged p tile (q is
// never stored); x += alpha p; r -= alpha q; partial |r|^2 and r.z.
// HBM traffic 21 B per cell: p, code, x, r in (TMA), x, r out (STG.128).
// No CTA-level barrier inside the tile loop: a warp only needs the TMA data.
template <int TH>
__global__ void __launch_bounds__((TH + 1) * 32, 1)
k_cg_update(const __grid_constant__ CgMaps maps, float* __restrict__ x, float* __restrict__ r,
            int ld, int ny, int tiles_x, int n_tiles, const CgCoef coef,
            CgScalars* __restrict__ s, double* __restrict__ partials)
{
  if (s->done) return;
  using St = UpdStage<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  __shared__ float4 lut[8];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float alpha = s->abs_new / (float)s->pq; // Eigen: alpha = absNew / p.dot(tmp)
  const float nalpha = -alpha;

  load_lut(lut, coef);
  if (threadIdx.x == 0)
  {
    for (int k = 0; k < kStages; ++k)
    {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], TH);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == TH)
  {
    if (lane == 0)
    {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it)
      {
        const int st = it % kStages;
        if (it >= kStages) mbar_wait(&empty[st], ((it / kStages) - 1) & 1);
        const int c0 = (tile % tiles_x) * kTileW;
        const int j0 = (tile / tiles_x) * TH;
        unsigned char* base = smem + st * St::kBytes;
        mbar_expect_tx(&full[st], St::kTx);
        tma_load_2d(base + St::oP, &maps.halo_a, c0 - 4, j0 - 1, &full[st]);
        tma_load_2d(base + St::oX, &maps.inner_a, c0, j0, &full[st]);
        tma_load_2d(base + St::oR, &maps.inner_b, c0, j0, &full[st]);
        tma_load_2d(base + St::oC, &maps.code, c0 - 16, j0 - 1, &full[st]);
      }
    }
    return;
  }

  double acc_r2 = 0.0, acc_rz = 0.0;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it)
  {
    const int st = it % kStages;
    const int c0 = (tile % tiles_x) * kTileW;
    const int j0 = (tile / tiles_x) * TH;
    unsigned char* base = smem + st * St::kBytes;
    const float* sp = reinterpret_cast<const float*>(base + St::oP);
    const float* sx = reinterpret_cast<const float*>(base + St::oX);
    const float* sr = reinterpret_cast<const float*>(base + St::oR);
    const unsigned char* sc = base + St::oC;
    mbar_wait(&full[st], (it / kStages) & 1);

    const int row = (int)warp + 1;
    const int fo = row * kHaloW + 4 + (int)lane * 4;
    const uint32_t c4 = *reinterpret_cast<const uint32_t*>(sc + row * kCodeW + 16 + lane * 4);
    const float4 pc = *reinterpret_cast<const float4*>(sp + fo);
    const float4 k0 = lut[c4 & 0xff], k1 = lut[(c4 >> 8) & 0xff], k2 = lut[(c4 >> 16) & 0xff],
                 k3 = lut[c4 >> 24];
    const float4 q = apply_a4(sp + fo, lane, pc, k0, k1, k2, k3);
    float4 xo = *reinterpret_cast<const float4*>(sx + warp * kTileW + lane * 4);
    float4 ro = *reinterpret_cast<const float4*>(sr + warp * kTileW + lane * 4);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]); // everything this warp needs is in registers

    if (c4 != 0) // four non-liquid cells: x and r stay exactly zero, nothing to write
    {
      xo.x = fmaf(alpha, pc.x, xo.x); ro.x = fmaf(nalpha, q.x, ro.x);
      xo.y = fmaf(alpha, pc.y, xo.y); ro.y = fmaf(nalpha, q.y, ro.y);
      xo.z = fmaf(alpha, pc.z, xo.z); ro.z = fmaf(nalpha, q.z, ro.z);
      xo.w = fmaf(alpha, pc.w, xo.w); ro.w = fmaf(nalpha, q.w, ro.w);
      const int j = j0 + (int)warp, ci = c0 + (int)lane * 4; // c4 != 0 implies inside the grid
      const size_t o = (size_t)j * ld + ci;
      *reinterpret_cast<float4*>(x + o) = xo;
      *reinterpret_cast<float4*>(r + o) = ro;
      const float4 z = make_float4(k0.x * ro.x, k1.x * ro.y, k2.x * ro.z, k3.x * ro.w);
      acc_r2 += (double)dot4(ro, ro);
      acc_rz += (double)dot4(ro, z);
    }
  }

  __shared__ double s_part[2][32];
  __shared__ bool s_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    acc_r2 += __shfl_down_sync(0xffffffffu, acc_r2, o);
    acc_rz += __shfl_down_sync(0xffffffffu, acc_rz, o);
  }
  if (lane == 0)
  {
    s_part[0][warp] = acc_r2;
    s_part[1][warp] = acc_rz;
  }
  consumer_sync(TH * 32);
  if (warp == 0)
  {
    double v0 = (lane < TH) ? s_part[0][lane] : 0.0;
    double v1 = (lane < TH) ? s_part[1][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      v0 += __shfl_down_sync(0xffffffffu, v0, o);
      v1 += __shfl_down_sync(0xffffffffu, v1, o);
    }
    if (lane == 0)
    {
      partials[blockIdx.x] = v0;
      partials[gridDim.x + blockIdx.x] = v1;
      __threadfence();
      s_last = (atomicAdd(&s->ticket[2], 1u) == gridDim.x - 1);
    }
  }
  consumer_sync(TH * 32);
  if (s_last)
  {
    __threadfence();
    const volatile double* part = partials;
    double v0 = 0.0, v1 = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += TH * 32)
    {
      v0 += part[k];
      v1 += part[gridDim.x + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      v0 += __shfl_down_sync(0xffffffffu, v0, o);
      v1 += __shfl_down_sync(0xffffffffu, v1, o);
    }
    if (lane == 0)
    {
      s_part[0][warp] = v0;
      s_part[1][warp] = v1;
    }
    consumer_sync(TH * 32);
    if (threadIdx.x == 0)
    {
      double tr2 = 0.0, trz = 0.0;
      for (int k = 0; k < TH; ++k)
      {
        tr2 += s_part[0][k];
        trz += s_part[1][k];
      }
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

// Tile height = consumer warps per CTA: 16 rows when that still gives every SM
// several tiles, else 8.
int pick_tile_rows(const fsb_ctx* c)
{
  if (const char* e = getenv("FSB_CG_TILE_ROWS")) // tuning knob for profiling runs
  {
    const int th = atoi(e);
    if (th == 8 || th == 16) return th;
  }
  const int tiles_x = fsb_div_up(c->ld, kTileW);
  if ((int64_t)tiles_x * fsb_div_up(c->ny, 16) >= (int64_t)4 * c->sm_count) return 16;
  return 8;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tiled tensor map over a pitched grid; out-of-range box elements read as zero
int make_map(fsb_ctx* c, EncodeTiledFn encode, CUtensorMap* map, void* base, bool is_f32,
             int box_w, int box_h)
{
  const cuuint64_t esz = is_f32 ? 4 : 1;
  const cuuint64_t gdim[2] = {(cuuint64_t)c->ld, (cuuint64_t)c->ny};
  const cuuint64_t gstride[1] = {(cuuint64_t)c->ld * esz};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                            2, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fsb_fail(c, FSB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d box", (int)r,
                    box_w, box_h);
  return FSB_OK;
}

template <int TH>
int configure_kernels(fsb_ctx* c, int64_t n_tiles)
{
  const int threads = (TH + 1) * 32;
  const int smem_dir = kStages * DirStage<TH>::kBytes, smem_upd = kStages * UpdStage<TH>::kBytes;
  FSB_CUDA(c, cudaFuncSetAttribute(k_cg_direction<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   smem_dir));
  FSB_CUDA(c, cudaFuncSetAttribute(k_cg_update<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   smem_upd));
  int occ_dir = 1, occ_upd = 1;
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, k_cg_direction<TH>, threads,
                                                            smem_dir));
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, k_cg_update<TH>, threads,
                                                            smem_upd));
  int cap = 2; // CTAs per SM worth keeping resident (each already keeps kStages-1 tiles in flight)
  if (const char* e = getenv("FSB_CG_CTAS_PER_SM")) cap = std::max(1, atoi(e));
  occ_dir = std::max(1, std::min(occ_dir, cap));
  occ_upd = std::max(1, std::min(occ_upd, cap));
  // persistent-style grids: exactly one resident wave, each CTA walks the tile list
  c->cg_grid_dir = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ_dir);
  c->cg_grid_upd = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ_upd);
  return FSB_OK;
}

int configure_cg(fsb_ctx* c)
{
  if (c->cg_tile_rows != 0) return FSB_OK;
  const int th = pick_tile_rows(c);
  const int64_t n_tiles = (int64_t)fsb_div_up(c->ld, kTileW) * fsb_div_up(c->ny, th);
  if (th == 16) FSB_TRY(configure_kernels<16>(c, n_tiles));
  else FSB_TRY(configure_kernels<8>(c, n_tiles));

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FSB_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fsb_fail(c, FSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  CUtensorMap halo_r, halo_p[2], inner_x, inner_r, code;
  FSB_TRY(make_map(c, encode, &halo_r, c->cg_r, true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &halo_p[0], c->cg_p[0], true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &halo_p[1], c->cg_p[1], true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &inner_x, c->cg_x, true, kTileW, th));
  FSB_TRY(make_map(c, encode, &inner_r, c->cg_r, true, kTileW, th));
  FSB_TRY(make_map(c, encode, &code, c->cg_code, false, kCodeW, th + 2));
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
  const int n_tiles = tiles_x * fsb_div_up(c->ny, th);
  const CgMaps& md = *reinterpret_cast<const CgMaps*>(c->cg_maps_dir[cur]);
  const CgMaps& mu = *reinterpret_cast<const CgMaps*>(c->cg_maps_upd[cur]);
#define FSB_CG_LAUNCH(TH)                                                                          \
  k_cg_direction<TH><<<c->cg_grid_dir, (TH + 1) * 32, kStages * DirStage<TH>::kBytes, c->stream>>>( \
      md, c->cg_p[cur ^ 1], c->ld, c->ny, tiles_x, n_tiles, coef, c->scal, c->partials);           \
  k_cg_update<TH><<<c->cg_grid_upd, (TH + 1) * 32, kStages * UpdStage<TH>::kBytes, c->stream>>>(   \
      mu, c->cg_x, c->cg_r, c->ld, c->ny, tiles_x, n_tiles, coef, c->scal, c->partials)
  if (th == 16) { FSB_CG_LAUNCH(16); }
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
  const GridDims d = make_grid_dims(c->nx, c->ny, c->ld, c->dx, c->dy);
  const CgCoef coef = make_coef(c);

  // ---- build
  fsb_prof_begin(c, FSB_PROF_RHS);
  const int64_t total = (int64_t)c->ld * c->ny;
  const int build_blocks = (int)std::min<int64_t>(fsb_div_up(total, 256), c->sm_count * 8);
  FSB_TRY(configure_cg(c));
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
