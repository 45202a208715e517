// Particle-side stages: cell-binned stable sort, particle-to-grid transfer as
// an atomics-free gather over the sorted particles, grid-to-particle transfer
// fused with the PIC/FLIP blend and the advection, RK3 particle tracing.
#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"

namespace {

constexpr int kBlock = 256;

inline GridDims dims(const fsb_ctx* c) { return make_grid_dims(c->nx, c->ny, c->ld, c->dx, c->dy); }
// the P2G accumulators carry the memory pool's deltas (src/FluidSolver.cpp:13-16)
inline GridDims pool_dims(const fsb_ctx* c)
{
  return make_grid_dims(c->nx, c->ny, c->ld, c->pool_dx, c->pool_dy);
}

// ------------------------------------------------------------------ sort --
// Key = the particle's interpolation base cell (int)(pos/delta), index-clamped
// (formula 3 of SURVEY.md A.4), dense index ci + cj*nx.  Runs of equal keys
// inside a warp are aggregated so that only the run head touches the counter.
//
// MARK: the same pass also performs FluidDomain::classifyCells' marking (src/FluidDomain.cpp:
// 157-167, the formula of k_mark_liquid: (int)((pos / length) * size), clamped, border skipped),
// so that a step reads the particle set once for both.
//
// PER particles per thread (block-strided, every access coalesced): a particle is a chain of dependent
// round trips -- its record, then the counter atomic whose return value is its rank -- and with one
// particle per thread the pass is bound by that chain times the number of thread waves.  The records of
// all PER particles are requested first, then all atomics are issued, then everything is stored.
template <bool MARK, class D, int PER>
__global__ void __launch_bounds__(256)
k_sort_count(const float4* __restrict__ part, int64_t n, const D d,
             int* __restrict__ count, int* __restrict__ key_out,
             int* __restrict__ rank_out, uint8_t* __restrict__ cell,
             const GridDims gd, const D glen)
{
  const int64_t k0 = (int64_t)blockIdx.x * (blockDim.x * PER) + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  float4 p[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    p[q] = (k < n) ? part[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int key[PER], base[PER], start[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const bool live = k0 + (int64_t)q * blockDim.x < n;
    key[q] = -1 - (int)lane; // distinct per lane, never equal to a real key
    if (live)
    {
      const int ci = clampi((int)div_dx(d, p[q].x), 0, d.nx - 1);
      const int cj = clampi((int)div_dy(d, p[q].y), 0, d.ny - 1);
      key[q] = ci + cj * d.nx;
      if (MARK)
      {
        const int x = clampi((int)(div_dx(glen, p[q].x) * (float)gd.nx), 0, gd.nx - 1);
        const int y = clampi((int)(div_dy(glen, p[q].y) * (float)gd.ny), 0, gd.ny - 1);
        if (!(x == 0 || y == 0 || x == gd.nx - 1 || y == gd.ny - 1))
          cell[x + (size_t)y * gd.ld] = FSB_LIQUID;
      }
    }
    const int prev = __shfl_up_sync(0xffffffffu, key[q], 1);
    const bool head = (lane == 0) || (key[q] != prev);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned below = heads & (0xffffffffu >> (31 - lane)); // heads at lanes <= mine
    start[q] = 31 - __clz(below);
    const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int len = above ? __ffs(above) : (32 - (int)lane); // run length seen from a head
    base[q] = 0;
    if (head && live) base[q] = atomicAdd(count + key[q], len);
  }
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    const int b = __shfl_sync(0xffffffffu, base[q], start[q]);
    if (k < n)
    {
      key_out[k] = key[q];
      rank_out[k] = b + ((int)lane - start[q]);
    }
  }
}

// exclusive scan of count[0..m) into start[0..m], three small kernels
constexpr int kScanBlock = 256;
constexpr int kScanItems = 16; // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ int block_scan_excl(int v, int* total)
{
  // exclusive scan of one int per thread over the block (blockDim.x a multiple of 32)
  __shared__ int s_warp[33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads(); // s_warp may still be read by a previous call
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0)
  {
    const int nw = blockDim.x >> 5;
    const int w = (lane < nw) ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;       // exclusive offset of warp `lane`
    if (lane == 31) s_warp[32] = wi; // block total (lanes >= nw contribute 0)
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + incl - v;
}

__global__ void k_scan_reduce(const int* __restrict__ count, int m, int* __restrict__ block_sum)
{
  const int base = blockIdx.x * kScanTile;
  int s = 0;
  for (int t = threadIdx.x; t < kScanTile; t += kScanBlock)
  {
    const int idx = base + t;
    if (idx < m) s += count[idx];
  }
  // block reduce
  __shared__ int s_w[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) s_w[wid] = s;
  __syncthreads();
  if (wid == 0)
  {
    s = (lane < (kScanBlock >> 5)) ? s_w[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) block_sum[blockIdx.x] = s;
  }
}

// single block: in-place exclusive scan of the block sums
__global__ void k_scan_blocks(int* __restrict__ block_sum, int nb)
{
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x)
  {
    const int idx = base + threadIdx.x;
    const int v = (idx < nb) ? block_sum[idx] : 0;
    int total;
    const int ex = block_scan_excl(v, &total);
    const int carry = s_carry;
    if (idx < nb) block_sum[idx] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
}

__global__ void k_scan_apply(const int* __restrict__ count, int m,
                             const int* __restrict__ block_sum, int* __restrict__ start)
{
  // each thread owns kScanItems consecutive entries
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int t = 0; t < kScanItems; ++t)
  {
    const int idx = base + t;
    v[t] = (idx < m) ? count[idx] : 0;
    s += v[t];
  }
  int total;
  int run = block_sum[blockIdx.x] + block_scan_excl(s, &total);
#pragma unroll
  for (int t = 0; t < kScanItems; ++t)
  {
    const int idx = base + t;
    if (idx < m) start[idx] = run;
    run += v[t];
  }
  // the grand total goes to start[m]
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanBlock - 1) start[m] = run;
}

__global__ void k_sort_place(const int* __restrict__ key, const int* __restrict__ rank, int64_t n,
                             const int* __restrict__ start, int* __restrict__ idx)
{
  // four particles per thread: the four dependent start[] look-ups are in flight together
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (k >= n) return;
  if (k + 4 <= n)
  {
    const int4 ky = *reinterpret_cast<const int4*>(key + k);
    const int4 rk = *reinterpret_cast<const int4*>(rank + k);
    const int s0 = __ldg(start + ky.x), s1 = __ldg(start + ky.y), s2 = __ldg(start + ky.z),
              s3 = __ldg(start + ky.w);
    idx[s0 + rk.x] = (int)k;
    idx[s1 + rk.y] = (int)k + 1;
    idx[s2 + rk.z] = (int)k + 2;
    idx[s3 + rk.w] = (int)k + 3;
    return;
  }
  for (int64_t q = k; q < n; ++q) idx[start[key[q]] + rank[q]] = (int)q;
}

// Make the order inside each cell canonical: ascending ORIGINAL index (the index the caller knows
// the particle by; the global id in a slab-partitioned run).  This turns the atomic ranking into a
// deterministic sort whose result does not depend on the order the particles were stored in
// before: every floating-point sum over a cell's particles is reproducible across runs, across
// state files and across particle partitions.
//
// PER cells per thread (block-strided): a cell is three dependent round trips (its bounds, its entries,
// their keys); the loads of all PER cells are issued level by level so that the chains overlap.
template <int PER>
__global__ void __launch_bounds__(256)
k_sort_canon(const int* __restrict__ start, int m, int* __restrict__ idx, const int* __restrict__ orig)
{
  const int c0 = blockIdx.x * (blockDim.x * PER) + threadIdx.x;
  int a[PER], n[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int c = c0 + q * (int)blockDim.x;
    a[q] = 0;
    n[q] = 0;
    if (c < m)
    {
      a[q] = start[c];
      n[q] = start[c + 1] - a[q];
    }
  }
  // the common case (2 .. 8 entries): all entries and their keys loaded at once (independent loads), a
  // fixed 19-exchange sorting network in registers, and a write-back only where the order changed
  int v[PER][8], key[PER][8];
#pragma unroll
  for (int q = 0; q < PER; ++q)
#pragma unroll
    for (int t = 0; t < 8; ++t) v[q][t] = (n[q] >= 2 && n[q] <= 8 && t < n[q]) ? idx[a[q] + t] : -1;
#pragma unroll
  for (int q = 0; q < PER; ++q)
#pragma unroll
    for (int t = 0; t < 8; ++t) key[q][t] = (v[q][t] >= 0) ? __ldg(orig + v[q][t]) : 0x7fffffff;
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    if (n[q] <= 1) continue;
    if (n[q] <= 8)
    {
      int w[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) w[t] = v[q][t];
#define FSB_CX(p, r)                                              \
  {                                                               \
    const bool sw = key[q][p] > key[q][r];                        \
    const int klo = sw ? key[q][r] : key[q][p];                   \
    const int khi = sw ? key[q][p] : key[q][r];                   \
    const int wlo = sw ? w[r] : w[p];                             \
    const int whi = sw ? w[p] : w[r];                             \
    key[q][p] = klo; key[q][r] = khi; w[p] = wlo; w[r] = whi;     \
  }
      FSB_CX(0, 1) FSB_CX(2, 3) FSB_CX(4, 5) FSB_CX(6, 7)
      FSB_CX(0, 2) FSB_CX(1, 3) FSB_CX(4, 6) FSB_CX(5, 7)
      FSB_CX(1, 2) FSB_CX(5, 6) FSB_CX(0, 4) FSB_CX(3, 7)
      FSB_CX(1, 5) FSB_CX(2, 6)
      FSB_CX(1, 4) FSB_CX(3, 6)
      FSB_CX(2, 4) FSB_CX(3, 5)
      FSB_CX(3, 4)
#undef FSB_CX
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t < n[q] && w[t] != v[q][t]) idx[a[q] + t] = w[t];
      continue;
    }
    const int lo = a[q], hi = a[q] + n[q];
    for (int s = lo + 1; s < hi; ++s)
    {
      const int vv = idx[s];
      const int kv = orig[vv];
      int t = s - 1;
      while (t >= lo && orig[idx[t]] > kv)
      {
        idx[t + 1] = idx[t];
        --t;
      }
      idx[t + 1] = vv;
    }
  }
}

template <int PER>
__global__ void __launch_bounds__(256)
k_sort_gather(const float4* __restrict__ src, const int* __restrict__ src_orig,
              const int* __restrict__ idx, int64_t n, float4* __restrict__ dst,
              int* __restrict__ dst_orig)
{
  // PER destinations per thread, block-strided: all permutation entries first, then all records, then the stores
  const int64_t k0 = (int64_t)blockIdx.x * (blockDim.x * PER) + threadIdx.x;
  int sidx[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    sidx[q] = (k < n) ? idx[k] : 0;
  }
  float4 v[PER];
  int o[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const bool live = k0 + (int64_t)q * blockDim.x < n;
    v[q] = live ? src[sidx[q]] : make_float4(0.f, 0.f, 0.f, 0.f);
    o[q] = live ? src_orig[sidx[q]] : 0;
  }
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    if (k < n)
    {
      dst[k] = v[q];
      dst_orig[k] = o[q];
    }
  }
}

// ------------------------------------------------------------------- P2G --
// src/FluidSolver.cpp:873-919 as a register-streaming, atomics-free,
// deterministic transfer over the cell-sorted particles.
//
// A warp owns a strip of 32 cell columns (lane = column; lanes 1..30 produce
// output, lanes 0 and 31 are the halo columns) and walks up a chunk of
// kP2gRows output rows.  In every cell row each lane visits ITS OWN cell's
// particles once, computes the reference's four splat targets and weights
// (include/Grid.h:152-184: truncation, fraction before clamping, independent
// clamps, y-split then x-split) and adds them into rolling per-lane node
// accumulators:  u targets lie in columns {ci, ci+1} and rows {c-1, c, c+1}
// of the particle's sort cell (ci, c), v targets in columns {ci-1, ci, ci+1}
// and rows {c, c+1}.  After cell row c the u nodes of row c-1 and the v nodes
// of row c are complete: the cross-column parts move one lane over by warp
// shuffle, the face is written as sum / weight where weight > 1e-6 (else the
// back buffer keeps its stale value, :902-915), and the accumulators rotate.
// Each particle is read 1.13 times (halo columns / rows); no shared memory,
// no atomics, and the summation order is fixed.
constexpr int kP2gRows = 32;
constexpr int kP2gCols = 30;

struct P2gAcc
{
  // u: [row slot 0..2 = c-1, c, c+1][column offset 0..1]; v: [row slot 0..1 = c, c+1][column offset -1..1]
  float us[3][2], uw[3][2];
  float vs[2][3], vw[2][3];
};

// generic (border / out-of-domain) accumulation: target node (it, jt) -> slot by comparison
__device__ __forceinline__ void p2g_add_u(P2gAcc& a, int di, int dj, float val, float w)
{
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (dj == r - 1 && di == q)
      {
        a.us[r][q] += val;
        a.uw[r][q] += w;
      }
}
__device__ __forceinline__ void p2g_add_v(P2gAcc& a, int di, int dj, float val, float w)
{
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (dj == r && di == q - 1)
      {
        a.vs[r][q] += val;
        a.vw[r][q] += w;
      }
}

template <class D>
__device__ __forceinline__ void p2g_particle(P2gAcc& a, const float4 p, const D& d, int ci, int c,
                                             float half_dx, float half_dy)
{
  // ---- u: splat (vel_x, 1) at (px, py - dy/2)
  {
    const float xd = div_dx(d, p.x);
    const float yd = div_dy(d, p.y - half_dy);
    const int bi = (int)xd, bj = (int)yd;
    const float fi = xd - (float)bi, fj = yd - (float)bj;
    const float v0 = (1.0f - fj) * p.z, v1 = fj * p.z;
    const float w0 = (1.0f - fj) * 1.0f, w1 = fj * 1.0f;
    const float s00 = (1.0f - fi) * v0, s10 = fi * v0, s01 = (1.0f - fi) * v1, s11 = fi * v1;
    const float t00 = (1.0f - fi) * w0, t10 = fi * w0, t01 = (1.0f - fi) * w1, t11 = fi * w1;
    const bool regular = bi == ci && bi + 1 < d.nx && bj >= 0 && bj + 1 < d.ny;
    if (regular && bj == c - 1)
    {
      a.us[0][0] += s00; a.uw[0][0] += t00; a.us[0][1] += s10; a.uw[0][1] += t10;
      a.us[1][0] += s01; a.uw[1][0] += t01; a.us[1][1] += s11; a.uw[1][1] += t11;
    }
    else if (regular && bj == c)
    {
      a.us[1][0] += s00; a.uw[1][0] += t00; a.us[1][1] += s10; a.uw[1][1] += t10;
      a.us[2][0] += s01; a.uw[2][0] += t01; a.us[2][1] += s11; a.uw[2][1] += t11;
    }
    else
    {
      const int i0 = clampi(bi, 0, d.nx - 1) - ci, i1 = clampi(bi + 1, 0, d.nx - 1) - ci;
      const int j0 = clampi(bj, 0, d.ny - 1) - c, j1 = clampi(bj + 1, 0, d.ny - 1) - c;
      p2g_add_u(a, i0, j0, s00, t00);
      p2g_add_u(a, i1, j0, s10, t10);
      p2g_add_u(a, i0, j1, s01, t01);
      p2g_add_u(a, i1, j1, s11, t11);
    }
  }
  // ---- v: splat (vel_y, 1) at (px - dx/2, py)
  {
    const float xd = div_dx(d, p.x - half_dx);
    const float yd = div_dy(d, p.y);
    const int bi = (int)xd, bj = (int)yd;
    const float fi = xd - (float)bi, fj = yd - (float)bj;
    const float v0 = (1.0f - fj) * p.w, v1 = fj * p.w;
    const float w0 = (1.0f - fj) * 1.0f, w1 = fj * 1.0f;
    const float s00 = (1.0f - fi) * v0, s10 = fi * v0, s01 = (1.0f - fi) * v1, s11 = fi * v1;
    const float t00 = (1.0f - fi) * w0, t10 = fi * w0, t01 = (1.0f - fi) * w1, t11 = fi * w1;
    const bool regular = bj == c && bj + 1 < d.ny && bi >= 0 && bi + 1 < d.nx;
    if (regular && bi == ci - 1)
    {
      a.vs[0][0] += s00; a.vw[0][0] += t00; a.vs[0][1] += s10; a.vw[0][1] += t10;
      a.vs[1][0] += s01; a.vw[1][0] += t01; a.vs[1][1] += s11; a.vw[1][1] += t11;
    }
    else if (regular && bi == ci)
    {
      a.vs[0][1] += s00; a.vw[0][1] += t00; a.vs[0][2] += s10; a.vw[0][2] += t10;
      a.vs[1][1] += s01; a.vw[1][1] += t01; a.vs[1][2] += s11; a.vw[1][2] += t11;
    }
    else
    {
      const int i0 = clampi(bi, 0, d.nx - 1) - ci, i1 = clampi(bi + 1, 0, d.nx - 1) - ci;
      const int j0 = clampi(bj, 0, d.ny - 1) - c, j1 = clampi(bj + 1, 0, d.ny - 1) - c;
      p2g_add_v(a, i0, j0, s00, t00);
      p2g_add_v(a, i1, j0, s10, t10);
      p2g_add_v(a, i0, j1, s01, t01);
      p2g_add_v(a, i1, j1, s11, t11);
    }
  }
}

template <class D, bool PIPE>
__global__ void __launch_bounds__(256)
k_p2g_stream(const float4* __restrict__ part, const int* __restrict__ cell_start,
             float* __restrict__ ub, float* __restrict__ vb, const D d, float half_dx,
             float half_dy, int strips_x, int n_warps, int row_lo, int row_hi)
{
  const int warp_id = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (warp_id >= n_warps) return; // whole warps leave together
  const int lane = threadIdx.x & 31;
  const int ci = (warp_id % strips_x) * kP2gCols - 1 + lane;
  // output rows [ja, jb) of this warp; a node's sum is built in the same order whatever row the
  // chunk starts at (cell rows j-1, j, j+1 in turn), so a row range gives the same bits as a full pass
  const int ja = row_lo + (warp_id / strips_x) * kP2gRows;
  const int jb = min(ja + kP2gRows, row_hi);
  const bool col_ok = ci >= 0 && ci < d.nx;
  const bool writer = col_ok && lane >= 1 && lane <= kP2gCols;

  P2gAcc a;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 2; ++q) a.us[r][q] = a.uw[r][q] = 0.0f;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) a.vs[r][q] = a.vw[r][q] = 0.0f;

  for (int c = ja - 1; c <= jb; ++c)
  {
    if (col_ok && c >= 0 && c < d.ny)
    {
      const int p0 = __ldg(cell_start + ci + c * d.nx);
      const int p1 = __ldg(cell_start + ci + c * d.nx + 1);
      if (PIPE)
      {
        // the next particle's record is requested before this one's splats are computed
        if (p0 < p1)
        {
          float4 cur = __ldg(part + p0);
          for (int k = p0; k < p1; ++k)
          {
            const float4 nxt = __ldg(part + min(k + 1, p1 - 1));
            p2g_particle(a, cur, d, ci, c, half_dx, half_dy);
            cur = nxt;
          }
        }
      }
      else
        for (int k = p0; k < p1; ++k) p2g_particle(a, __ldg(part + k), d, ci, c, half_dx, half_dy);
    }
    // u nodes of row c-1: own column + the part the west neighbour lane holds for us
    {
      const float s_w = __shfl_up_sync(0xffffffffu, a.us[0][1], 1);
      const float w_w = __shfl_up_sync(0xffffffffu, a.uw[0][1], 1);
      const float su = a.us[0][0] + s_w, wu = a.uw[0][0] + w_w;
      const int j = c - 1;
      if (writer && j >= ja && j < jb && (double)wu > 0.000001) ub[ci + (size_t)j * d.ld] = su / wu;
    }
    // v nodes of row c: own column + west neighbour's east part + east neighbour's west part
    {
      const float s_w = __shfl_up_sync(0xffffffffu, a.vs[0][2], 1);
      const float w_w = __shfl_up_sync(0xffffffffu, a.vw[0][2], 1);
      const float s_e = __shfl_down_sync(0xffffffffu, a.vs[0][0], 1);
      const float w_e = __shfl_down_sync(0xffffffffu, a.vw[0][0], 1);
      const float sv = (a.vs[0][1] + s_w) + s_e, wv = (a.vw[0][1] + w_w) + w_e;
      if (writer && c >= ja && c < jb && (double)wv > 0.000001) vb[ci + (size_t)c * d.ld] = sv / wv;
    }
    // rotate the row slots
#pragma unroll
    for (int q = 0; q < 2; ++q)
    {
      a.us[0][q] = a.us[1][q]; a.uw[0][q] = a.uw[1][q];
      a.us[1][q] = a.us[2][q]; a.uw[1][q] = a.uw[2][q];
      a.us[2][q] = 0.0f; a.uw[2][q] = 0.0f;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
    {
      a.vs[0][q] = a.vs[1][q]; a.vw[0][q] = a.vw[1][q];
      a.vs[1][q] = 0.0f; a.vw[1][q] = 0.0f;
    }
  }
}

// ------------------------------------------------------------------- G2P --
// src/FluidSolver.cpp:921-963 fused with src/MarkerParticleSet.cpp:40-62.
// mode < 0: advect only.  dt_advect == 0 && !do_advect: transfer only.
template <bool DO_G2P, bool DO_ADVECT>
__global__ void k_g2p_advect(float4* __restrict__ part, int64_t n, const float* __restrict__ uf,
                             const float* __restrict__ vf, const float* __restrict__ ud_or_prev,
                             const float* __restrict__ vd_or_prev, int diff_is_prev,
                             const uint8_t* __restrict__ cell, const GridDims d, int mode,
                             float pic_ratio, float dt, int ensure_outside)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 p = part[k];
  if (DO_G2P)
  {
    float nvx, nvy;
    if (mode == FSB_G2P_PIC)
    {
      nvx = vel_x_interp(uf, d, p.x, p.y);
      nvy = vel_y_interp(vf, d, p.x, p.y);
    }
    else if (diff_is_prev && mode == FSB_G2P_PICFLIP)
    {
      // front and (front - previous) at the same point: one pass over the taps
      float pic_x, pic_y, dvx, dvy;
      grid_interp_pair(uf, ud_or_prev, d, p.x, p.y - d.dy * 0.5f, &pic_x, &dvx);
      grid_interp_pair(vf, vd_or_prev, d, p.x - d.dx * 0.5f, p.y, &pic_y, &dvy);
      const float flip_x = p.z + dvx;
      const float flip_y = p.w + dvy;
      nvx = pic_x * pic_ratio + flip_x * (1.0f - pic_ratio);
      nvy = pic_y * pic_ratio + flip_y * (1.0f - pic_ratio);
    }
    else
    {
      float dvx, dvy;
      if (diff_is_prev)
      {
        dvx = grid_interp_diff(uf, ud_or_prev, d, p.x, p.y - d.dy * 0.5f);
        dvy = grid_interp_diff(vf, vd_or_prev, d, p.x - d.dx * 0.5f, p.y);
      }
      else
      {
        dvx = vel_x_interp(ud_or_prev, d, p.x, p.y);
        dvy = vel_y_interp(vd_or_prev, d, p.x, p.y);
      }
      const float flip_x = p.z + dvx;
      const float flip_y = p.w + dvy;
      if (mode == FSB_G2P_FLIP)
      {
        nvx = flip_x;
        nvy = flip_y;
      }
      else
      {
        const float pic_x = vel_x_interp(uf, d, p.x, p.y);
        const float pic_y = vel_y_interp(vf, d, p.x, p.y);
        nvx = pic_x * pic_ratio + flip_x * (1.0f - pic_ratio);
        nvy = pic_y * pic_ratio + flip_y * (1.0f - pic_ratio);
      }
    }
    p.z = nvx;
    p.w = nvy;
  }
  if (DO_ADVECT)
  {
    // include/MarkerParticleSet.h:37-41
    p.x += p.z * dt;
    p.y += p.w * dt;
    if (ensure_outside)
    {
      // src/MarkerParticleSet.cpp:55-60: the roll-back is a second advect(-dt)
      const int x = (int)div_dx(d, p.x);
      const int y = (int)div_dy(d, p.y);
      if (cell_type(cell, d, x, y) == FSB_SOLID)
      {
        const float mdt = -dt;
        p.x += p.z * mdt;
        p.y += p.w * mdt;
      }
    }
  }
  part[k] = p;
}

// The transfer of the fused steps as straight-line code: mode and the delta fast path are
// compile-time, the (front - previous) difference is taken per tap (src/MacGrid.cpp:64-67), and
// all sixteen taps are independent loads the compiler can issue back to back.  Per particle the
// arithmetic is that of k_g2p_advect<true, true> with diff_is_prev (bit-identical).
// one particle of the fused transfer + blend + advection
template <int MODE, class D>
__device__ __forceinline__ float4 g2p_one(float4 p, const float* __restrict__ uf, const float* __restrict__ vf,
                                          const float* __restrict__ up, const float* __restrict__ vp,
                                          const uint8_t* __restrict__ cell, const D& d, float pic_ratio,
                                          float dt, int ensure_outside)
{
  float nvx, nvy;
  if (MODE == FSB_G2P_PIC)
  {
    nvx = vel_x_interp(uf, d, p.x, p.y);
    nvy = vel_y_interp(vf, d, p.x, p.y);
  }
  else if (MODE == FSB_G2P_FLIP)
  {
    nvx = p.z + grid_interp_diff(uf, up, d, p.x, p.y - d.dy * 0.5f);
    nvy = p.w + grid_interp_diff(vf, vp, d, p.x - d.dx * 0.5f, p.y);
  }
  else
  {
    float pic_x, pic_y, dvx, dvy;
    grid_interp_pair(uf, up, d, p.x, p.y - d.dy * 0.5f, &pic_x, &dvx);
    grid_interp_pair(vf, vp, d, p.x - d.dx * 0.5f, p.y, &pic_y, &dvy);
    const float flip_x = p.z + dvx;
    const float flip_y = p.w + dvy;
    nvx = pic_x * pic_ratio + flip_x * (1.0f - pic_ratio);
    nvy = pic_y * pic_ratio + flip_y * (1.0f - pic_ratio);
  }
  p.z = nvx;
  p.w = nvy;
  // include/MarkerParticleSet.h:37-41
  p.x += p.z * dt;
  p.y += p.w * dt;
  if (ensure_outside)
  {
    // src/MarkerParticleSet.cpp:55-60: the roll-back is a second advect(-dt)
    const int x = (int)div_dx(d, p.x);
    const int y = (int)div_dy(d, p.y);
    if (cell_type(cell, d, x, y) == FSB_SOLID)
    {
      const float mdt = -dt;
      p.x += p.z * mdt;
      p.y += p.w * mdt;
    }
  }
  return p;
}

// PER particles per thread, block-strided so that every access stays coalesced: a particle is two
// dependent rounds of loads (its record, then the 16 grid taps its position selects); with one particle per
// thread the kernel is bound by that chain times the number of thread waves (207 at 6.3e7 particles), not by
// bandwidth.  All records are loaded first, all stores come last, so the PER chains overlap.
template <int MODE, class D, int PER>
__global__ void __launch_bounds__(256)
k_g2p_step(float4* __restrict__ part, int64_t n, const float* __restrict__ uf,
           const float* __restrict__ vf, const float* __restrict__ up, const float* __restrict__ vp,
           const uint8_t* __restrict__ cell, const D d, float pic_ratio, float dt, int ensure_outside)
{
  const int64_t k0 = (int64_t)blockIdx.x * (blockDim.x * PER) + threadIdx.x;
  float4 p[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    p[q] = (k < n) ? part[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < PER; ++q)
    if (k0 + (int64_t)q * blockDim.x < n)
      p[q] = g2p_one<MODE, D>(p[q], uf, vf, up, vp, cell, d, pic_ratio, dt, ensure_outside);
#pragma unroll
  for (int q = 0; q < PER; ++q)
  {
    const int64_t k = k0 + (int64_t)q * blockDim.x;
    if (k < n) part[k] = p[q];
  }
}

// src/FluidSolver.cpp:774-791
template <class D>
__global__ void k_advect_particles_grid(float4* __restrict__ part, int64_t n,
                                        const float* __restrict__ uf, const float* __restrict__ vf,
                                        const D d, float dt, int integrator)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float4 p = part[k];
  float xn, yn;
  advected_position(uf, vf, d, integrator, p.x, p.y, dt, &xn, &yn);
  p.x = xn;
  p.y = yn;
  part[k] = p;
}

__global__ void k_unpermute(const float4* __restrict__ part, const int* __restrict__ orig,
                            int64_t n, float4* __restrict__ dst)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  dst[orig[k]] = part[k];
}

// src/FluidDomain.cpp:39-46: lattice in y-outer / x-inner order; the coordinate
// sequences (repeated float additions) are produced on the host.
__global__ void k_emit_source(float4* __restrict__ part, int* __restrict__ orig, int64_t first,
                              const float* __restrict__ xs, const float* __restrict__ ys,
                              int64_t count_x, int64_t count_y, float vel_x, float vel_y)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count_x * count_y) return;
  const int64_t ix = k % count_x, iy = k / count_x;
  part[first + k] = make_float4(xs[ix], ys[iy], vel_x, vel_y);
  orig[first + k] = (int)(first + k);
}

// src/FluidSolver.cpp:816-871 transferVelocityToGridGather.  weight = 1 - (|dx|/deltaX +
// |dy|/deltaY) >= 1 selects the particles numerically ON the face position (the sum must be
// <= 2^-25); such a particle lies in the 3x3 sort cells around the face position, so the
// reference's scan over the whole set becomes a scan of nine cells.  The reference adds the
// selected particles in particle-set order: they are picked by ascending original index.
__device__ __forceinline__ void gather_face(const float4* __restrict__ part,
                                            const int* __restrict__ orig,
                                            const int* __restrict__ cell_start, const GridDims& pd,
                                            const GridDims& d, float xf, float yf, bool want_x,
                                            float* wsum, float* vsum)
{
  const int ci = clampi((int)div_dx(pd, xf), 0, pd.nx - 1);
  const int cj = clampi((int)div_dy(pd, yf), 0, pd.ny - 1);
  float ws = 0.0f, vs = 0.0f;
  int last = -1;
  for (;;)
  {
    int best = 0x7fffffff;
    float bw = 0.0f, bv = 0.0f;
    for (int jj = max(cj - 1, 0); jj <= min(cj + 1, pd.ny - 1); ++jj)
      for (int ii = max(ci - 1, 0); ii <= min(ci + 1, pd.nx - 1); ++ii)
      {
        const int c0 = ii + jj * pd.nx;
        for (int k = cell_start[c0]; k < cell_start[c0 + 1]; ++k)
        {
          const int o = orig[k];
          if (o <= last || o >= best) continue;
          const float4 p = part[k];
          const float ax = fabsf(p.x - xf);
          const float ay = fabsf(p.y - yf);
          const float w = 1.0f - (ax / d.dx + ay / d.dy);
          if (w >= 1.0f)
          {
            best = o;
            bw = w;
            bv = want_x ? p.z : p.w;
          }
        }
      }
    if (best == 0x7fffffff) break;
    ws += bw;
    vs += bv;
    last = best;
  }
  *wsum = ws;
  *vsum = vs;
}

__global__ void k_p2g_gather(const float4* __restrict__ part, const int* __restrict__ orig,
                             const int* __restrict__ cell_start, float* __restrict__ ub,
                             float* __restrict__ vb, const GridDims pd, const GridDims d)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= d.nx || j >= d.ny) return;
  // :823-826: i * deltaX in float; (j + 0.5) * deltaY in double, rounded to MyFloat
  const float x_u = (float)i * d.dx;
  const float y_u = (float)(((double)j + 0.5) * (double)d.dy);
  const float x_v = (float)(((double)i + 0.5) * (double)d.dx);
  const float y_v = (float)j * d.dy;
  float w, v;
  gather_face(part, orig, cell_start, pd, d, x_u, y_u, true, &w, &v);
  if (w != 0.0f) ub[i + (size_t)j * d.ld] = v / w;
  gather_face(part, orig, cell_start, pd, d, x_v, y_v, false, &w, &v);
  if (w != 0.0f) vb[i + (size_t)j * d.ld] = v / w;
}

// ------------------------------------------------------------ particle slabs --
// Slab-partitioned particle sets (one process per GPU, every rank holds the full grids but only
// the particles of its row slab plus one ghost row of its neighbours').  Ownership is a pure
// function of the position: the rank whose rows contain the particle's sort-key row.
__device__ __forceinline__ int slab_row(const GridDims& d, float y)
{
  return clampi((int)div_dy(d, y), 0, d.ny - 1);
}
__device__ __forceinline__ int slab_owner(int row, int ny, int world)
{
  // rows [ny*q/world, ny*(q+1)/world) belong to rank q
  int q = (int)(((int64_t)row * world) / ny);
  while (q + 1 < world && row >= (int)(((int64_t)ny * (q + 1)) / world)) ++q;
  while (q > 0 && row < (int)(((int64_t)ny * q) / world)) --q;
  return q;
}

// particles that are not this rank's (ghosts) are marked dead (orig = -1) before the transfer back
__global__ void k_slab_mark_ghosts(const float4* __restrict__ part, int* __restrict__ orig, int64_t n,
                                   const GridDims d, int lo, int hi)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int row = slab_row(d, part[k].y);
  if (row < lo || row >= hi) orig[k] = -1;
}

// destination rank of every live particle; counts per rank (block histogram + atomics)
__global__ void k_slab_count(const float4* __restrict__ part, const int* __restrict__ orig, int64_t n,
                             const GridDims d, int world, int* __restrict__ dest,
                             unsigned long long* __restrict__ counts)
{
  __shared__ unsigned int s_cnt[kMaxRanks];
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n)
  {
    int q = -1;
    if (orig[k] >= 0) q = slab_owner(slab_row(d, part[k].y), d.ny, world);
    dest[k] = q;
    if (q >= 0) atomicAdd(&s_cnt[q], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kMaxRanks && s_cnt[threadIdx.x])
    atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// scatter into per-destination groups (order inside a group is irrelevant: the cell sort's
// in-cell order goes by original index)
__global__ void k_slab_scatter(const float4* __restrict__ part, const int* __restrict__ orig,
                               const int* __restrict__ dest, int64_t n,
                               unsigned long long* __restrict__ cursor, float4* __restrict__ part_out,
                               int* __restrict__ orig_out)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int q = dest[k];
  if (q < 0) return;
  const unsigned long long slot = atomicAdd(&cursor[q], 1ull);
  part_out[slot] = part[k];
  orig_out[slot] = orig[k];
}

// owned particles whose row is `row`: count, then compact into a buffer
__global__ void k_slab_row_select(const float4* __restrict__ part, const int* __restrict__ orig,
                                  int64_t n, const GridDims d, int row,
                                  unsigned long long* __restrict__ cursor, float4* __restrict__ part_out,
                                  int* __restrict__ orig_out)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || orig[k] < 0) return;
  if (slab_row(d, part[k].y) != row) return;
  const unsigned long long slot = atomicAdd(cursor, 1ull);
  if (part_out)
  {
    part_out[slot] = part[k];
    orig_out[slot] = orig[k];
  }
}

} // namespace

int fsb_k_slab_mark_ghosts(fsb_ctx* c)
{
  if (c->n == 0) return FSB_OK;
  k_slab_mark_ghosts<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->orig[c->pcur], c->n, pool_dims(c), c->slab_lo, c->slab_hi);
  FSB_LAUNCHED(c);
  return FSB_OK;
}

// groups the live particles by owner rank into the other buffer; counts[world] on return
int fsb_k_slab_sort_out(fsb_ctx* c, int64_t* counts)
{
  const int world = c->slab_world;
  for (int q = 0; q < kMaxRanks; ++q) c->slab_count[q] = 0;
  if (c->n > 0)
  {
    unsigned long long* dev = c->slab_ctr; // 2 * kMaxRanks words
    unsigned long long host[2 * kMaxRanks];
    FSB_CUDA(c, cudaMemsetAsync(dev, 0, sizeof(unsigned long long) * 2 * kMaxRanks, c->stream));
    k_slab_count<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->orig[c->pcur], c->n, pool_dims(c), world, c->sort_key, dev);
    FSB_LAUNCHED(c);
    FSB_CUDA(c, cudaMemcpyAsync(host, dev, sizeof(unsigned long long) * kMaxRanks,
                                cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    unsigned long long off = 0;
    for (int q = 0; q < kMaxRanks; ++q)
    {
      c->slab_count[q] = (int64_t)host[q];
      host[kMaxRanks + q] = off; // cursor start of group q
      off += host[q];
    }
    FSB_CUDA(c, cudaMemcpyAsync(dev + kMaxRanks, host + kMaxRanks, sizeof(unsigned long long) * kMaxRanks,
                                cudaMemcpyHostToDevice, c->stream));
    k_slab_scatter<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->orig[c->pcur], c->sort_key, c->n, dev + kMaxRanks, c->part[c->pcur ^ 1],
        c->orig[c->pcur ^ 1]);
    FSB_LAUNCHED(c);
    FSB_CUDA(c, cudaStreamSynchronize(c->stream)); // `host` leaves scope
    c->pcur ^= 1;
    c->n = (int64_t)off;
  }
  c->sort_valid = false;
  c->slab_grouped = true;
  for (int q = 0; q < world; ++q) counts[q] = c->slab_count[q];
  return FSB_OK;
}

// selects this rank's particles of one row into slab_buf; *n_out = how many
int fsb_k_slab_row_select(fsb_ctx* c, int row, int64_t* n_out)
{
  *n_out = 0;
  if (c->n == 0) return FSB_OK;
  unsigned long long* dev = c->slab_ctr;
  unsigned long long cnt = 0;
  const GridDims d = pool_dims(c);
  FSB_CUDA(c, cudaMemsetAsync(dev, 0, sizeof(unsigned long long), c->stream));
  k_slab_row_select<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->orig[c->pcur], c->n, d, row, dev, nullptr, nullptr);
  FSB_LAUNCHED(c);
  FSB_CUDA(c, cudaMemcpyAsync(&cnt, dev, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if ((int64_t)cnt > c->slab_buf_cap)
  {
    if (c->slab_buf_part) cudaFree(c->slab_buf_part);
    if (c->slab_buf_orig) cudaFree(c->slab_buf_orig);
    c->slab_buf_part = nullptr; c->slab_buf_orig = nullptr;
    const int64_t cap = std::max<int64_t>(1024, (int64_t)cnt * 2);
    FSB_CUDA(c, cudaMalloc(&c->slab_buf_part, sizeof(float4) * cap));
    FSB_CUDA(c, cudaMalloc(&c->slab_buf_orig, sizeof(int) * cap));
    c->slab_buf_cap = cap;
  }
  if (cnt > 0)
  {
    FSB_CUDA(c, cudaMemsetAsync(dev, 0, sizeof(unsigned long long), c->stream));
    k_slab_row_select<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->orig[c->pcur], c->n, d, row, dev, c->slab_buf_part, c->slab_buf_orig);
    FSB_LAUNCHED(c);
  }
  *n_out = (int64_t)cnt;
  return FSB_OK;
}

namespace {

} // namespace

int fsb_k_p2g_gather(fsb_ctx* c)
{
  FSB_TRY(fsb_k_sort_particles(c));
  fsb_prof_begin(c, FSB_PROF_P2G);
  k_p2g_gather<<<dim3(fsb_div_up(c->nx, kBlock), c->ny), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->orig[c->pcur], c->cell_start, fsb_ub(c), fsb_vb(c), pool_dims(c), dims(c));
  FSB_LAUNCHED(c);
  c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:870
  fsb_prof_end(c, FSB_PROF_P2G);
  return FSB_OK;
}

int fsb_k_sort_particles(fsb_ctx* c, bool mark_labels)
{
  if (c->sort_valid) return mark_labels ? fsb_fail(c, FSB_ERR_INVALID, "sort already valid") : FSB_OK;
  const int m = c->nx * c->ny;
  fsb_prof_begin(c, FSB_PROF_SORT);
  FSB_CUDA(c, cudaMemsetAsync(c->cell_count, 0, (size_t)m * sizeof(int), c->stream));
  if (c->n > 0)
  {
    // lengthX() is recomputed as size * delta in float (include/Grid.h:54-55)
    const GridDims len =
        make_grid_dims(c->nx, c->ny, c->ld, (float)c->nx * c->dx, (float)c->ny * c->dy);
    const GridDims pd = pool_dims(c);
    const int per = c->sort_per; // particles per thread (FSB_SORT_PER: 1, 2 or 4)
    const dim3 grid(fsb_div_up(c->n, (int64_t)kBlock * per));
#define FSB_COUNT_N(MARK, D, pd_, len_, PER)                                                             \
  k_sort_count<MARK, D, PER><<<grid, kBlock, 0, c->stream>>>(c->part[c->pcur], c->n, pd_, c->cell_count, \
                                                             c->sort_key, c->sort_rank, c->cell, dims(c), \
                                                             len_)
#define FSB_COUNT(MARK, D, pd_, len_)                       \
  do {                                                      \
    if (per == 4) FSB_COUNT_N(MARK, D, pd_, len_, 4);       \
    else if (per == 2) FSB_COUNT_N(MARK, D, pd_, len_, 2);  \
    else FSB_COUNT_N(MARK, D, pd_, len_, 1);                \
  } while (0)
    const bool p2 = pd.pow2 == 3 && (!mark_labels || len.pow2 == 3);
    if (p2 && mark_labels) FSB_COUNT(true, GridDimsP2, as_pow2(pd), as_pow2(len));
    else if (p2) FSB_COUNT(false, GridDimsP2, as_pow2(pd), as_pow2(len));
    else if (mark_labels) FSB_COUNT(true, GridDims, pd, len);
    else FSB_COUNT(false, GridDims, pd, len);
#undef FSB_COUNT
#undef FSB_COUNT_N
    FSB_LAUNCHED(c);
  }
  const int nb = fsb_div_up(m, kScanTile);
  k_scan_reduce<<<nb, kScanBlock, 0, c->stream>>>(c->cell_count, m, c->scan_block);
  FSB_LAUNCHED(c);
  k_scan_blocks<<<1, 1024, 0, c->stream>>>(c->scan_block, nb);
  FSB_LAUNCHED(c);
  k_scan_apply<<<nb, kScanBlock, 0, c->stream>>>(c->cell_count, m, c->scan_block, c->cell_start);
  FSB_LAUNCHED(c);
  if (c->n > 0)
  {
    k_sort_place<<<fsb_div_up(fsb_div_up(c->n, 4), kBlock), kBlock, 0, c->stream>>>(
        c->sort_key, c->sort_rank, c->n, c->cell_start, c->sort_idx);
    FSB_LAUNCHED(c);
    if (c->canon_per == 2)
      k_sort_canon<2><<<fsb_div_up(m, kBlock * 2), kBlock, 0, c->stream>>>(c->cell_start, m, c->sort_idx, c->orig[c->pcur]);
    else
      k_sort_canon<1><<<fsb_div_up(m, kBlock), kBlock, 0, c->stream>>>(c->cell_start, m, c->sort_idx, c->orig[c->pcur]);
    FSB_LAUNCHED(c);
    const int per = c->sort_per;
    const int gather_grid = fsb_div_up(c->n, (int64_t)kBlock * per);
    if (per == 4)
      k_sort_gather<4><<<gather_grid, kBlock, 0, c->stream>>>(c->part[c->pcur], c->orig[c->pcur], c->sort_idx, c->n,
                                                              c->part[c->pcur ^ 1], c->orig[c->pcur ^ 1]);
    else if (per == 2)
      k_sort_gather<2><<<gather_grid, kBlock, 0, c->stream>>>(c->part[c->pcur], c->orig[c->pcur], c->sort_idx, c->n,
                                                              c->part[c->pcur ^ 1], c->orig[c->pcur ^ 1]);
    else
      k_sort_gather<1><<<gather_grid, kBlock, 0, c->stream>>>(c->part[c->pcur], c->orig[c->pcur], c->sort_idx, c->n,
                                                              c->part[c->pcur ^ 1], c->orig[c->pcur ^ 1]);
    FSB_LAUNCHED(c);
    c->pcur ^= 1;
  }
  c->sort_valid = true;
  fsb_prof_end(c, FSB_PROF_SORT);
  return FSB_OK;
}

int fsb_k_p2g(fsb_ctx* c)
{
  FSB_TRY(fsb_k_sort_particles(c));
  fsb_prof_begin(c, FSB_PROF_P2G);
  const GridDims d = pool_dims(c);
  const int strips_x = fsb_div_up(c->nx, kP2gCols);
  // slab-partitioned particles: only the faces of this rank's rows (the other rows come from their owners)
  const int row_lo = c->slab_world > 1 ? c->slab_lo : 0, row_hi = c->slab_world > 1 ? c->slab_hi : c->ny;
  const int64_t n_warps = (int64_t)strips_x * fsb_div_up(row_hi - row_lo, kP2gRows);
#define FSB_P2G(D, d_, PIPE)                                                                         \
  k_p2g_stream<D, PIPE><<<fsb_div_up(n_warps * 32, 256), 256, 0, c->stream>>>(                        \
      c->part[c->pcur], c->cell_start, fsb_ub(c), fsb_vb(c), d_, 0.5f * c->dx, 0.5f * c->dy, strips_x, \
      (int)n_warps, row_lo, row_hi)
  if (d.pow2 == 3) { if (c->p2g_pipe) FSB_P2G(GridDimsP2, as_pow2(d), true); else FSB_P2G(GridDimsP2, as_pow2(d), false); }
  else { if (c->p2g_pipe) FSB_P2G(GridDims, d, true); else FSB_P2G(GridDims, d, false); }
#undef FSB_P2G
  FSB_LAUNCHED(c);
  c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:918
  fsb_prof_end(c, FSB_PROF_P2G);
  return FSB_OK;
}

int fsb_k_g2p(fsb_ctx* c, int mode, float pic_ratio)
{
  if (c->n == 0) return FSB_OK;
  fsb_prof_begin(c, FSB_PROF_G2P);
  k_g2p_advect<true, false><<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->n, fsb_uf(c), fsb_vf(c), c->u_diff, c->v_diff, 0, c->cell, dims(c), mode,
      pic_ratio, 0.0f, 0);
  FSB_LAUNCHED(c);
  fsb_prof_end(c, FSB_PROF_G2P);
  return FSB_OK;
}

int fsb_k_advect_particles(fsb_ctx* c, float dt, int ensure_outside)
{
  if (c->n == 0) return FSB_OK;
  fsb_prof_begin(c, FSB_PROF_G2P);
  k_g2p_advect<false, true><<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->n, nullptr, nullptr, nullptr, nullptr, 0, c->cell, dims(c), 0, 0.0f, dt,
      ensure_outside);
  FSB_LAUNCHED(c);
  c->sort_valid = false;
  fsb_prof_end(c, FSB_PROF_G2P);
  return FSB_OK;
}

// G2P on (front - previous) taken per tap + blend + advect in one pass: the
// diff buffer is never materialised (src/MacGrid.cpp:58-70 folded in).
int fsb_k_g2p_advect(fsb_ctx* c, int mode, float pic_ratio, float dt, int ensure_outside)
{
  if (c->n == 0) return FSB_OK;
  fsb_prof_begin(c, FSB_PROF_G2P);
  const GridDims d = dims(c);
  const dim3 grid(fsb_div_up(c->n, kBlock));
  const int per = c->g2p_per; // particles per thread (FSB_G2P_PER: 1, 2 or 4)
  const dim3 grid_per(fsb_div_up(c->n, (int64_t)kBlock * per));
  if (c->stage_v1)
    k_g2p_advect<true, true><<<grid, kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->n, fsb_uf(c), fsb_vf(c), c->u_prev, c->v_prev, 1, c->cell, d, mode,
        pic_ratio, dt, ensure_outside);
  else
  {
#define FSB_G2P_N(MODE, D, d_, PER)                                                                 \
  k_g2p_step<MODE, D, PER><<<grid_per, kBlock, 0, c->stream>>>(                                      \
      c->part[c->pcur], c->n, fsb_uf(c), fsb_vf(c), c->u_prev, c->v_prev, c->cell, d_, pic_ratio, dt, \
      ensure_outside)
#define FSB_G2P(MODE, D, d_)                          \
  do {                                                \
    if (per == 4) FSB_G2P_N(MODE, D, d_, 4);          \
    else if (per == 2) FSB_G2P_N(MODE, D, d_, 2);     \
    else FSB_G2P_N(MODE, D, d_, 1);                   \
  } while (0)
    const bool p2 = d.pow2 == 3;
    if (mode == FSB_G2P_PIC) { if (p2) FSB_G2P(FSB_G2P_PIC, GridDimsP2, as_pow2(d)); else FSB_G2P(FSB_G2P_PIC, GridDims, d); }
    else if (mode == FSB_G2P_FLIP) { if (p2) FSB_G2P(FSB_G2P_FLIP, GridDimsP2, as_pow2(d)); else FSB_G2P(FSB_G2P_FLIP, GridDims, d); }
    else { if (p2) FSB_G2P(FSB_G2P_PICFLIP, GridDimsP2, as_pow2(d)); else FSB_G2P(FSB_G2P_PICFLIP, GridDims, d); }
#undef FSB_G2P
#undef FSB_G2P_N
  }
  FSB_LAUNCHED(c);
  c->sort_valid = false;
  fsb_prof_end(c, FSB_PROF_G2P);
  return FSB_OK;
}

int fsb_k_advect_particles_grid(fsb_ctx* c, float dt)
{
  if (c->n == 0) return FSB_OK;
  fsb_prof_begin(c, FSB_PROF_ADVECT_PART);
  const GridDims d = dims(c);
  if (d.pow2 == 3)
    k_advect_particles_grid<GridDimsP2><<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->n, fsb_uf(c), fsb_vf(c), as_pow2(d), dt, c->integrator);
  else
    k_advect_particles_grid<GridDims><<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(
        c->part[c->pcur], c->n, fsb_uf(c), fsb_vf(c), d, dt, c->integrator);
  FSB_LAUNCHED(c);
  c->sort_valid = false;
  fsb_prof_end(c, FSB_PROF_ADVECT_PART);
  return FSB_OK;
}

int fsb_k_unpermute(fsb_ctx* c, float4* dst_dense)
{
  if (c->n == 0) return FSB_OK;
  k_unpermute<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(c->part[c->pcur],
                                                                  c->orig[c->pcur], c->n, dst_dense);
  FSB_LAUNCHED(c);
  return FSB_OK;
}

int fsb_k_emit_source_dev(fsb_ctx* c, int64_t first, const float* xs_dev, const float* ys_dev,
                          int64_t count_x, int64_t count_y, float vel_x, float vel_y)
{
  const int64_t total = count_x * count_y;
  if (total == 0) return FSB_OK;
  k_emit_source<<<fsb_div_up(total, kBlock), kBlock, 0, c->stream>>>(
      c->part[c->pcur], c->orig[c->pcur], first, xs_dev, ys_dev, count_x, count_y, vel_x, vel_y);
  FSB_LAUNCHED(c);
  return FSB_OK;
}
