// One-sweep Jacobi-preconditioned CG: the default solve of the pressure projection
// (src/FluidSolver.cpp:418-426, Eigen's ConjugateGradient as restated in SURVEY.md Appendix B).
//
// Eigen's iteration has two reduction points (p.Ap before the update, r.z after it), which forces
// two sweeps over the grid and two grid-wide (multi-GPU: cross-GPU) synchronisations per
// iteration.  This kernel runs the SAME iteration -- same alpha, same x / r / p updates in the
// same statement order, same stopping test on the exact |r|^2 -- with ONE sweep and ONE
// reduction point per iteration, and no extra vectors:
//
//   sweep k  (knows alpha_k, beta_k; reads p_k with a two-cell halo, r_k with a one-cell halo)
//     on the tile and its one-cell halo:  q_k = A p_k,  r_{k+1} = r_k - alpha_k q_k,
//                                         z_{k+1} = D^-1 r_{k+1},  p_{k+1} = z_{k+1} + beta_k p_k
//     on the tile:                        q_{k+1} = A p_{k+1}   (registers only)
//                                         x += alpha_k p_k      (odd k: two pending updates at once)
//     partial sums over the tile:  p_{k+1}.q_{k+1},  r_{k+1}.z_{k+1},  |r_{k+1}|^2,
//                                  z_{k+1}.q_{k+1},  q_{k+1}.D^-1 q_{k+1}
//   reduction point:  stop if |r_{k+1}|^2 < threshold (Eigen's test, on the exact norm);
//                     alpha_{k+1} = r_{k+1}.z_{k+1} / p_{k+1}.q_{k+1}            (as Eigen)
//                     beta_{k+1}  = (r_{k+2}.z_{k+2}) / (r_{k+1}.z_{k+1})  with the numerator from
//                                   the identity  r'.z' = r.z - 2 alpha z.q + alpha^2 q.D^-1 q
//                                   (r' = r - alpha q), evaluated in double from the exact dot
//                                   products of the vectors of THIS sweep.
// The identity is exact algebra on the vectors actually stored; the only difference to Eigen's
// beta is that r' enters before its fp32 rounding (relative 1e-7, the size of beta's own
// rounding).  Nothing is carried by recurrence from one iteration to the next: every alpha comes
// from exact dot products, so there is no drift.  numpy fp32 study: identical iteration counts
// (910 / 1714 at 256^2 / 512^2 to 1e-6, 529 / 1007 at 128^2 / 256^2 to FLT_EPSILON) and pressure
// within 2e-6 of the two-reduction iteration (tools/studies/cg_one_sweep_study.py).
//
// HBM traffic per cell and iteration: r 4 + p 4 + code 1 in, r 4 + p 4 out, x (4 + 4 + 4) / 2
// (deferred: x is touched on odd iterations only) = 23 B, against 32 B for the two-sweep kernel
// of fsb_cg.cu and 45 B for the textbook formulation.  r and p are double-buffered
// (neighbouring tiles still read the old halo while a tile is being written).
//
// Frame: as k_cg_solve (fsb_cg_frame.cuh) -- one persistent cooperative kernel for the whole
// solve, 8 consumer warps + 1 TMA producer warp per CTA, 2 CTAs per SM, mbarrier ring; the
// grid barrier carries the reduction and every CTA derives the scalars itself.  Sharded solves:
// a slab's two boundary rows of p and one of r are stored straight into the neighbours' ghost
// rows (NVLink), and the five sums cross the GPUs through ONE mailbox round per iteration.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "fsb_cg_frame.cuh"
#include "fsb_cg_one_scalars.h"

namespace {

template <int TH>
struct OneStage // p_old (two-cell halo), r_old (one-cell halo), code (one-cell halo)
{
  static constexpr int kP = kHaloW * (TH + 4) * 4; // rows j0-2 .. j0+TH+1
  static constexpr int kR = kHaloW * (TH + 2) * 4; // rows j0-1 .. j0+TH
  static constexpr int kCode = kCodeW * (TH + 2);
  static constexpr int oP = 0, oR = align128(kP), oC = oR + align128(kR);
  static constexpr int kBytes = align128(oC + kCode);
  static constexpr int kTx = kP + kR + kCode;
};

struct OneMaps
{
  CUtensorMap halo_p[2]; // fp32, box kHaloW x (TH+4)
  CUtensorMap halo_r[2]; // fp32, box kHaloW x (TH+2)
  CUtensorMap code;      // u8,   box kCodeW x (TH+2)
};
static_assert(sizeof(OneMaps) <= sizeof(((fsb_ctx*)nullptr)->cg_maps_one), "cg_maps_one too small");

// peers' copies of the direction / residual buffers (whole arrays; null: no neighbour that side)
struct OnePeers
{
  float *p_lo[2], *p_hi[2], *r_lo[2], *r_hi[2];
};

constexpr int kNSums = 5; // p.q, r.z, |r|^2, z.q, q.D^-1 q

__device__ __forceinline__ void st_mail2(unsigned long long* p, unsigned long long a, unsigned long long b)
{
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_mail2(const unsigned long long* p, unsigned long long& a,
                                         unsigned long long& b)
{
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Barrier + reduction of the kNSums sums for the NW consumer warps of every CTA, then the scalar
// step of the iteration on the CTA's own copy.  Single GPU: warp 0 of EVERY CTA waits for the
// arrival counter and folds all partials itself in a fixed order (identical bits everywhere, no
// broadcast hop).  Sharded: CTA 0 folds the slab's partials and posts them into every rank's
// mailbox (lane q -> rank q); warp 0 of every CTA polls its own rank's mailbox and adds the
// `world` entries in rank order.
template <int NW>
__device__ __forceinline__ void one_reduce(double (&acc)[kNSums], CgScalars* s,
                                           double* __restrict__ partials, unsigned phase_id,
                                           const ShardArgs& sh, OneState* ss, bool pushed,
                                           int sys_flags, double* __restrict__ dbg = nullptr)
{
  __shared__ double s_part[kNSums][NW];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Every thread orders its own generic-proxy stores of this sweep against the async proxy (the TMA
  // loads of the next sweep): a proxy fence is not cumulative.  Visibility at GPU / system scope
  // follows from lane 0's fence below, which is cumulative over everything the CTA barrier in
  // between has ordered before it (the pattern of cooperative-groups grid.sync()).
  fence_proxy_async_all();
#pragma unroll
  for (int n = 0; n < kNSums; ++n)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_down_sync(0xffffffffu, acc[n], o);
    if (lane == 0) s_part[n][warp] = acc[n];
  }
  consumer_sync(NW * 32);
  __shared__ double s_fold[kNSums][NW];
  __shared__ int s_folder;
  const int G = (int)gridDim.x;
  const unsigned int target = phase_id * (unsigned int)G;
  double tot[kNSums];
#pragma unroll
  for (int n = 0; n < kNSums; ++n) tot[n] = 0.0;
  // FSB_CG_DEBUG_TIMES (with FSB_CG_DEBUG_SUMS): %globaltimer stamps of reductions 100..163 instead of the sums
  double* tdbg = nullptr;
  if (dbg && (sys_flags & 4) && threadIdx.x == 0 && phase_id >= 100 && phase_id < 164)
    tdbg = dbg + ((size_t)(phase_id - 100) * G + blockIdx.x) * 10;
#define FSB_STAMP(i) do { if (tdbg) tdbg[i] = (double)global_ns(); } while (0)
  if (warp == 0)
  {
#pragma unroll
    for (int n = 0; n < kNSums; ++n)
    {
      double v = (lane < NW) ? s_part[n][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      tot[n] = v;
    }
    FSB_STAMP(0);
    // partials: one 64-byte record per CTA (five sums), two regions used alternately
    if (lane == 0)
    {
      double2* rec = reinterpret_cast<double2*>(partials + (size_t)blockIdx.x * 8);
      rec[0] = make_double2(tot[0], tot[1]);
      rec[1] = make_double2(tot[2], tot[3]);
      partials[(size_t)blockIdx.x * 8 + 4] = tot[4];
      if (dbg && !(sys_flags & 4) && phase_id <= 64) // FSB_CG_DEBUG_SUMS: this CTA's own sums of the sweep
        for (int n = 0; n < kNSums; ++n) dbg[((size_t)(phase_id - 1) * G + blockIdx.x) * 10 + n] = tot[n];
      // the CTA's global stores (p / r / x rows, peer rows) must be visible to the TMA loads of the
      // next sweep on every SM (and GPU) before the arrival is
      fence_proxy_async_all();
      if (pushed) fence_acq_rel_sys(); // peer rows: acknowledged by the peer before this CTA arrives
      // Who folds the partials: single GPU -- every CTA, as soon as the counter is complete (no
      // broadcast hop); sharded -- the CTA that arrived LAST (it needs no wait at all), which then
      // posts the slab's sums into every rank's mailbox.  The arrival is a release (the CTA's stores,
      // ordered before it by the CTA barrier above, are visible to whoever sees the count), the wait an
      // acquire -- attached to the atomic / the polling load themselves, no separate fences.
      int folder = 1;
      if (sh.world == 1)
      {
        red_release_gpu_add(&s->bar_count, 1u); // no return value needed: a reduction, not a round trip
        FSB_STAMP(1);
        SpinGuard g;
        while (ld_acquire_gpu(&s->bar_count) < target) g.tick();
      }
      else
      {
        folder = (atom_acq_rel_gpu_add(&s->bar_count, 1u) + 1u == target) ? 1 : 0;
        FSB_STAMP(1);
      }
      s_folder = folder;
      FSB_STAMP(2);
    }
  }
  consumer_sync(NW * 32);
  const bool folder = s_folder != 0;
  if (folder)
  {
    // the fold is done by the whole CTA: thread t adds the records t, t + NW * 32, ... (at most two),
    // a butterfly inside each warp, then warp 0 adds the NW warp sums in warp order -- one fixed order
    // of additions, the same on every CTA; all loads are in flight together
    double f[kNSums];
#pragma unroll
    for (int n = 0; n < kNSums; ++n) f[n] = 0.0;
    for (int k = (int)threadIdx.x; k < G; k += NW * 32)
    {
      const double2* rec = reinterpret_cast<const double2*>(partials + (size_t)k * 8);
      const double2 a = __ldcg(rec), b2 = __ldcg(rec + 1);
      const double c = __ldcg(partials + (size_t)k * 8 + 4);
      f[0] += a.x; f[1] += a.y; f[2] += b2.x; f[3] += b2.y; f[4] += c;
    }
#pragma unroll
    for (int n = 0; n < kNSums; ++n)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) f[n] += __shfl_xor_sync(0xffffffffu, f[n], o);
      if (lane == 0) s_fold[n][warp] = f[n];
    }
    consumer_sync(NW * 32);
    if (warp == 0)
    {
#pragma unroll
      for (int n = 0; n < kNSums; ++n)
      {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_fold[n][w];
        tot[n] = v;
      }
    }
  }
  if (warp == 0)
  {
    FSB_STAMP(3);
    if (tdbg) tdbg[7] = folder ? 1.0 : 0.0;
    bool ok = true;
    const unsigned long long seq = ss->seq + 1;
    if (sh.world > 1)
    {
      const unsigned long long tag = seq & 0xffffffffull;
      if (folder)
      {
        // every local CTA's arrival (and its peer rows, each fenced at system scope by the CTA that
        // stored them) precedes the entry
        if (sys_flags & 2) fence_acq_rel_gpu();
        else fence_acq_rel_sys();
        if ((int)lane < sh.world)
        {
          unsigned long long* out = reinterpret_cast<unsigned long long*>(sh.mail[lane]) + kOneMailWord +
                                    ((int)(seq & 1ull) * kMaxRanks + sh.rank) * kOneMailStride;
#pragma unroll
          for (int n = 0; n < kNSums; ++n)
          {
            const unsigned long long b = (unsigned long long)__double_as_longlong(tot[n]);
            st_mail2(out + 2 * n, mail_word((unsigned int)b, seq), mail_word((unsigned int)(b >> 32), seq));
          }
        }
      }
      FSB_STAMP(4);
      double v[kNSums];
#pragma unroll
      for (int n = 0; n < kNSums; ++n) v[n] = 0.0;
      if ((int)lane < sh.world)
      {
        const unsigned long long* in = reinterpret_cast<const unsigned long long*>(sh.mail[sh.rank]) +
                                       kOneMailWord + ((int)(seq & 1ull) * kMaxRanks + (int)lane) * kOneMailStride;
        const unsigned long long t0 = global_ns();
        unsigned int spins = 0;
#pragma unroll
        for (int n = 0; n < kNSums; ++n)
        {
          unsigned long long w0, w1;
          for (;;)
          {
            ld_mail2(in + 2 * n, w0, w1);
            if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
            if ((++spins & 1023u) == 0 && global_ns() - t0 > kMailTimeoutNs)
            {
              if (ok && blockIdx.x == 0)
              {
                s->comm_diag[0] = (unsigned long long)(phase_id);
                s->comm_diag[1] = lane * 16ull + n;
                s->comm_diag[2] = tag;
                s->comm_diag[3] = w0;
              }
              ok = false;
              break;
            }
          }
          v[n] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
      }
      ok = __all_sync(0xffffffffu, ok);
      FSB_STAMP(5);
      // the peers' boundary rows were stored, and fenced at system scope, before their entries; they
      // live in THIS GPU's memory and the next sweep reads them with TMA loads that the producer issues
      // only after it has seen the release below (FSB_CG_POLL_FENCE_SYS=1: a system-scope fence here)
      if (!(sys_flags & 1)) fence_acq_rel_sys();
#pragma unroll
      for (int n = 0; n < kNSums; ++n)
      {
        double t = 0.0;
        for (int q = 0; q < sh.world; ++q) t += __shfl_sync(0xffffffffu, v[n], q); // rank order
        tot[n] = t;
      }
    }
    if (lane == 0)
    {
      if (dbg && !(sys_flags & 4) && phase_id <= 64) // ... and the totals as this CTA sees them
        for (int n = 0; n < kNSums; ++n) dbg[((size_t)(phase_id - 1) * gridDim.x + blockIdx.x) * 10 + 5 + n] = tot[n];
      ss->seq = seq;
      if (!ok)
      {
        ss->comm_error = 1;
        ss->done = 1;
      }
      else
      {
        one_advance(ss, tot[0], tot[1], tot[2], tot[3], tot[4]);
      }
      __threadfence_block();
      ss->released = phase_id;
      FSB_STAMP(6);
    }
  }
  consumer_sync(NW * 32);
}

template <int NW, int RPW>
__global__ void __launch_bounds__((NW + 1) * 32, 2)
k_cg_solve1(const __grid_constant__ OneMaps maps, float* __restrict__ x, float* __restrict__ r_a,
            float* __restrict__ r_b, float* __restrict__ p_a, float* __restrict__ p_b, int ld,
            int tiles_x, int n_tiles, int stages, const CgCoef coef, CgScalars* __restrict__ s,
            double* __restrict__ partials, const __grid_constant__ ShardArgs sh,
            const __grid_constant__ OnePeers peers, int flags, double* __restrict__ dbg)
{
  constexpr int TH = NW * RPW;
  using St = OneStage<TH>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  __shared__ float4 lut[8];
  __shared__ OneState ss;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool serp = (flags & 1) != 0; // consecutive sweeps walk the tile list in opposite directions
  const bool xdefer = (flags & 128) != 0;
  const bool xhint = (flags & 2) != 0;

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
    ss.rz = vs->rz; ss.r2 = vs->r2;
    ss.alpha = 0.0f; ss.beta = 0.0f; ss.thr = vs->thr;
    ss.iter = vs->iter; ss.done = vs->done; ss.max_iters = vs->max_iters; ss.comm_error = 0;
    ss.sweep = -1;
    ss.seq = vs->seq[3];
    ss.released = 0;
  }
  __syncthreads();
  const int* __restrict__ tile_list = s->tile_list; // fixed for the whole solve
  // flags bit 8 (FSB_CG_DEBUG_NOTILES, measurement only): sweeps without tiles -- the time per
  // "iteration" is then the bare cost of the barrier + reduction + scalar step
  const bool no_tiles = (flags & 256) != 0;
  // FSB_CG_ROTATE=1: rotate the CTA -> tile assignment from round to round (TileWalk::rot).  Measured
  // at 4096^2: no gain (69.8 vs 69.2 us per iteration) -- the 11 us between the first and the last CTA
  // reaching the barrier are not the wall tiles' slower path -- so it is off by default.
  const int rot = (flags & 16384) ? 1 : 0;
  const int n_walk = no_tiles ? 0 : tile_list ? s->n_active_tiles : n_tiles;
  const int n_prefix = tile_list ? s->n_prefix_tiles : 0;
  const int G = (int)gridDim.x;

  if (warp == NW)
  {
    // ---- producer: one elected lane feeds the ring, sweep after sweep
    if (lane != 0 || ss.done) return;
    RingPos rp = {0, 0};
    int cur = 0, sweep = -1;
    unsigned phase_id = 0;
    int npre = 0; // leading tiles of the sweep whose loads are already out
    // The next sweep's loads depend on the other CTAs' stores of this sweep -- that is, on every
    // CTA having ARRIVED at the barrier -- but not on the scalars: they go out as soon as the
    // arrival counter is complete, while warp 0 still folds the partials and advances the
    // scalars.  (Sharded: the peers' boundary rows are only known to have landed once their
    // mailbox entries are in, so tiles next to a slab boundary wait for the release.)
    const bool early = (flags & 4) != 0;
    // sharded: tiles that read a neighbour's ghost rows wait for the release (the peers' entries)
    auto needs_peer = [&](const TileWalk& t) {
      const int j0 = sh.row_lo + t.ty * TH;
      return sh.world > 1 && ((sh.rank > 0 && t.ty == 0) || (sh.rank < sh.world - 1 && j0 + TH + 2 > sh.row_hi));
    };
    // FSB_CG_PHINT=1: the old direction and residual are dead once this sweep has read them (the next
    // sweep overwrites them): evict-first, so that the L2 keeps what this sweep WRITES, which is what
    // the next sweep reads
    const bool dead_hint = (flags & 8) != 0;
    const uint64_t pol_dead = l2_policy_evict_first();
    auto issue = [&](const TileWalk& t, int cr) {
      if (rp.round > 0) mbar_wait_guarded(&empty[rp.st], (rp.round - 1) & 1);
      unsigned char* base = smem + rp.st * St::kBytes;
      const int c0 = t.tx * kTileW, j0 = sh.row_lo + t.ty * TH;
      mbar_expect_tx(&full[rp.st], St::kTx);
      if (dead_hint)
      {
        tma_load_2d_hint(base + St::oP, &maps.halo_p[cr], c0 - 4, j0 - 2, &full[rp.st], pol_dead);
        tma_load_2d_hint(base + St::oR, &maps.halo_r[cr], c0 - 4, j0 - 1, &full[rp.st], pol_dead);
      }
      else
      {
        tma_load_2d(base + St::oP, &maps.halo_p[cr], c0 - 4, j0 - 2, &full[rp.st]);
        tma_load_2d(base + St::oR, &maps.halo_r[cr], c0 - 4, j0 - 1, &full[rp.st]);
      }
      tma_load_2d(base + St::oC, &maps.code, c0 - 16, j0 - 1, &full[rp.st]);
      rp.advance(stages);
    };
    for (;;)
    {
      TileWalk t(blockIdx.x, G, tiles_x, n_walk, serp && (sweep & 1) == 0, tile_list, n_prefix, rot);
      int k = 0;
      for (; k < npre; ++k) t.next();
      for (; k < t.count; ++k, t.next()) issue(t, cur);
      ++phase_id;
      npre = 0;
      const RingPos pre = rp;
      if (early)
      {
        const unsigned int target = phase_id * (unsigned int)G;
        {
          SpinGuard g;
          while (ld_acquire_gpu(&s->bar_count) < target) g.tick();
        }
        fence_proxy_async_all();
        TileWalk tn(blockIdx.x, G, tiles_x, n_walk, serp && ((sweep + 1) & 1) == 0, tile_list, n_prefix, rot);
        const int want = min(stages, tn.count);
        for (; npre < want && !needs_peer(tn); ++npre, tn.next()) issue(tn, cur ^ 1);
      }
      {
        SpinGuard g;
        while (ss.released < phase_id) g.tick();
      }
      fence_proxy_async_all();
      if (ss.done)
      {
        // nobody will consume the tiles already requested: let them land before the CTA may exit
        RingPos w = pre;
        for (int m = 0; m < npre; ++m)
        {
          mbar_wait_guarded(&full[w.st], w.round & 1);
          w.advance(stages);
        }
        return;
      }
      cur ^= 1;
      ++sweep;
    }
  }

  // ---- consumers
  const float inv5 = coef.invdiag[4], diag5 = coef.diag[4], off = coef.off;
  const int r0 = (int)warp * RPW; // first own tile row
  // stage-relative offsets of this lane (floats / bytes): P box row r0 + i, R / code box row r0 + m
  const int fo = r0 * kHaloW + 4 + (int)lane * 4;
  const int co = r0 * kCodeW + 16 + (int)lane * 4;
  // lanes 0 / 31 also own the west / east halo column: P / R box columns 3 (and 2 beyond it) for
  // lane 0, 132 (and 133) for lane 31
  const bool edge = (lane == 0 || lane == 31);
  const bool west = (lane == 0);
  const int hfo = r0 * kHaloW + (west ? 3 : 4 + kTileW);     // the halo column itself
  const int hff = r0 * kHaloW + (west ? 2 : 4 + kTileW + 1); // the column beyond it
  const int hco = r0 * kCodeW + (west ? 15 : 16 + kTileW);
  const bool sharded = sh.world > 1;
  RingPos rp = {0, 0};
  int cur = 0, sweep = -1;
  unsigned phase_id = 0;
  float alpha_prev = 0.0f;       // alpha of the previous iteration (pending x update)
  bool pending = false;          // x lacks the update of the last iteration
  const float* last_p = nullptr; // direction of the last iteration
  while (!ss.done)
  {
    const float alpha = ss.alpha, beta = ss.beta, nalpha = -alpha;
    const bool iter_sweep = sweep >= 0;
    const bool with_x = iter_sweep && (!xdefer || (sweep & 1));
    const bool two_x = with_x && xdefer; // the previous iteration's update is still pending
    const float* __restrict__ p_old = cur ? p_b : p_a;
    float* __restrict__ p_new = cur ? p_a : p_b;
    float* __restrict__ r_new = cur ? r_a : r_b;
    double acc[kNSums];
#pragma unroll
    for (int n = 0; n < kNSums; ++n) acc[n] = no_tiles ? 1.0 : 0.0;
    bool pushed = false;
    TileWalk t(blockIdx.x, G, tiles_x, n_walk, serp && (sweep & 1) == 0, tile_list, n_prefix, rot);
    for (int tk = 0; tk < t.count; ++tk, t.next())
    {
      const unsigned char* base = smem + rp.st * St::kBytes;
      const float* sp = reinterpret_cast<const float*>(base + St::oP);
      const float* sr = reinterpret_cast<const float*>(base + St::oR);
      const unsigned char* sc = base + St::oC;
      const int ci = t.tx * kTileW + (int)lane * 4;
      const int j0 = sh.row_lo + t.ty * TH;
      const int jb = j0 + r0;
      // a tile holding one of the slab's two first / last rows stores into peer memory
      const bool push_tile = sharded && (t.ty == 0 || j0 + TH >= sh.row_hi - 1);
      pushed |= push_tile;
      const bool inside = (jb + RPW <= sh.row_hi) && (ci < ld); // all own rows of this lane are stored
      const size_t o0 = (size_t)jb * ld + ci;

      // x and the previous direction do not pass through the ring: they are read and written by
      // this thread only, so the loads are simply issued before the wait for the tile
      float4 xv[RPW], pv[RPW];
#pragma unroll
      for (int k = 0; k < RPW; ++k)
      {
        xv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (with_x && jb + k < sh.row_hi && ci < ld)
        {
          if (xhint)
          {
            // FSB_CG_XHINT=1: x (and the previous direction, read once more) are pure streams
            xv[k] = __ldcs(reinterpret_cast<const float4*>(x + o0 + (size_t)k * ld));
            if (two_x) pv[k] = __ldcs(reinterpret_cast<const float4*>(p_new + o0 + (size_t)k * ld));
          }
          else
          {
            xv[k] = *reinterpret_cast<const float4*>(x + o0 + (size_t)k * ld);
            if (two_x) pv[k] = *reinterpret_cast<const float4*>(p_new + o0 + (size_t)k * ld); // p_{k-1}
          }
        }
      }
      // the NEXT tile's x and previous-direction rows into the L2 now: its loads, one tile from here,
      // find them there instead of in DRAM (no registers held in between)
      if (with_x && tk + 1 < t.count && lane < 4 * RPW * (two_x ? 2 : 1))
      {
        const int ntx = t.list ? (t.ahead & 0xffff) : t.tx, nty = t.list ? (t.ahead >> 16) : t.ty;
        const int row = (int)lane >> 2 & (RPW - 1), which = (int)lane / (4 * RPW);
        const int nj = sh.row_lo + nty * TH + r0 + row, nci = ntx * kTileW + ((int)lane & 3) * 32;
        if (t.list && nj < sh.row_hi && nci < ld)
        {
          const float* a = (which ? p_new : x) + (size_t)nj * ld + nci;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
      mbar_wait(&full[rp.st], rp.round & 1);

      float4 pk[RPW + 4], rk[RPW + 2];
      uint32_t cd[RPW + 2];
      float hpc[RPW + 2]; // lanes 0 / 31: p_old of the halo-column cell, rows of the stage-1 block
      float hfar[RPW], hr[RPW];
      uint32_t hcd[RPW];
#pragma unroll
      for (int i = 0; i < RPW + 4; ++i) pk[i] = *reinterpret_cast<const float4*>(sp + fo + i * kHaloW);
      bool ok = inside && !push_tile && (flags & 512) == 0; // bit 9 (FSB_CG_DEBUG_NOFAST): table path only
#pragma unroll
      for (int m = 0; m < RPW + 2; ++m)
      {
        rk[m] = *reinterpret_cast<const float4*>(sr + fo + m * kHaloW);
        cd[m] = *reinterpret_cast<const uint32_t*>(sc + co + m * kCodeW);
        hpc[m] = edge ? sp[hfo + (m + 1) * kHaloW] : 0.0f;
        ok &= (cd[m] == kInterior4);
      }
#pragma unroll
      for (int k = 0; k < RPW; ++k)
      {
        hfar[k] = 0.0f;
        hr[k] = 0.0f;
        hcd[k] = 5u;
        if (edge)
        {
          hfar[k] = sp[hff + (k + 2) * kHaloW];
          hr[k] = sr[hfo + (k + 1) * kHaloW];
          hcd[k] = sc[hco + (k + 1) * kCodeW];
        }
        ok &= (hcd[k] == 5u);
      }
      // The ring slot is released only after everything staged has ARRIVED in registers, not merely been
      // requested: an mbarrier arrive does not wait for the warp's shared-memory loads still in flight,
      // and a slot released with loads outstanding can be refilled under them.  (Seen on B200 as a
      // timing-dependent last-bit wobble of the sums that involve the halo-column cells, whose loads are
      // the last ones issued.)  The arrive is therefore made data-dependent on every load of the tile:
      // one word of each, OR-ed together and folded over the warp.
      {
        uint32_t dep = 0;
#pragma unroll
        for (int i = 0; i < RPW + 4; ++i) dep |= __float_as_uint(pk[i].w) | __float_as_uint(pk[i].x);
#pragma unroll
        for (int m = 0; m < RPW + 2; ++m)
          dep |= __float_as_uint(rk[m].w) | __float_as_uint(rk[m].x) | cd[m] | __float_as_uint(hpc[m]);
#pragma unroll
        for (int k = 0; k < RPW; ++k) dep |= __float_as_uint(hfar[k]) | __float_as_uint(hr[k]) | hcd[k];
        dep = __reduce_or_sync(0xffffffffu, dep); // every lane's loads, and the warp converges here
        if (lane == 0) mbar_arrive_after(&empty[rp.st], dep);
      }
      rp.advance(stages);
      // FAST: every cell this warp touches in this tile (own rows, the rows above and below, the
      // halo-column cells) is LIQUID with four non-SOLID neighbours, all own rows lie inside the
      // slab and the grid, no peer stores: register constants, no table, no bounds logic
      const bool fast = __all_sync(0xffffffffu, ok);

      // x += alpha p (odd iterations: the pending update of the previous iteration first, the same
      // rounding order as one update per iteration)
      if (with_x)
      {
#pragma unroll
        for (int k = 0; k < RPW; ++k)
          if (jb + k < sh.row_hi && ci < ld)
          {
            float4 xn = xv[k];
            if (two_x) xn = fma4(alpha_prev, pv[k], xn);
            xn = fma4(alpha, pk[k + 2], xn);
            if (xhint) __stcs(reinterpret_cast<float4*>(x + o0 + (size_t)k * ld), xn);
            else *reinterpret_cast<float4*>(x + o0 + (size_t)k * ld) = xn;
          }
      }

      auto body = [&](auto fast_c) {
        constexpr bool FAST = decltype(fast_c)::value;
        // the update of iteration k for the halo-column cell of the own rows (lanes 0 / 31), with
        // the same association of the sums as apply_a4 / direction4, so that it equals bit for bit
        // what the owner of that cell stores:  q = A p, r' = r - alpha q, p' = D^-1 r' + beta p
        float he[RPW];
#pragma unroll
        for (int k = 0; k < RPW; ++k)
        {
          he[k] = 0.0f;
          if (edge)
          {
            const int m = k + 1;
            const float pc = hpc[m];
            const float pw = west ? hfar[k] : pk[m + 1].w;
            const float pe = west ? pk[m + 1].x : hfar[k];
            float4 kf = make_float4(inv5, diag5, off, 0.f);
            if (!FAST) kf = lut[hcd[k]];
            const float q = fmaf(kf.y, pc, kf.z * ((pw + pe) + (hpc[m - 1] + hpc[m + 1])));
            const float rr = fmaf(nalpha, q, hr[k]);
            he[k] = fmaf(beta, pc, kf.x * rr);
          }
        }
        // the same update on the warp's rows and one row above / below (stage-1 row m = grid row jb - 1 + m)
        float4 rn[RPW + 2], pn[RPW + 2];
#pragma unroll
        for (int m = 0; m < RPW + 2; ++m)
        {
          const float4 pc = pk[m + 1];
          float w = __shfl_up_sync(0xffffffffu, pc.w, 1);
          float e = __shfl_down_sync(0xffffffffu, pc.x, 1);
          if (edge) // lane 0: west neighbour, lane 31: east neighbour = the halo-column cell
          {
            if (west) w = hpc[m];
            else e = hpc[m];
          }
          const uint32_t c4 = FAST ? kInterior4 : cd[m];
          const float4 q = apply_a4(pc, w, e, pk[m], pk[m + 2], c4, lut, diag5, off);
          rn[m] = fma4(nalpha, q, rk[m]);
          pn[m] = direction4(rn[m], pc, c4, lut, inv5, beta);
        }
        // q' = A p' on the own rows, the five sums, the stores
#pragma unroll
        for (int k = 0; k < RPW; ++k)
        {
          const int m = k + 1;
          float w = __shfl_up_sync(0xffffffffu, pn[m].w, 1);
          float e = __shfl_down_sync(0xffffffffu, pn[m].x, 1);
          if (edge)
          {
            if (west) w = he[k];
            else e = he[k];
          }
          const uint32_t c4 = FAST ? kInterior4 : cd[m];
          const float4 q2 = apply_a4(pn[m], w, e, pn[m - 1], pn[m + 1], c4, lut, diag5, off);
          const int j = jb + k;
          if (FAST || (j < sh.row_hi && ci < ld))
          {
            const size_t o = o0 + (size_t)k * ld;
            *reinterpret_cast<float4*>(p_new + o) = pn[m];
            *reinterpret_cast<float4*>(r_new + o) = rn[m];
            if (!FAST && push_tile)
            {
              // the slab's first / last two rows of p and first / last row of r also go straight
              // into the neighbours' ghost rows (NVLink stores)
              float* const pp_lo = peers.p_lo[cur ^ 1];
              float* const pp_hi = peers.p_hi[cur ^ 1];
              float* const pr_lo = peers.r_lo[cur ^ 1];
              float* const pr_hi = peers.r_hi[cur ^ 1];
              if (pp_lo && j <= sh.row_lo + 1) *reinterpret_cast<float4*>(pp_lo + o) = pn[m];
              if (pp_hi && j >= sh.row_hi - 2) *reinterpret_cast<float4*>(pp_hi + o) = pn[m];
              if (pr_lo && j == sh.row_lo) *reinterpret_cast<float4*>(pr_lo + o) = rn[m];
              if (pr_hi && j == sh.row_hi - 1) *reinterpret_cast<float4*>(pr_hi + o) = rn[m];
            }
            float4 iv = make_float4(inv5, inv5, inv5, inv5);
            if (!FAST && c4 != kInterior4)
              iv = make_float4(lut[c4 & 0xff].x, lut[(c4 >> 8) & 0xff].x, lut[(c4 >> 16) & 0xff].x,
                               lut[c4 >> 24].x);
            const float4 z = make_float4(iv.x * rn[m].x, iv.y * rn[m].y, iv.z * rn[m].z, iv.w * rn[m].w);
            const float4 mq = make_float4(iv.x * q2.x, iv.y * q2.y, iv.z * q2.z, iv.w * q2.w);
            acc[0] += (double)dot4(pn[m], q2);
            acc[1] += (double)dot4(rn[m], z);
            acc[2] += (double)dot4(rn[m], rn[m]);
            acc[3] += (double)dot4(z, q2);
            acc[4] += (double)dot4(q2, mq);
          }
        }
      };
      if (fast) body(std::true_type{});
      else body(std::false_type{});
    }
    ++phase_id;
    one_reduce<NW>(acc, s, partials + (phase_id & 1u) * (8 * G), phase_id, sh, &ss, pushed, (flags >> 11) & 7, dbg);
    if (iter_sweep)
    {
      alpha_prev = alpha;
      pending = !with_x;
      last_p = p_old;
    }
    cur ^= 1;
    ++sweep;
  }
  if (pending && !ss.comm_error)
  {
    // the solve ended on an iteration whose x update was deferred: x += alpha p of that iteration.
    // Same tile -> thread mapping as the sweeps; p is exactly zero outside LIQUID cells.
    TileWalk t(blockIdx.x, G, tiles_x, n_walk, false, tile_list, n_prefix, rot);
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
          const float4 x4 = *reinterpret_cast<const float4*>(x + o);
          *reinterpret_cast<float4*>(x + o) = fma4(alpha_prev, p4, x4);
        }
      }
    }
  }
  // every CTA holds the same final scalars; CTA 0 publishes them for the host
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    s->r2 = ss.r2;
    s->rz = ss.rz;
    s->iter = ss.iter;
    s->done = ss.done;
    s->seq[3] = ss.seq;
    if (ss.comm_error) s->comm_error = 1;
  }
}

template <int RPW>
int configure_one_shape(fsb_ctx* c, int64_t n_tiles)
{
  constexpr int TH = kNWOne * RPW;
  const int threads = (kNWOne + 1) * 32;
  const int budget = (227 * 1024 - 2 * 2048) / 2; // two resident CTAs per SM
  // Ring depth: 3.  Deeper rings fit (5 stages of 21 KB at two CTAs per SM) but measured slower -- one
  // B200, same box: 4096^2 76.7 (5 stages) / 74.8 (4) / 72.7 (3) / 73.3 (2) us per iteration, the
  // 8192 x 1024 slab of an 8-GPU solve 42.0 / 39.2 / 39.9 (5 / 3 / 2), 8192^2 291.9 / 279.1 (5 / 3): the
  // shared memory a shorter ring leaves free goes to the L1, through which x and the spilled registers pass.
  int stages = std::max(2, std::min(3, budget / OneStage<TH>::kBytes));
  if (const char* e = getenv("FSB_CG_STAGES")) // tuning knob for profiling runs
  {
    const int v = atoi(e);
    if (v >= 2 && v <= kMaxStages) stages = v;
  }
  const int smem = stages * OneStage<TH>::kBytes;
  if (smem > 227 * 1024 - 2048)
    return fsb_fail(c, FSB_ERR_INVALID, "CG ring of %d stages does not fit shared memory", stages);
  auto kf = k_cg_solve1<kNWOne, RPW>;
  FSB_CUDA(c, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 1;
  FSB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, threads, smem));
  int cap = 2;
  if (const char* e = getenv("FSB_CG_CTAS_PER_SM")) cap = std::max(1, atoi(e));
  occ = std::max(1, std::min(occ, cap));
  c->cg_one_stages = stages;
  c->cg_one_grid = (int)std::min<int64_t>(n_tiles, (int64_t)c->sm_count * occ);
  return FSB_OK;
}

int configure_one(fsb_ctx* c)
{
  const int th = c->cg_tile_rows; // chosen by configure_cg (the active-tile list uses it too)
  if (c->cg_one_th == th) return FSB_OK;
  const int64_t n_tiles = (int64_t)fsb_div_up(c->ld, kTileW) *
                          fsb_div_up(c->shard.row_hi - c->shard.row_lo, th);
  if (th == 4 * kNWOne) FSB_TRY(configure_one_shape<4>(c, n_tiles));
  else if (th == 2 * kNWOne) FSB_TRY(configure_one_shape<2>(c, n_tiles));
  else FSB_TRY(configure_one_shape<1>(c, n_tiles));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FSB_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fsb_fail(c, FSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  OneMaps* m = reinterpret_cast<OneMaps*>(c->cg_maps_one);
  memset(m, 0, sizeof(OneMaps));
  FSB_TRY(make_map(c, encode, &m->halo_p[0], c->cg_p[0], true, kHaloW, th + 4));
  FSB_TRY(make_map(c, encode, &m->halo_p[1], c->cg_p[1], true, kHaloW, th + 4));
  FSB_TRY(make_map(c, encode, &m->halo_r[0], c->cg_r, true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &m->halo_r[1], c->cg_r2, true, kHaloW, th + 2));
  FSB_TRY(make_map(c, encode, &m->code, c->cg_code, false, kCodeW, th + 2));
  c->cg_one_th = th;
  return FSB_OK;
}

} // namespace

int fsb_cg_one_partials(const fsb_ctx* c) { return 2 * 8 * std::max(c->cg_one_grid, c->sm_count * 2); }

// Called after the set-up (k_cg_build*: x = 0, r = b, stencil codes, |b|^2 and the threshold in
// the CG scalars, the active-tile list) with the solve not yet finished.
int fsb_k_cg_one_solve(fsb_ctx* c, const CgCoef& coef)
{
  FSB_TRY(configure_one(c));
  const int th = c->cg_tile_rows;
  const ShardArgs& sh = c->shard;
  int tiles_x = fsb_div_up(c->ld, kTileW);
  int n_tiles = tiles_x * fsb_div_up(sh.row_hi - sh.row_lo, th);
  const int need = 2 * 8 * c->cg_one_grid;
  if (need > c->partials_cap)
    return fsb_fail(c, FSB_ERR_INVALID, "partials buffer too small for the one-sweep solve");
  {
    // The sweeps rely on both direction buffers and the second residual buffer being exactly zero
    // wherever no LIQUID cell is (skipped tiles, masked cells) and on p = 0 before the first
    // iteration: clear this rank's OWN rows of them.  The neighbours' ghost rows of the buffers
    // the set-up sweep WRITES (p[1], r2) belong to the peers, which may already be storing into
    // them; those of p[0] are read by the set-up sweep as "p = 0" and no peer writes them before
    // this rank has finished that sweep, so they are cleared here as well.
    const int lo = sh.world > 1 ? sh.row_lo : 0;
    const int hi = sh.world > 1 ? sh.row_hi : c->ny;
    const size_t off = (size_t)lo * c->ld, bytes = sizeof(float) * (size_t)(hi - lo) * c->ld;
    const int glo = std::max(0, lo - 2), ghi = std::min(c->ny, hi + 2);
    FSB_CUDA(c, cudaMemsetAsync(c->cg_p[0] + (size_t)glo * c->ld, 0,
                                sizeof(float) * (size_t)(ghi - glo) * c->ld, c->stream));
    FSB_CUDA(c, cudaMemsetAsync(c->cg_p[1] + off, 0, bytes, c->stream));
    FSB_CUDA(c, cudaMemsetAsync(c->cg_r2 + off, 0, bytes, c->stream));
  }
  const OneMaps& maps = *reinterpret_cast<const OneMaps*>(c->cg_maps_one);
  OnePeers peers;
  const bool south = sh.world > 1 && sh.rank > 0, north = sh.world > 1 && sh.rank < sh.world - 1;
  for (int k = 0; k < 2; ++k)
  {
    peers.p_lo[k] = south ? c->peer_p[k][sh.rank - 1] : nullptr;
    peers.p_hi[k] = north ? c->peer_p[k][sh.rank + 1] : nullptr;
  }
  peers.r_lo[0] = south ? c->peer_r[sh.rank - 1] : nullptr;
  peers.r_hi[0] = north ? c->peer_r[sh.rank + 1] : nullptr;
  peers.r_lo[1] = south ? c->peer_r2[sh.rank - 1] : nullptr;
  peers.r_hi[1] = north ? c->peer_r2[sh.rank + 1] : nullptr;
  int ld = c->ld, stages = c->cg_one_stages;
  CgCoef cf = coef;
  int flags = c->cg_flags;
  // FSB_CG_DEBUG_SUMS=<file prefix> (measurement / debugging only): per sweep and CTA, the CTA's own five
  // sums and the totals as it sees them, for the first 64 sweeps, appended to <prefix>.<solve number>
  static int dbg_solves = 0;
  double* dbg = nullptr;
  const char* dbg_path = getenv("FSB_CG_DEBUG_SUMS");
  const size_t dbg_count = (size_t)64 * c->cg_one_grid * 10 + 8 + 64 * 8;
  if (dbg_path)
  {
    FSB_CUDA(c, cudaMalloc(&dbg, sizeof(double) * dbg_count));
    FSB_CUDA(c, cudaMemsetAsync(dbg, 0, sizeof(double) * dbg_count, c->stream));
  }
  void* args[] = {(void*)&maps, &c->cg_x, &c->cg_r, &c->cg_r2, &c->cg_p[0], &c->cg_p[1], &ld, &tiles_x,
                  &n_tiles, &stages, &cf, &c->scal, &c->partials, (void*)&sh, &peers, &flags, &dbg};
  const dim3 grid(c->cg_one_grid), block((kNWOne + 1) * 32);
  cudaError_t e;
  if (th == 4 * kNWOne)
    e = cudaLaunchCooperativeKernel((void*)k_cg_solve1<kNWOne, 4>, grid, block, args,
                                    (size_t)stages * OneStage<4 * kNWOne>::kBytes, c->stream);
  else if (th == 2 * kNWOne)
    e = cudaLaunchCooperativeKernel((void*)k_cg_solve1<kNWOne, 2>, grid, block, args,
                                    (size_t)stages * OneStage<2 * kNWOne>::kBytes, c->stream);
  else
    e = cudaLaunchCooperativeKernel((void*)k_cg_solve1<kNWOne, 1>, grid, block, args,
                                    (size_t)stages * OneStage<1 * kNWOne>::kBytes, c->stream);
  if (e != cudaSuccess)
    return fsb_fail(c, FSB_ERR_CUDA, "cooperative launch of the one-sweep CG solve failed: %s",
                    cudaGetErrorString(e));
  c->launches += 1;
  if (dbg)
  {
    std::vector<double> h(dbg_count);
    FSB_CUDA(c, cudaMemcpyAsync(h.data(), dbg, sizeof(double) * dbg_count, cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(dbg);
    char name[512];
    snprintf(name, sizeof name, "%s.%d", dbg_path, dbg_solves++);
    if (FILE* f = fopen(name, "wb"))
    {
      fwrite(h.data(), sizeof(double), dbg_count, f);
      fclose(f);
    }
  }
  return FSB_OK;
}
