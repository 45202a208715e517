// Tile frame shared by the CG solve kernels (fsb_cg.cu: two sweeps per iteration; fsb_cg_one.cu: one
// sweep per iteration): TMA / mbarrier wrappers, the tile walk, the per-float4 stencil arithmetic,
// the self-validating mailbox word and the spin watchdog.  Included by both translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"
#include "fsb_cg_math.cuh"

namespace {

// ------------------------------------------------------------ tile frame --
// Both iteration kernels are TMA-fed, warp-specialised, persistent kernels:
//   * CTA = NW consumer warps + 1 producer warp; tiles of kTileW = 128 columns
//     x TH = NW * RPW rows.  Consumer warp w owns the RPW consecutive tile rows
//     w*RPW ..; lane l owns columns 4l..4l+3 (one 16-byte shared-memory access).
//   * the producer's elected lane walks the CTA's tile list and issues
//     cp.async.bulk.tensor (TMA) box loads into a STAGES-deep shared-memory
//     ring, completion on full[stage] (mbarrier, expect_tx); a consumer warp
//     releases a stage (arrive on empty[stage]) as soon as it has pulled what it
//     needs into registers.  Boxes include the one-cell halo (fp32: 136 x (TH+2)
//     starting at (c0-4, j0-1); code bytes: 160 x (TH+2) starting at (c0-16,
//     j0-1)); TMA zero-fills everything outside the grid, so there is no bounds
//     logic on the load side, and (STAGES-1) tiles of loads stay in flight per
//     CTA regardless of what the consumer warps are doing.
//   * consumers never write shared memory and never synchronise with each
//     other inside the tile loop: k_cg_direction re-computes the new direction
//     on the two halo rows and two halo columns of a warp's row block from the
//     staged r / p_old / code instead of exchanging it between warps.
//   * coefficients: a float4 group whose four cells are all liquid with four
//     non-SOLID neighbours (code word 0x05050505, the bulk of any scene) uses
//     register constants; other groups read a shared-memory table of float4
//     {inverse diagonal, diagonal, off-diagonal, 0} indexed by the stencil code
//     (code 0 -> zeros, so masked cells come out exactly 0 without branches).
//     (An indexed kernel-parameter array compiles to indexed LDC on the XU
//     pipe: measured 78 % XU-bound, profiles/r01b.)
//   * arithmetic uses explicit FMA: the CG is held to the solver tolerance and
//     comparable iteration counts, not to Eigen's rounding.
constexpr int kTileW = 128;
constexpr int kHaloW = kTileW + 8;  // fp32 halo box width: columns c0-4 .. c0+131
constexpr int kCodeW = kTileW + 32; // code halo box width: columns c0-16 .. c0+143
constexpr int kMaxStages = 8;

__host__ __device__ constexpr int align128(int x) { return (x + 127) / 128 * 128; }

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
// Release / acquire fences.  __threadfence() is a SEQUENTIALLY CONSISTENT fence (SASS: MEMBAR.SC.GPU +
// CCTL.IVALL); the barrier protocols here only need release before an arrival and acquire after a
// wait, which fence.acq_rel provides (SASS: MEMBAR.ALL.GPU).
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Arrival / wait of the software grid barrier with the ordering attached to the operation itself
// instead of a separate fence on each side.
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v)
{
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_acq_rel_gpu_add(unsigned int* p, unsigned int v)
{
  unsigned int old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// An arrive that cannot be issued before the value `dep` is available: its address is formed from a
// term that is always zero (bit 1 of a square) but that the assembler cannot fold away.  Used to
// release a ring slot as soon as the shared-memory loads that produced `dep` have RETURNED -- an
// mbarrier arrive by itself does not wait for the warp's loads in flight.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep)
{
  asm volatile(
      "{\n"
      ".reg .b32 t;\n"
      "mul.lo.u32 t, %1, %1;\n"
      "and.b32 t, t, 2;\n"
      "add.u32 t, t, %0;\n"
      "mbarrier.arrive.shared::cta.b64 _, [t];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(dep)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      // (a suspend-time hint on this try_wait -- 20 us -- was measured to cost 10 % at 4096^2: the warp
      // wakes up late)
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
// programmatic dependent launch: let the next kernel of the stream be scheduled early / wait
// for the previous kernel's memory before touching anything it wrote
__device__ __forceinline__ void pdl_launch_dependents()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
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

// walk of the CTA's tile list (tiles blockIdx.x, + gridDim.x, ...) without a division per tile.
// `reverse` walks the SAME list from its last tile down: consecutive sweeps over the grid then
// meet at the rows the previous sweep touched last, which are the ones still in the L2.
struct TileWalk
{
  int tx, ty, step_x, step_y, tiles_x, count;
  bool rev;
  // optional list of ACTIVE tiles (packed ty << 16 | tx): the walk then runs over list positions
  // instead of tile numbers, so tiles without a LIQUID cell are never visited.  The first `prefix`
  // list entries (a slab's boundary-row tiles in a sharded solve) are visited FIRST in both walk
  // directions: their rows go to the neighbour GPUs while the rest of the sweep is still running,
  // so the system-scope fence before the reduction finds the peer stores already acknowledged.
  const int* list;
  int first, stride, n_total, k, n_pre;
  int ahead; // list entry of tile k + 1, requested one tile early (its latency is off the critical path)
  // rot != 0: round r of the walk gives this CTA the list position r * stride + (first + r * rot) %
  // stride instead of r * stride + first, so that a CTA's tiles do not all sit in the same few tile
  // columns (stride 296 over 32 tile columns: every fourth tile of CTAs 0, 7 mod 8 would be a wall
  // tile, which takes the slower table path).  Only without a prefix.
  int rot;
  __device__ __forceinline__ TileWalk(int first_tile, int stride_, int tiles_x_, int n_tiles,
                                      bool reverse = false, const int* list_ = nullptr,
                                      int prefix = 0, int rot_ = 0)
  {
    rot = (list_ && prefix == 0) ? rot_ : 0;
    tiles_x = tiles_x_;
    rev = reverse;
    list = list_;
    first = first_tile;
    stride = stride_;
    n_total = n_tiles;
    k = 0;
    ahead = 0;
    count = first_tile < n_tiles ? (n_tiles - 1 - first_tile) / stride_ + 1 : 0;
    if (rot)
    {
      const int full = n_tiles / stride_, rem = n_tiles - full * stride_;
      count = full + (((first_tile + full * rot) % stride_ < rem) ? 1 : 0);
    }
    // this CTA's list positions below `prefix`
    n_pre = (list && first_tile < prefix) ? (prefix - 1 - first_tile) / stride_ + 1 : 0;
    if (n_pre > count) n_pre = count;
    const int start = reverse ? first_tile + (count - 1) * stride_ : first_tile;
    step_x = stride_ % tiles_x;
    step_y = stride_ / tiles_x;
    tx = 0;
    ty = 0;
    if (list)
    {
      if (count > 0)
      {
        const int t = __ldg(list + position(0));
        tx = t & 0xffff;
        ty = t >> 16;
        if (count > 1) ahead = __ldg(list + position(1));
      }
    }
    else
    {
      tx = start % tiles_x;
      ty = start / tiles_x;
    }
  }
  // list position of the kk-th tile of this CTA: prefix entries ascending, then the rest in walk order
  __device__ __forceinline__ int position(int kk) const
  {
    if (kk < n_pre) return first + kk * stride;
    const int m = kk - n_pre; // m-th of the (count - n_pre) non-prefix entries
    const int r = rev ? (count - 1 - m) : (n_pre + m);
    if (rot) return r * stride + (first + r * rot) % stride;
    return first + r * stride;
  }
  __device__ __forceinline__ void next()
  {
    if (list)
    {
      ++k;
      if (k < count)
      {
        tx = ahead & 0xffff;
        ty = ahead >> 16;
        if (k + 1 < count) ahead = __ldg(list + position(k + 1));
      }
      return;
    }
    if (!rev)
    {
      tx += step_x;
      ty += step_y;
      if (tx >= tiles_x)
      {
        tx -= tiles_x;
        ++ty;
      }
    }
    else
    {
      tx -= step_x;
      ty -= step_y;
      if (tx < 0)
      {
        tx += tiles_x;
        --ty;
      }
    }
  }
};

// ---- peer-memory mailbox (row-slab sharding, see fsb_internal.cuh)
// Every 8-byte word of a slot is self-validating: 32 bits of payload + the low 32 bits of the
// sequence number.  8-byte stores are single transactions, so the reader needs no flag that is
// ordered after the payload and the writer needs no fence between them (a system-scope fence
// costs an NVLink round trip).  One double = two words.
__device__ __forceinline__ unsigned long long mail_word(unsigned int payload, unsigned long long seq)
{
  return (unsigned long long)payload | (seq << 32);
}

__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A peer that stays silent for kMailTimeoutNs raises comm_error and ends the solve instead of
// hanging the GPU.
constexpr unsigned long long kMailTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
// Watchdog of the persistent kernel's spin loops: a wait that outlives every legitimate cause
// (the mailbox time-out included) traps, so a protocol bug ends the launch with an error instead
// of occupying the GPU forever.
constexpr unsigned long long kHangNs = 45ull * 1000ull * 1000ull * 1000ull;
#ifndef FSB_SPIN_SLEEP_NS
#define FSB_SPIN_SLEEP_NS 32
#endif
constexpr unsigned kSpinSleepNs = FSB_SPIN_SLEEP_NS;
struct SpinGuard
{
  unsigned int n = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void tick()
  {
    // a spinning warp (the producer waiting for a ring slot, warp 0 and the producer waiting at the grid
    // barrier) takes issue slots from the consumer warps of its scheduler: 21 % of all executed warp
    // instructions of the solve kernel were these loops (profiles/r02_cg_solve1_ncu_full.md); back off
    __nanosleep(kSpinSleepNs);
    if ((++n & 4095u) == 0)
    {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kHangNs) __trap();
    }
  }
};
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity)
{
  SpinGuard g;
  while (!mbar_try(bar, parity)) g.tick();
}

__device__ __forceinline__ void fence_proxy_async_all()
{
  asm volatile("fence.proxy.async;" ::: "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// evict-last on a fraction q/4 of the accesses (q = 1..4); the fraction must be an immediate
__device__ __forceinline__ uint64_t l2_policy_evict_last(int q)
{
  uint64_t pol;
  if (q >= 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (q == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.75;" : "=l"(pol));
  else if (q == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.5;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.25;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1,
                                                 uint64_t* bar, uint64_t policy)
{
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void st_f4_hint(float* ptr, const float4 v, uint64_t policy)
{
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
               : "memory");
}

// ring slot bookkeeping shared by the producer and the consumers
struct RingPos
{
  int st, round;
  __device__ __forceinline__ void advance(int stages)
  {
    if (++st == stages)
    {
      st = 0;
      ++round;
    }
  }
};

// Kernel shapes: 8 consumer warps x RPW rows.  RPW = 2 (16-row tiles) when that still
// gives every SM several tiles, else RPW = 1 (8-row tiles, small grids).
constexpr int kNW = 8;
// the one-sweep kernel (fsb_cg_one.cu) needs more registers per thread: 7 consumer warps + the producer
// = 256 threads, 128 registers at two CTAs per SM
constexpr int kNWOne = 7;

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

} // namespace
