// Internal declarations shared by the libfsb.so translation units.
// Product code: never includes or links anything under oracle/.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fsb.h"

// ---------------------------------------------------------------------------
// Data layout in HBM
//   * every grid (u, v, prev, diff, CG vectors: fp32; labels, masks, stencil
//     codes: u8) is stored row-major with a padded row pitch `ld` (elements),
//     ld = nx rounded up to 32, so every row starts on a 128-byte boundary and
//     every kernel can use 16-byte vector accesses.  Pad columns hold zeros /
//     non-liquid codes and are never copied to the host.
//   * particles: AoS float4 {pos_x,pos_y,vel_x,vel_y} (one 16-byte access per
//     particle), kept SORTED by cell on the device, with `orig[k]` = index the
//     caller knows the particle by.  cell_start[c] .. cell_start[c+1] is the
//     range of cell c (dense index i + j*nx).
// ---------------------------------------------------------------------------

struct CgScalars
{
  double rhs2; // |b|^2
  double pq;   // p . A p
  double r2;   // |r|^2 after the update
  double rz;   // r . z
  float abs_new, abs_old, beta, thr, tol;
  int iter;      // Eigen's `i`
  int done;      // 1: converged or capped; all later CG launches are no-ops
  int max_iters;
  int n_liquid;
  unsigned int ticket[4]; // last-block counters, one per reducing kernel
  // row-slab sharding: sequence numbers of the peer-memory mailbox reductions (one per type),
  // and a flag raised when a peer did not answer in time
  unsigned long long seq[4];
  int comm_error;
  // where a mailbox wait gave up (diagnostics for FSB_ERR_COMM): sweep, missing rank, expected tag,
  // word seen
  unsigned long long comm_diag[4];
  // software grid barrier of the persistent solve kernel: arrivals so far / last released phase
  unsigned int bar_count;
  unsigned int bar_release;
  // active-tile list of the solve (tiles with at least one LIQUID cell); null: all tiles
  const int* tile_list;
  int n_active_tiles;
  int n_prefix_tiles; // leading list entries that every walk visits first (slab boundary rows)
};

// ---------------------------------------------------------------------------
// Row-slab sharding of the CG across GPUs (one process per GPU).
//   Every rank holds the full-size vectors; rank q iterates on rows
//   [row_lo, row_hi) and keeps the two ghost rows row_lo-1 and row_hi current:
//   the kernel that produces a boundary row (k_cg_direction: p, k_cg_update: r)
//   also stores it straight into the neighbour's copy over NVLink (peer pointers
//   from CUDA IPC).  The two dot products per iteration are combined through a
//   mailbox in each rank's HBM: the last CTA of the producing kernel writes its
//   partial sums and a sequence number into EVERY rank's mailbox, a one-warp
//   combine kernel waits for all `world` entries and adds them in rank order, so
//   all ranks derive bit-identical alpha / beta / `done`.  No NCCL call in the loop.
constexpr int kMaxRanks = 8;
constexpr int kMailTypes = 12; // 0: p.Ap, 1: (|r|^2, r.z), 2: end-of-solve barrier; 4-11: one-sweep solve
// the one-sweep solve (fsb_cg_one.cu) posts five doubles per reduction: rank q's entry of reduction
// number `seq` is the 16 self-validating words starting at word
// kOneMailWord + ((seq & 1) * kMaxRanks + q) * kOneMailStride of a mailbox.  Two entries per rank,
// used alternately: a rank can only post reduction n + 2 after it has received every rank's n + 1,
// and a rank posts n + 1 only after all of its CTAs have finished reading the entries of n -- so an
// entry is never overwritten while somebody may still have to read it.
constexpr int kOneMailWord = 4 * kMaxRanks * 4;
constexpr int kOneMailStride = 16;

struct MailSlot
{
  double v[2];
  unsigned long long seq;
  unsigned long long pad;
};

struct ShardArgs
{
  int world, rank;
  int row_lo, row_hi;         // rows this rank iterates on
  MailSlot* mail[kMaxRanks];  // mail[q]: rank q's mailbox (kMailTypes x kMaxRanks slots)
};

struct CgCoef
{
  float off;        // (float)(1 / pow(dx, 2))           src/FluidSolver.cpp:382
  float diag[5];    // (float)(-n / pow(dx, 2)), n=0..4  src/FluidSolver.cpp:409-410
  float invdiag[5]; // Eigen DiagonalPreconditioner: diag != 0 ? 1/diag : 1
};

struct fsb_mg_state;  // multigrid hierarchy of the opt-in preconditioner (fsb_mg.cu)

struct fsb_ctx
{
  int nx = 0, ny = 0, ld = 0;
  float dx = 0, dy = 0;           // MacGrid deltas (src/MacGrid.cpp:8)
  float pool_dx = 0, pool_dy = 0; // FluidSolverMemoryPool deltas (src/FluidSolver.cpp:56-65)
  int pool_nx = 0, pool_ny = 0;
  float density = 0, pic_ratio = 0;
  float grav_x = 0, grav_y = 0;
  int integrator = FSB_INTEGRATOR_RK3;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  // MacGrid
  float* u[2] = {nullptr, nullptr};
  float* v[2] = {nullptr, nullptr};
  int front = 0;
  float *u_prev = nullptr, *v_prev = nullptr, *u_diff = nullptr, *v_diff = nullptr;
  uint8_t* cell = nullptr;
  // FluidSolverMemoryPool: extension masks
  uint8_t* mask_x[2] = {nullptr, nullptr};
  uint8_t* mask_y[2] = {nullptr, nullptr};
  int mask_front = 0;

  // particles (double-buffered for the cell sort)
  float4* part[2] = {nullptr, nullptr};
  int* orig[2] = {nullptr, nullptr};
  int pcur = 0;
  int64_t n = 0, cap = 0;
  bool sort_valid = false;
  int* cell_start = nullptr; // nx*ny + 1
  int* cell_count = nullptr; // nx*ny
  int* sort_key = nullptr;   // cap
  int* sort_rank = nullptr;  // cap
  int* sort_idx = nullptr;   // cap
  int* scan_block = nullptr; // block sums for the scan
  float* stage = nullptr;    // device staging for host copies (dense), grown on demand
  size_t stage_bytes = 0;

  // CG
  float *cg_x = nullptr, *cg_r = nullptr;
  float* cg_r2 = nullptr; // second residual buffer of the one-sweep solve (r is ping-ponged there)
  float* cg_p[2] = {nullptr, nullptr}; // search direction, ping-pong
  uint8_t* cg_code = nullptr;
  double* partials = nullptr;
  int partials_cap = 0;
  CgScalars* scal = nullptr;   // device
  CgScalars* scal_h = nullptr; // pinned host mirror, two slots (poll runs one chunk behind)
  cudaEvent_t cg_ev[2] = {nullptr, nullptr};
  cudaGraphExec_t cg_graph = nullptr;
  int cg_graph_state = 0; // 0: not built, 1: usable, -1: capture unavailable (direct launches)
  int cg_tile_rows = 0, cg_grid_dir = 0, cg_grid_upd = 0, cg_stages_dir = 0, cg_stages_upd = 0;
  // TMA descriptors of the two iteration kernels for both ping-pong phases (CgMaps in fsb_cg.cu)
  alignas(64) unsigned char cg_maps_dir[2][5 * sizeof(CUtensorMap)];
  alignas(64) unsigned char cg_maps_upd[2][5 * sizeof(CUtensorMap)];
  alignas(64) unsigned char cg_maps_fused[8 * sizeof(CUtensorMap)]; // SolveMaps
  alignas(64) unsigned char cg_maps_one[5 * sizeof(CUtensorMap)]; // OneMaps (fsb_cg_one.cu)
  int cg_one_th = 0, cg_one_grid = 0, cg_one_stages = 0;
  bool cg_one = true;    // one sweep + one reduction per iteration (default; fsb_cg_one.cu)
  bool last_solve_one = false;
  bool cg_fused = false; // persistent cooperative solve kernel in use
  bool cg_pdl = true;    // programmatic dependent launch between the iteration kernels
  int cg_flags = 0;      // tuning bits of the iteration kernels, see configure_cg
  int cg_persist_mb = 0; // L2 set-aside for the residual during a solve (0: none)
  // preconditioner: FSB_PRECOND_JACOBI (the reference's, default) or FSB_PRECOND_MULTIGRID (opt-in)
  int precond = 0;
  int mg_max_iters = 200;    // beyond this the multigrid iteration is abandoned for Jacobi
  int mg_sweeps = 3;         // damped-Jacobi pre- and post-sweeps per level (equal: symmetric V-cycle)
  bool mg_graph = true;   // FSB_MG_GRAPH=0: the V-cycle as plain launches instead of one captured CUDA graph
  int mg_stop = 4;        // FSB_MG_STOP: coarsen until both sides are <= this (kMgStopDefault; at most 32)
  bool mg_renorm = true;  // FSB_MG_RENORM=0: plain bilinear / full-weighting transfers (round-1 behaviour)
  bool last_solve_mg = false;
  fsb_mg_state* mg = nullptr;
  bool cg_persist_miss_normal = false;
  bool cg_skip_tiles = true; // sweeps visit only tiles that hold a LIQUID cell
  bool cg_edge_first = false; // sharded solves: slab boundary tiles first in every sweep (knob)
  int* cg_tile_flags = nullptr;
  int* cg_tile_list = nullptr;
  int cg_tile_cap = 0;
  int cg_last_active_tiles = 0; // active-tile count of the last solve (measurement)
  int cg_grid_fused = 0, cg_fused_stages = 0, cg_fused_stage_bytes = 0;
  int max_iters = 100;
  float tol = 1.1920929e-7f;
  int iters = 0;
  float err = 0;
  bool pressure_valid = false;
  // fused FLIP steps interpolate (front - previous) per tap; the diff buffer is
  // materialised only when somebody reads it
  bool diff_pending = false;

  // semi-Lagrangian velocity advection (fsb_sl.cu): one 16-byte record per face, largest displacement
  void *sl_rec_u = nullptr, *sl_rec_v = nullptr;
  int* sl_maxd = nullptr;
  size_t sl_cells = 0;
  int sl_reach = 0;         // reach (source cells) of the last gather
  int canon_per = 2;            // FSB_CANON_PER: cells per thread of the sort's in-cell ordering pass (1, 2)
  int p2g_pipe = 1;             // FSB_P2G_PIPE: particle-to-grid kernel requests the next record before computing this one
  int sort_per = 2;             // FSB_SORT_PER: particles per thread of the sort's counting and gather passes (1, 2, 4)
  int g2p_per = 2;              // FSB_G2P_PER: particles per thread of the grid-to-particle kernel (1, 2, 4)
  int build_blocks_per_sm = 0;  // grid of the pressure set-up kernel: 0 = one resident wave (FSB_BUILD_BLOCKS_PER_SM)
  bool sl_atomic = false;   // FSB_SL_ATOMIC=1: the first-generation float-atomics scatter

  // row-slab sharding (fsb_shard_*); world == 1: not sharded
  ShardArgs shard = {1, 0, 0, 0, {nullptr}};
  MailSlot* mail_local = nullptr;
  float* peer_r[kMaxRanks] = {nullptr};
  float* peer_r2[kMaxRanks] = {nullptr};
  float* peer_p[2][kMaxRanks] = {{nullptr}};
  float* peer_x[kMaxRanks] = {nullptr};
  float** peer_x_dev = nullptr; // device array of the world-1 peer x pointers
  void* ipc_opened[6 * kMaxRanks] = {nullptr};
  int n_ipc_opened = 0;

  // slab-partitioned particles (fsb_slab_*): world == 1: the whole set lives here
  int slab_world = 1, slab_rank = 0, slab_lo = 0, slab_hi = 0;
  int64_t slab_count[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // group sizes after fsb_slab_sort_out
  bool slab_grouped = false;
  bool slab_partitioned = false; // the set holds a slab's particles with GLOBAL ids (after keep_own / add)
  float4* slab_buf_part = nullptr; // boundary-row selection
  int* slab_buf_orig = nullptr;
  int64_t slab_buf_cap = 0;
  int64_t slab_sel = 0;                   // size of the last boundary-row selection
  unsigned long long* slab_ctr = nullptr; // device counters / cursors (2 * kMaxRanks)

  // measurement
  bool profiling = false;
  static constexpr int kProfPool = 512; // event pairs recorded between two drains
  cudaEvent_t prof_ev[kProfPool][2];
  int prof_stage[kProfPool];
  int prof_used = 0;
  bool prof_made = false;
  float prof_ms[FSB_PROF_COUNT];
  int prof_calls[FSB_PROF_COUNT];
  cudaEvent_t timer_ev[2] = {nullptr, nullptr};
  int64_t launches = 0;
  int sm_count = 148;
  // FSB_STAGE_KERNELS=v1 selects the first-generation one-cell-per-thread stage kernels and the
  // unfused stage sequence (kept for A/B measurements and as a second implementation in the tests)
  bool stage_v1 = false;

  std::string err_msg;
};

// error plumbing ------------------------------------------------------------
int fsb_fail(fsb_ctx* ctx, int code, const char* fmt, ...);
#define FSB_CUDA(ctx, expr)                                                          \
  do                                                                                 \
  {                                                                                  \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return fsb_fail((ctx), FSB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)
#define FSB_TRY(expr)                 \
  do                                  \
  {                                   \
    int _rc = (expr);                 \
    if (_rc != FSB_OK) return _rc;    \
  } while (0)
// after a kernel launch
#define FSB_LAUNCHED(ctx)                    \
  do                                         \
  {                                          \
    (ctx)->launches++;                       \
    FSB_CUDA((ctx), cudaGetLastError());     \
  } while (0)

static inline int fsb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline float* fsb_uf(fsb_ctx* c) { return c->u[c->front]; }
static inline float* fsb_vf(fsb_ctx* c) { return c->v[c->front]; }
static inline float* fsb_ub(fsb_ctx* c) { return c->u[c->front ^ 1]; }
static inline float* fsb_vb(fsb_ctx* c) { return c->v[c->front ^ 1]; }

void fsb_prof_begin(fsb_ctx* ctx, int stage);
void fsb_prof_end(fsb_ctx* ctx, int stage);

// stage launchers (each returns FSB_OK or an error code) ----------------------
// grid stages: fsb_grid.cu
int fsb_k_classify(fsb_ctx* c);
int fsb_k_classify_reset(fsb_ctx* c);
int fsb_k_clear_labels(fsb_ctx* c);
int fsb_k_save_previous(fsb_ctx* c);
int fsb_k_update_diff(fsb_ctx* c);
int fsb_k_add_acceleration(fsb_ctx* c, float ax, float ay, float dt);
int fsb_k_enforce_dirichlet(fsb_ctx* c);
int fsb_k_prev_gravity_dirichlet(fsb_ctx* c, float ax, float ay, float dt, int save_prev);
int fsb_k_extend_velocity(fsb_ctx* c, int n_iter);
int fsb_k_extend_velocity_avg(fsb_ctx* c, int n_iter);
int fsb_k_advect_velocity_sl(fsb_ctx* c, float dt);
// particle stages: fsb_particles.cu
int fsb_k_sort_particles(fsb_ctx* c, bool mark_labels = false);
int fsb_k_p2g(fsb_ctx* c);
int fsb_k_g2p(fsb_ctx* c, int mode, float pic_ratio);
int fsb_k_advect_particles(fsb_ctx* c, float dt, int ensure_outside);
int fsb_k_g2p_advect(fsb_ctx* c, int mode, float pic_ratio, float dt, int ensure_outside);
int fsb_k_advect_particles_grid(fsb_ctx* c, float dt);
int fsb_k_unpermute(fsb_ctx* c, float4* dst_dense);
int fsb_k_p2g_gather(fsb_ctx* c);
int fsb_k_slab_mark_ghosts(fsb_ctx* c);
int fsb_k_slab_sort_out(fsb_ctx* c, int64_t* counts);
int fsb_k_slab_row_select(fsb_ctx* c, int row, int64_t* n_out);
int fsb_k_emit_source_dev(fsb_ctx* c, int64_t first, const float* xs_dev, const float* ys_dev,
                          int64_t count_x, int64_t count_y, float vel_x, float vel_y);
// semi-Lagrangian velocity advection as a deterministic gather: fsb_sl.cu
int fsb_k_advect_velocity_sl_gather(fsb_ctx* c, float dt, int* done);
void fsb_sl_free(fsb_ctx* c);
// pressure: fsb_cg.cu
int fsb_k_pressure_solve(fsb_ctx* c, float density, float dt, bool fuse_dirichlet = false);
void fsb_cg_reconfigure(fsb_ctx* c); // drop the CG launch configuration and graph (sharding changed)
// multigrid-preconditioned CG (fsb_mg.cu)
int fsb_k_mg_solve(fsb_ctx* c, int* converged);
void fsb_mg_free(fsb_ctx* c);
// one-sweep Jacobi-PCG (fsb_cg_one.cu): the default solve
int fsb_k_cg_one_solve(fsb_ctx* c, const CgCoef& coef);
int fsb_cg_one_partials(const fsb_ctx* c);

