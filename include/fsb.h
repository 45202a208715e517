/* fsb.h -- C ABI of the B200-native PIC/FLIP hot path (libfsb.so).
 *
 * Drop-in boundary for the per-step path of kbladin/Fluid_Simulation.  The
 * reference has no FFI: its boundary is the public C++ class surface of
 * fluidsim_lib (include/FluidSolver.h:45-54, include/FluidDomain.h:40-68,
 * include/MacGrid.h:19-159, include/MarkerParticleSet.h:53-74).  The C++ host
 * classes in include/fsb/ keep those class and method names and forward every
 * call to the functions below; each entry point cites the reference routine it
 * replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - plain pointers and sizes only; no CUDA or torch types in any signature
 *    (a CUDA stream is passed as void*).
 *  - every function returns FSB_OK or an FSB_ERR_* code; fsb_last_error()
 *    gives the text.  The C++ wrappers rethrow std::runtime_error, which is
 *    what the reference's step* functions throw (src/FluidSolver.cpp:101-107).
 *  - one context = one simulation domain on one CUDA device, one stream.  Calls
 *    are asynchronous on that stream; fsb_get_* and fsb_synchronize wait.
 *    A context is not thread-safe (neither is the reference).
 *  - all state lives in HBM.  Host buffers are caller-owned, dense row-major
 *    `i + j*size_x` exactly like Grid<T> (include/Grid.h:27-31); particles are
 *    AoS {pos_x,pos_y,vel_x,vel_y} like MarkerParticle
 *    (include/MarkerParticleSet.h:43-50) and always in the caller's original
 *    order, whatever ordering the device uses internally.
 *  - cell labels: 0 LIQUID, 1 AIR, 2 SOLID (include/MacGrid.h:14-17), one
 *    byte per cell.
 *  - there is no CPU fallback: every entry point fails with FSB_ERR_CUDA when
 *    no sm_100 device is usable.
 */
#ifndef FSB_H
#define FSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsb_ctx fsb_ctx;

enum {
  FSB_OK = 0,
  FSB_ERR_INVALID = 1, /* bad argument, or the reference's validate() mismatch */
  FSB_ERR_CUDA = 2,    /* CUDA runtime failure (incl. no device) */
  FSB_ERR_NOMEM = 3,
  FSB_ERR_COMM = 4     /* multi-GPU exchange failure / timeout */
};

/* MacGrid buffers (include/MacGrid.h:164-174) */
enum {
  FSB_U_FRONT = 0, FSB_V_FRONT = 1, FSB_U_BACK = 2, FSB_V_BACK = 3,
  FSB_U_PREV = 4, FSB_V_PREV = 5, FSB_U_DIFF = 6, FSB_V_DIFF = 7
};
enum { FSB_LIQUID = 0, FSB_AIR = 1, FSB_SOLID = 2 };
enum { FSB_G2P_PIC = 0, FSB_G2P_FLIP = 1, FSB_G2P_PICFLIP = 2 };
enum { FSB_STEP_SEMILAGRANGIAN = 0, FSB_STEP_PIC = 1, FSB_STEP_FLIP = 2, FSB_STEP_PICFLIP = 3 };
/* include/OdeSolver.h: RK3 :102-113 (the reference's hard-wired choice,
 * include/FluidSolver.h:144), EulerExplicit :78-86 */
enum { FSB_INTEGRATOR_RK3 = 0, FSB_INTEGRATOR_EULER = 1 };
/* preconditioner of the pressure CG: Eigen's DiagonalPreconditioner as the reference uses it
 * (include/FluidSolver.h:114), or -- opt-in, not the reference's algorithm -- one geometric
 * multigrid V-cycle (fluid_simulation_b200/csrc/fsb_mg.cu) */
enum { FSB_PRECOND_JACOBI = 0, FSB_PRECOND_MULTIGRID = 1 };

/* ---- life cycle ------------------------------------------------------- */

/* FluidDomain(size_x,size_y,length_x,length_y,density,pic_ratio)
 * (src/FluidDomain.cpp:54-68) + FluidSolverMemoryPool(domain)
 * (src/FluidSolver.cpp:28-54) + FluidSolver(pool) (:78-82) in one object.
 * `device` is the CUDA ordinal.  Labels start SOLID-border / AIR-interior
 * (src/MacGrid.cpp:24), velocities zero, no particles, CG cap 100 and
 * tolerance FLT_EPSILON (src/FluidSolver.cpp:81 and Eigen's default). */
int fsb_create(fsb_ctx** out, int size_x, int size_y, float length_x, float length_y,
               float density, float pic_ratio, int device);
void fsb_destroy(fsb_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL: last create error). */
const char* fsb_last_error(const fsb_ctx* ctx);
const char* fsb_version(void);
/* Run all later work of this context on `cuda_stream` (a cudaStream_t). */
int fsb_set_stream(fsb_ctx* ctx, void* cuda_stream);
int fsb_synchronize(fsb_ctx* ctx);

/* GridInterface getters (include/Grid.h:50-55) */
int fsb_size_x(const fsb_ctx* ctx);
int fsb_size_y(const fsb_ctx* ctx);
float fsb_delta_x(const fsb_ctx* ctx);
float fsb_delta_y(const fsb_ctx* ctx);

/* ---- parameters ------------------------------------------------------- */

/* _cg_solver.setMaxIterations / setTolerance (src/FluidSolver.cpp:81).
 * max_iters < 0 means Eigen's default cap (2 * unknowns). */
int fsb_set_cg(fsb_ctx* ctx, int max_iters, float tol);
/* Same system, stopping rule and result; the iteration count drops from O(grid size) to a few
 * tens.  Single-GPU solves only; if the multigrid iteration breaks down
 * (scattered single-cell obstacles defeat a geometric hierarchy) the solve is repeated with the
 * Jacobi preconditioner, so the flag never changes what is computed, only how fast. */
int fsb_set_preconditioner(fsb_ctx* ctx, int kind);
/* iterations() and error() of the last solve */
int fsb_get_cg_info(const fsb_ctx* ctx, int* iterations, float* error);
/* FluidDomain::setPicRatio (src/FluidDomain.cpp:99-102): clamped to [0,1] */
int fsb_set_pic_ratio(fsb_ctx* ctx, float pic_ratio);
int fsb_set_density(fsb_ctx* ctx, float density);
int fsb_set_integrator(fsb_ctx* ctx, int integrator);
/* FluidSolverMemoryPool(size_x,size_y,delta_x,delta_y) as held by the FluidSolver
 * (src/FluidSolver.cpp:5-26,56-65): the step functions compare it with the domain
 * (validate(), :89-97) and the P2G accumulators carry its deltas.  The default is
 * the pool FluidSolver(FluidSolverMemoryPool(domain)) ends up with: the domain's
 * sizes and delta_x for BOTH deltas (the copy constructor passes deltaX twice). */
int fsb_set_pool(fsb_ctx* ctx, int size_x, int size_y, float delta_x, float delta_y);
/* gravity used by the fused steps; default (0, (float)-9.82)
 * (src/FluidSolver.cpp:115,154,192,232) */
int fsb_set_gravity(fsb_ctx* ctx, float ax, float ay);

/* ---- state transfer --------------------------------------------------- */

/* MarkerParticleSet::clear + addParticle (include/MarkerParticleSet.h:56-58) */
int fsb_set_particles(fsb_ctx* ctx, const float* aos4, int64_t n);
int fsb_append_particles(fsb_ctx* ctx, const float* aos4, int64_t n);
int64_t fsb_num_particles(const fsb_ctx* ctx);
/* iteration over the set (include/MarkerParticleSet.h:68-74), original order */
int fsb_get_particles(fsb_ctx* ctx, float* aos4);
/* FluidSource::update spawn branch (src/FluidDomain.cpp:34-50) executed on the
 * device: appends the delta/2.5 lattice over [x_min,x_max) x [y_min,y_max)
 * in the reference's order.  *n_added receives the count (may be NULL). */
int fsb_emit_source(fsb_ctx* ctx, float x_min, float x_max, float y_min, float y_max,
                    float delta_x, float delta_y, float vel_x, float vel_y,
                    int64_t* n_added);

int fsb_set_grid(fsb_ctx* ctx, int which, const float* src);
int fsb_get_grid(fsb_ctx* ctx, int which, float* dst);
/* MacGrid::setCellType / cellType (include/MacGrid.h:92-97,140-143) */
int fsb_set_cell_types(fsb_ctx* ctx, const uint8_t* src);
int fsb_get_cell_types(fsb_ctx* ctx, uint8_t* dst);
/* the CG solution of the last pressure solve on the full grid (0 where not
 * LIQUID); the reference discards it (src/FluidSolver.cpp:423-460) */
int fsb_get_pressure(fsb_ctx* ctx, float* dst);

/* ---- one entry per reference stage ----------------------------------- */

/* FluidDomain::classifyCells(MarkerParticleSet&) src/FluidDomain.cpp:150-180 */
int fsb_classify_cells(fsb_ctx* ctx);
/* FluidSolver::transferVelocityToGridSpread src/FluidSolver.cpp:873-919 */
int fsb_p2g_spread(fsb_ctx* ctx);
/* MacGrid::updatePreviousVelocityBuffer src/MacGrid.cpp:52-56 */
int fsb_save_previous(fsb_ctx* ctx);
/* MacGrid::clearCellTypeBuffer src/MacGrid.cpp:32-50: border SOLID, interior AIR */
int fsb_clear_cell_types(fsb_ctx* ctx);
/* MacGrid::swapVelocityBuffers src/MacGrid.cpp:89-93 */
int fsb_swap_velocity_buffers(fsb_ctx* ctx);
/* FluidSolver::addExternalAcceleration src/FluidSolver.cpp:276-295 */
int fsb_add_acceleration(fsb_ctx* ctx, float ax, float ay, float dt);
/* FluidSolver::enforceDirichlet src/FluidSolver.cpp:297-321 */
int fsb_enforce_dirichlet(fsb_ctx* ctx);
/* FluidSolver::extendVelocityIndividual src/FluidSolver.cpp:485-622 */
int fsb_extend_velocity(fsb_ctx* ctx, int n_iterations);
/* FluidSolver::pressureSolve src/FluidSolver.cpp:323-483 (Eigen CG replaced by
 * the matrix-free Jacobi-PCG kernels) */
int fsb_pressure_solve(fsb_ctx* ctx, float density, float dt);
/* MacGrid::updateVelocityDiffBuffer src/MacGrid.cpp:58-70 */
int fsb_update_diff(fsb_ctx* ctx);
/* FluidSolver::transferVelocityToParticles{PIC,FLIP,PICFLIP} src/FluidSolver.cpp:921-963 */
int fsb_g2p(fsb_ctx* ctx, int mode, float pic_ratio);
/* MarkerParticleSet::advect / advectAndEnsureOutsideObstacles src/MarkerParticleSet.cpp:40-62 */
int fsb_advect_particles(fsb_ctx* ctx, float dt, int ensure_outside_obstacles);
/* FluidSolver::advectVelocitySemiLagrangian src/FluidSolver.cpp:709-772 */
int fsb_advect_velocity_sl(fsb_ctx* ctx, float dt);
/* FluidSolver::advectParticlesWithGrid src/FluidSolver.cpp:774-791 */
int fsb_advect_particles_grid(fsb_ctx* ctx, float dt);

/* FluidSolver routines that no step* calls (API completeness):
 * addExternalForce src/FluidSolver.cpp:253-274 -- F / density * dt on the left and bottom faces
 * of LIQUID cells, with the context's density */
int fsb_add_external_force(fsb_ctx* ctx, float fx, float fy, float dt);
/* transferVelocityToGridGather src/FluidSolver.cpp:816-871 -- every face takes the mean velocity
 * of the particles whose hat weight is >= 1 (particles numerically ON the face position), in
 * particle-set order; other faces keep the back buffer's value; swap.  The reference scans all
 * particles per face; the device scans the face's 3x3 cells of the cell-sorted set. */
int fsb_p2g_gather(fsb_ctx* ctx);
/* extendVelocityAvarageing src/FluidSolver.cpp:625-707 -- per-CELL validity (the x masks of the pool);
 * an unmarked non-SOLID cell takes the mean of the cell-centred back-buffer velocities of its marked
 * neighbours (order (i-1,j), (i,j-1), (i,j+1), (i+1,j)) and writes it to both of its faces per
 * component, so a sweep depends on the row-major scan order: the device runs it as skewed wavefronts
 * (t = i + 2 j) that reproduce the sequential scan bit for bit.  Masks swapped per sweep, velocities
 * at the end.  FSB_ERR_INVALID unless the border cells are SOLID (the reference asserts there). */
int fsb_extend_velocity_averaging(fsb_ctx* ctx, int n_iterations);

/* ---- fused steps: FluidSolver::step* src/FluidSolver.cpp:99-251 ------- */

/* kind = FSB_STEP_*.  Uses the context's density and pic ratio the way the
 * reference reads them from the FluidDomain.  FSB_ERR_INVALID when the
 * reference's validate() (src/FluidSolver.cpp:89-97) would throw. */
int fsb_step(fsb_ctx* ctx, int kind, float dt);

/* ---- state dump / reload -------------------------- */

/* Everything a step reads, as one little-endian file: a 96-byte header (magic "FSBSTATE",
 * version, sizes, deltas, density, pic ratio, gravity, integrator, CG cap and tolerance, pool,
 * particle count), the labels (size_x*size_y bytes), the eight MacGrid buffers in FSB_U_FRONT ..
 * FSB_V_DIFF order (dense fp32), the particles in the DEVICE's order (AoS fp32 x 4; cell-sorted
 * by the last step) and the int32 map from that order to the caller's indices (particle k of the
 * file is the caller's particle map[k]).  A run continued from a reloaded file is bit-identical
 * to the uninterrupted run (the cell sort orders every cell by original index, so the stored
 * order does not matter).  The reference has no state format (its only output is the PPM frame,
 * src/Renderer.cpp). */
int fsb_save_state(fsb_ctx* ctx, const char* path);
/* The context must have the file's grid size. */
int fsb_load_state(fsb_ctx* ctx, const char* path);

/* ---- frames -------------------------------------- */

/* The frame examples/simple.cpp:73-82 draws with the reference's software renderer --
 * Renderer::clearCanvas + renderGridCellsToCanvas + renderParticlesToCanvas (src/Renderer.cpp:
 * 14-56,141-162, src/Canvas.cpp:62-92) and the byte conversion of writeCanvasToPpm (:217-248) --
 * rasterised on the device from the state in HBM: width*height*3 bytes, row j of the canvas at
 * offset j*width*3, byte-identical to the reference's PPM payload.  (x_min..y_max) is the
 * world-space area of Renderer's constructor. */
int fsb_render_rgb(fsb_ctx* ctx, int width, int height, float x_min, float x_max, float y_min,
                   float y_max, uint8_t* rgb);
/* Renderer::writeCanvasToPpm of that frame: "P6\n<w> <h>\n255\n" + the bytes */
int fsb_write_ppm(fsb_ctx* ctx, const char* path, int width, int height, float x_min, float x_max,
                  float y_min, float y_max);

/* ---- multi-GPU: row-slab sharding of the pressure solve ---------------- */

/* One process per GPU, each holding the same domain (every stage except the CG
 * runs replicated and is deterministic); the CG iterates on `world` row slabs.
 * Slab halos travel as direct peer-memory stores over NVLink and the dot
 * products through peer-memory mailboxes (CUDA IPC), so an iteration issues no
 * collective call.  The host exchanges the handle blobs once (any transport:
 * torch.distributed, MPI, files):
 *   fsb_shard_export  on every rank -> blob of FSB_SHARD_BLOB_BYTES
 *   all-gather the blobs in rank order
 *   fsb_shard_connect(rank, world, all_blobs)
 * After that fsb_pressure_solve / fsb_step on every rank cooperate; every rank
 * must make the same sequence of calls.  FSB_ERR_COMM: a peer did not answer. */
#define FSB_SHARD_BLOB_BYTES 512
int fsb_shard_export(fsb_ctx* ctx, void* blob);
int fsb_shard_connect(fsb_ctx* ctx, int rank, int world, const void* all_blobs);
int fsb_shard_disconnect(fsb_ctx* ctx);
/* rows [*row_lo, *row_hi) of the grid this rank iterates on */
int fsb_shard_rows(const fsb_ctx* ctx, int* row_lo, int* row_hi);

/* ---- multi-GPU: particle slabs --------------------------- */

/* One process per GPU; every rank keeps the full grids, but only the particles whose row (the
 * P2G sort-key row, src/FluidSolver.cpp:873-919) lies in its slab, plus one ghost row from each
 * neighbour while a step is in flight.  The canonical in-cell particle order goes by global id, so
 * a slab-partitioned run produces the SAME BITS as the single-GPU run.  A PIC / FLIP / PIC-FLIP step
 * is three device phases with two exchanges in between; the transport is the caller's
 * (fluid_simulation_b200/sharding.py: torch.distributed / NCCL; tests: in-process copies):
 *
 *   fsb_slab_boundary(side) + _take  -> send to the neighbour -> fsb_slab_add      (ghost rows)
 *   fsb_slab_step_a(kind)            labels from the local particles, P2G of the own rows
 *   fsb_get_rows / fsb_set_rows      all-gather the label, u and v rows of every slab
 *   fsb_slab_step_b(kind, dt)        gravity, walls, extension, pressure solve (optionally sharded)
 *   fsb_slab_step_c(kind, dt)        ghosts retired; G2P + blend + advection of the own particles
 *   fsb_slab_sort_out -> fsb_slab_take(dest) -> send -> fsb_slab_keep_own -> fsb_slab_add (migration)
 *
 * A semi-Lagrangian step (src/FluidSolver.cpp:99-134) uses the same three phases without ghost rows and with
 * the label rows only in the exchange: a marker particle labels its own cell, and the velocity lives on the
 * grid, advected by every rank with identical bits (the gather form of fsb_advect_velocity_sl):
 *
 *   fsb_slab_step_a(SL)      labels of the own rows        -> all-gather the label rows
 *   fsb_slab_step_b(SL, dt)  velocity advection, gravity, walls, pressure solve (optionally sharded)
 *   fsb_slab_step_c(SL, dt)  RK3 trace of the own particles -> migration
 *
 * Buffers may be host or device pointers (cudaMemcpyDefault).  Particle ids are the caller's
 * global indices (fsb_set_particles / fsb_emit_source number them 0 .. n-1 identically on all ranks). */
#define FSB_ROWS_LABELS 8
int fsb_slab_configure(fsb_ctx* ctx, int rank, int world);
int fsb_slab_rows(const fsb_ctx* ctx, int* row_lo, int* row_hi);
int fsb_slab_add(fsb_ctx* ctx, const float* aos4, const int32_t* ids, int64_t n);
/* groups the live particles by owner rank; counts[world] */
int fsb_slab_sort_out(fsb_ctx* ctx, int64_t* counts);
int fsb_slab_take(fsb_ctx* ctx, int dest, float* aos4, int32_t* ids);
int fsb_slab_keep_own(fsb_ctx* ctx);
/* own particles of the slab's first (side 0) / last (side 1) row: count, then copy */
int fsb_slab_boundary(fsb_ctx* ctx, int side, int64_t* n);
int fsb_slab_boundary_take(fsb_ctx* ctx, float* aos4, int32_t* ids);
/* this rank's particles and their global ids (fsb_num_particles of them), device order */
int fsb_slab_get(fsb_ctx* ctx, float* aos4, int32_t* ids);
/* rows [row_lo, row_hi) of a MacGrid buffer (FSB_U_FRONT ..) or of the labels (FSB_ROWS_LABELS), dense */
int fsb_get_rows(fsb_ctx* ctx, int which, int row_lo, int row_hi, void* dst);
int fsb_set_rows(fsb_ctx* ctx, int which, int row_lo, int row_hi, const void* src);
int fsb_slab_step_a(fsb_ctx* ctx, int kind);
int fsb_slab_step_b(fsb_ctx* ctx, int kind, float dt);
int fsb_slab_step_c(fsb_ctx* ctx, int kind, float dt);

/* ---- measurement ------------------------------------------------------ */

/* Per-stage CUDA-event timing on the context's stream.  Enabling it makes
 * every stage record an event pair; fsb_profile_read drains them (this
 * synchronises).  Stage ids: FSB_PROF_*. */
enum {
  FSB_PROF_CLASSIFY = 0, FSB_PROF_SORT, FSB_PROF_P2G, FSB_PROF_GRID_PRE,
  FSB_PROF_EXTEND, FSB_PROF_RHS, FSB_PROF_CG, FSB_PROF_PATCH, FSB_PROF_G2P,
  FSB_PROF_ADVECT_SL, FSB_PROF_ADVECT_PART, FSB_PROF_COUNT
};
int fsb_profile_enable(fsb_ctx* ctx, int on);
/* ms[FSB_PROF_COUNT], calls[FSB_PROF_COUNT]: totals since the last read */
int fsb_profile_read(fsb_ctx* ctx, float* ms, int* calls);
/* launch mode of the pressure solve as configured by the last solve: 0 not configured yet,
 * 1 two kernels per iteration in a CUDA graph (FSB_CG_MODE=graph), 2 one persistent cooperative
 * kernel with two sweeps per iteration (FSB_CG_MODE=fused),
 * 3 the last solve was a multigrid-preconditioned CG (FSB_PRECOND_MULTIGRID),
 * 4 one persistent cooperative kernel with ONE sweep and one reduction per iteration (default) */
int fsb_cg_launch_mode(const fsb_ctx* ctx);
/* cells the sweeps of the last pressure solve visited per iteration on THIS rank: the cells of the
 * tiles that hold at least one LIQUID cell (all tiles of the rank's rows when the active-tile list is
 * off); 0 before the first solve.  Measurement only: the byte count of the roofline figure. */
int64_t fsb_cg_swept_cells(const fsb_ctx* ctx);
/* number of kernels this library launched on the context since creation */
int64_t fsb_launch_count(const fsb_ctx* ctx);
/* a pair of events around an arbitrary region of this context's stream */
int fsb_timer_start(fsb_ctx* ctx);
int fsb_timer_stop(fsb_ctx* ctx, float* elapsed_ms); /* synchronises */

#ifdef __cplusplus
}
#endif
#endif /* FSB_H */
