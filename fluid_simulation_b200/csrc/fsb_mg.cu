// Opt-in multigrid-preconditioned CG for the pressure Poisson system.
//
// The reference hands the system to Eigen's Jacobi-preconditioned CG (src/FluidSolver.cpp:424-425);
// that is what fsb_cg.cu reproduces and what every parity test and the headline benchmark use.
// Jacobi-PCG needs O(N) iterations on an N x N grid (14 000 at 4096^2).  This file replaces the
// preconditioner -- not the system, the stopping rule or the result -- by one geometric multigrid
// V-cycle (after McAdams, Sifakis, Teran: "A parallel multigrid Poisson solver for fluids
// simulation on large grids", 2010):
//   * levels: 2x2 cell coarsening down to <= 4 x 4; a coarse cell is AIR (Dirichlet) if any child
//     is AIR, else LIQUID if any child is LIQUID, else SOLID; the 5-point operator is re-discretised
//     on the coarse labels with h -> 2h (same matrix-free form as the fine operator: no stored
//     matrix on any level);
//   * smoother: damped Jacobi (omega = 2/3), 3 pre- and 3 post-sweeps (FSB_MG_SWEEPS); restriction = tensor
//     product of (1 3 3 1)/8, prolongation = 4 x its transpose (cell-centred bilinear), both with the weights
//     renormalised over the non-SOLID coarse parents next to walls (fsb_mg_kernels.cuh: without that the
//     restriction is not conservative at a wall and the iteration count grows with the grid -- 17 / 24 / 50 at
//     1024^2 / 2048^2 / 4096^2 in round 1, 7 - 8 at every size with it; tools/studies/mg_transfer_study.py);
//     the coarsest level is solved by 40 Jacobi sweeps inside one CTA's shared memory.  Equal pre/post sweeps
//     and P = 4 R^T keep the V-cycle symmetric, as CG requires.
// The outer iteration is Eigen's statement order with z = V(r) in place of z = D^-1 r and the same
// stopping rule |r|^2 < tol^2 |b|^2.  Every kernel is a memory-bound sweep with four cells per
// thread; reductions are per-CTA fp64 partials folded in a fixed order by the last CTA.
// A geometric hierarchy cannot represent scattered single-cell obstacles: if the iteration breaks
// down (non-finite scalars, r.z or p.Ap of the wrong sign, residual growth, iteration cap) the
// caller falls back to the Jacobi solve, so the flag can never produce a wrong answer.
#include <cfloat>
#include <cmath>
#include <vector>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"
#include "fsb_mg_kernels.cuh"

namespace {

constexpr int kMgMaxLevels = 16;
// coarsening continues to <= 4 x 4 (kMgStopDefault; FSB_MG_STOP): with the wall-conservative transfers of
// fsb_mg_kernels.cuh the coarse levels represent the smooth modes well, and 40 sweeps solve a 4 x 4 level
// exactly -- 40 sweeps on a 32 x 32 coarsest level leave its lowest modes almost untouched (per-sweep factor
// 0.9995 for the fundamental), which showed as a plateau of ~10 PCG iterations.
constexpr int kMgCoarseSweeps = 40;

struct MgLevel
{
  int nx = 0, ny = 0, ld = 0;
  float inv_h2 = 0;
  uint8_t* lab = nullptr;  // level 0: the context's labels (not owned)
  uint8_t* code = nullptr; // level 0: the CG's stencil codes (not owned)
  float* x[2] = {nullptr, nullptr};
  float* b = nullptr; // level 0: the CG residual (not owned)
  float* r = nullptr;
};

struct MgScalars
{
  double pq, r2, rz, rz_old, rhs2;
  float alpha, beta, thr;
  int iter, done, fail;
  unsigned int ticket;
};

} // namespace

struct fsb_mg_state
{
  int n_levels = 0;
  MgLevel lv[kMgMaxLevels];
  MgScalars* scal = nullptr;   // device
  MgScalars* scal_h = nullptr; // pinned
  double* partials = nullptr;
  int partials_cap = 0;
  int nx = 0, ny = 0;
  // the V-cycle as one CUDA graph (about 90 launches, most of them on levels of a few thousand cells whose
  // cost is the launch itself): captured on first use -- buffers, sizes and the sweep count are fixed for
  // the life of the hierarchy -- and replayed for every application of the preconditioner
  cudaGraphExec_t vexec = nullptr;
  int vexec_cur = 0;      // which of lv[0].x[] holds the result
  int vexec_launches = 0; // kernels inside the graph (for the launch counter)
  bool vexec_failed = false;
};

namespace {

MgCoef make_mg_coef(float inv_h2)
{
  MgCoef k;
  k.inv_h2 = inv_h2;
  k.wdinv[0] = 0.0f;
  for (int n = 1; n < 5; ++n) k.wdinv[n] = kMgOmega * (-1.0f / ((float)n * inv_h2));
  return k;
}

__device__ __forceinline__ bool mg_last_block(unsigned int* ticket)
{
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  return s_last;
}
__device__ __forceinline__ double mg_fold(const volatile double* part, int n)
{
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) s += part[k];
  return block_sum(s);
}

// ---- outer PCG kernels (grid-stride over (row, 1024-column segment) work items) ------------
#define MG_FOR_GROUPS(ld, ny)                                                   \
  const int segs = ((ld) + 1023) / 1024;                                        \
  for (int w = blockIdx.x; w < segs * (ny); w += gridDim.x)                     \
    for (int j = w / segs, i0 = ((w - j * segs) * 256 + threadIdx.x) * 4, once = 1; once && i0 < (ld); once = 0)

// q = A p, partial p.q
__global__ void __launch_bounds__(256)
k_mg_apply_dot(const float* __restrict__ p, const uint8_t* __restrict__ code, float* __restrict__ q,
               int nx, int ny, int ld, float inv_h2, MgScalars* __restrict__ s,
               double* __restrict__ partials)
{
  double acc = 0.0;
  MG_FOR_GROUPS(ld, ny)
  {
    const size_t k = i0 + (size_t)j * ld;
    const uint32_t c4 = *reinterpret_cast<const uint32_t*>(code + k);
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 != 0u)
    {
      const float4 pc = *reinterpret_cast<const float4*>(p + k);
      out = mg_apply4(p, ld, ny, i0, j, pc, c4, inv_h2);
      acc += (double)pc.x * out.x + (double)pc.y * out.y + (double)pc.z * out.z + (double)pc.w * out.w;
    }
    *reinterpret_cast<float4*>(q + k) = out;
  }
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  if (mg_last_block(&s->ticket))
  {
    const double pq = mg_fold(partials, gridDim.x);
    if (threadIdx.x == 0)
    {
      s->pq = pq;
      // A is negative definite on the liquid cells and so must be the preconditioner: rz < 0, pq < 0
      if (!(pq < 0.0) || !(s->rz < 0.0)) s->fail = 1;
      s->alpha = (float)(s->rz / pq);
      s->ticket = 0;
    }
  }
}

// x += alpha p, r -= alpha q, partial |r|^2; then the stopping rule
__global__ void __launch_bounds__(256)
k_mg_update(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p,
            const float* __restrict__ q, int ny, int ld, MgScalars* __restrict__ s,
            double* __restrict__ partials, int max_iters)
{
  const float alpha = s->alpha;
  double acc = 0.0;
  MG_FOR_GROUPS(ld, ny)
  {
    const size_t k = i0 + (size_t)j * ld;
    const float4 p4 = *reinterpret_cast<const float4*>(p + k);
    const float4 q4 = *reinterpret_cast<const float4*>(q + k);
    float4 x4 = *reinterpret_cast<const float4*>(x + k);
    float4 r4 = *reinterpret_cast<const float4*>(r + k);
    x4.x = fmaf(alpha, p4.x, x4.x); x4.y = fmaf(alpha, p4.y, x4.y);
    x4.z = fmaf(alpha, p4.z, x4.z); x4.w = fmaf(alpha, p4.w, x4.w);
    r4.x = fmaf(-alpha, q4.x, r4.x); r4.y = fmaf(-alpha, q4.y, r4.y);
    r4.z = fmaf(-alpha, q4.z, r4.z); r4.w = fmaf(-alpha, q4.w, r4.w);
    *reinterpret_cast<float4*>(x + k) = x4;
    *reinterpret_cast<float4*>(r + k) = r4;
    acc += (double)r4.x * r4.x + (double)r4.y * r4.y + (double)r4.z * r4.z + (double)r4.w * r4.w;
  }
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  if (mg_last_block(&s->ticket))
  {
    const double r2 = mg_fold(partials, gridDim.x);
    if (threadIdx.x == 0)
    {
      s->r2 = r2;
      if (!(r2 == r2) || r2 > 1.0e8 * s->rhs2) s->fail = 1; // NaN or runaway
      if ((float)r2 < s->thr) s->done = 1;                  // Eigen: converged, break before i++
      else
      {
        s->iter += 1;
        if (s->iter >= max_iters) s->done = 1;
      }
      s->ticket = 0;
    }
  }
}

// partial r.z; the last CTA forms beta = rz_new / rz_old
__global__ void __launch_bounds__(256)
k_mg_dot_rz(const float* __restrict__ r, const float* __restrict__ z, int ny, int ld,
            MgScalars* __restrict__ s, double* __restrict__ partials, int first)
{
  double acc = 0.0;
  MG_FOR_GROUPS(ld, ny)
  {
    const size_t k = i0 + (size_t)j * ld;
    const float4 r4 = *reinterpret_cast<const float4*>(r + k);
    const float4 z4 = *reinterpret_cast<const float4*>(z + k);
    acc += (double)r4.x * z4.x + (double)r4.y * z4.y + (double)r4.z * z4.z + (double)r4.w * z4.w;
  }
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  if (mg_last_block(&s->ticket))
  {
    const double rz = mg_fold(partials, gridDim.x);
    if (threadIdx.x == 0)
    {
      s->rz_old = s->rz;
      s->rz = rz;
      s->beta = first ? 0.0f : (float)(rz / s->rz_old);
      if (!(rz < 0.0)) s->fail = 1;
      s->ticket = 0;
    }
  }
}

// p = z + beta p
__global__ void __launch_bounds__(256)
k_mg_direction(float* __restrict__ p, const float* __restrict__ z, int ny, int ld,
               const MgScalars* __restrict__ s)
{
  const float beta = s->beta;
  MG_FOR_GROUPS(ld, ny)
  {
    const size_t k = i0 + (size_t)j * ld;
    const float4 z4 = *reinterpret_cast<const float4*>(z + k);
    float4 p4 = *reinterpret_cast<const float4*>(p + k);
    p4.x = fmaf(beta, p4.x, z4.x); p4.y = fmaf(beta, p4.y, z4.y);
    p4.z = fmaf(beta, p4.z, z4.z); p4.w = fmaf(beta, p4.w, z4.w);
    *reinterpret_cast<float4*>(p + k) = p4;
  }
}

__global__ void k_mg_init_scalars(MgScalars* s, const CgScalars* cg, int max_iters)
{
  s->pq = 0.0; s->r2 = cg->r2; s->rz = 0.0; s->rz_old = 0.0; s->rhs2 = cg->rhs2;
  s->alpha = 0.0f; s->beta = 0.0f; s->thr = cg->thr;
  s->iter = 0; s->done = 0; s->fail = 0; s->ticket = 0;
  (void)max_iters;
}

template <class T>
int mg_alloc(fsb_ctx* c, T** p, size_t count)
{
  *p = nullptr;
  if (cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)) != cudaSuccess)
    return fsb_fail(c, FSB_ERR_NOMEM, "cudaMalloc of %zu bytes failed (multigrid levels)", sizeof(T) * count);
  FSB_CUDA(c, cudaMemsetAsync(*p, 0, sizeof(T) * (count ? count : 1), c->stream));
  return FSB_OK;
}

int mg_build_hierarchy(fsb_ctx* c)
{
  if (c->mg && c->mg->nx == c->nx && c->mg->ny == c->ny) return FSB_OK;
  fsb_mg_free(c);
  fsb_mg_state* m = new fsb_mg_state();
  c->mg = m;
  m->nx = c->nx; m->ny = c->ny;
  const double dx2 = std::pow((double)c->dx, 2);
  int nx = c->nx, ny = c->ny;
  float inv_h2 = (float)(1 / dx2);
  for (int l = 0; l < kMgMaxLevels; ++l)
  {
    MgLevel& L = m->lv[l];
    L.nx = nx; L.ny = ny; L.ld = (nx + 31) / 32 * 32; L.inv_h2 = inv_h2;
    const size_t cells = (size_t)L.ld * ny;
    if (l == 0)
    {
      L.lab = c->cell; L.code = c->cg_code; L.b = c->cg_r;
    }
    else
    {
      FSB_TRY(mg_alloc(c, &L.lab, cells));
      FSB_TRY(mg_alloc(c, &L.code, cells));
      FSB_TRY(mg_alloc(c, &L.b, cells));
    }
    FSB_TRY(mg_alloc(c, &L.x[0], cells));
    FSB_TRY(mg_alloc(c, &L.x[1], cells));
    FSB_TRY(mg_alloc(c, &L.r, cells));
    m->n_levels = l + 1;
    if (nx <= c->mg_stop && ny <= c->mg_stop) break;
    nx = (nx + 1) / 2; ny = (ny + 1) / 2; inv_h2 *= 0.25f;
  }
  const MgLevel& last = m->lv[m->n_levels - 1];
  if (last.nx > kMgCoarsest || last.ny > kMgCoarsest)
    return fsb_fail(c, FSB_ERR_INVALID, "grid too large for %d multigrid levels", kMgMaxLevels);
  FSB_TRY(mg_alloc(c, &m->scal, 1));
  if (cudaMallocHost((void**)&m->scal_h, sizeof(MgScalars)) != cudaSuccess)
    return fsb_fail(c, FSB_ERR_NOMEM, "cudaMallocHost failed");
  m->partials_cap = c->sm_count * 8;
  FSB_TRY(mg_alloc(c, &m->partials, (size_t)m->partials_cap));
  return FSB_OK;
}

inline dim3 grid4(const MgLevel& L) { return dim3(fsb_div_up(L.ld, 1024), L.ny); }
inline dim3 grid1(const MgLevel& L) { return dim3(fsb_div_up(L.ld, 256), L.ny); }

// labels and stencil codes of the coarse levels (level 0 shares the CG's)
int mg_setup_labels(fsb_ctx* c)
{
  fsb_mg_state* m = c->mg;
  for (int l = 1; l < m->n_levels; ++l)
  {
    const MgLevel& F = m->lv[l - 1];
    const MgLevel& C = m->lv[l];
    k_mg_coarsen_labels<<<grid1(C), 256, 0, c->stream>>>(F.lab, F.nx, F.ny, F.ld, C.lab, C.nx, C.ny, C.ld);
    FSB_LAUNCHED(c);
    k_mg_codes<<<grid1(C), 256, 0, c->stream>>>(C.lab, C.code, C.nx, C.ny, C.ld);
    FSB_LAUNCHED(c);
  }
  return FSB_OK;
}

// z = V(r): result in lv[0].x[*cur0]
int mg_vcycle(fsb_ctx* c, int* cur_out)
{
  fsb_mg_state* m = c->mg;
  int cur[kMgMaxLevels];
  const int last = m->n_levels - 1;
  for (int l = 0; l < last; ++l)
  {
    MgLevel& L = m->lv[l];
    const MgCoef kf = make_mg_coef(L.inv_h2);
    cur[l] = 0;
    k_mg_smooth<true><<<grid4(L), 256, 0, c->stream>>>(nullptr, L.b, L.code, L.x[0], L.nx, L.ny, L.ld, kf);
    FSB_LAUNCHED(c);
    for (int s = 1; s < c->mg_sweeps; ++s)
    {
      k_mg_smooth<false><<<grid4(L), 256, 0, c->stream>>>(L.x[cur[l]], L.b, L.code, L.x[cur[l] ^ 1], L.nx, L.ny, L.ld, kf);
      FSB_LAUNCHED(c);
      cur[l] ^= 1;
    }
    k_mg_residual<<<grid4(L), 256, 0, c->stream>>>(L.x[cur[l]], L.b, L.code, L.r, L.nx, L.ny, L.ld, L.inv_h2);
    FSB_LAUNCHED(c);
    MgLevel& C = m->lv[l + 1];
    k_mg_restrict<<<grid1(C), 256, 0, c->stream>>>(L.r, L.nx, L.ny, L.ld, C.code, C.b, C.nx, C.ny, C.ld,
                                                   c->mg_renorm ? C.lab : nullptr);
    FSB_LAUNCHED(c);
  }
  {
    MgLevel& L = m->lv[last];
    cur[last] = 0;
    k_mg_coarse_solve<<<1, 1024, 0, c->stream>>>(L.b, L.code, L.x[0], L.nx, L.ny, L.ld,
                                                  make_mg_coef(L.inv_h2), kMgCoarseSweeps);
    FSB_LAUNCHED(c);
  }
  for (int l = last - 1; l >= 0; --l)
  {
    MgLevel& L = m->lv[l];
    MgLevel& C = m->lv[l + 1];
    const MgCoef kf = make_mg_coef(L.inv_h2);
    k_mg_prolong_add<<<grid4(L), 256, 0, c->stream>>>(L.x[cur[l]], L.code, L.nx, L.ny, L.ld,
                                                      C.x[cur[l + 1]], C.nx, C.ny, C.ld,
                                                      c->mg_renorm ? C.lab : nullptr);
    FSB_LAUNCHED(c);
    for (int s = 0; s < c->mg_sweeps; ++s)
    {
      k_mg_smooth<false><<<grid4(L), 256, 0, c->stream>>>(L.x[cur[l]], L.b, L.code, L.x[cur[l] ^ 1], L.nx, L.ny, L.ld, kf);
      FSB_LAUNCHED(c);
      cur[l] ^= 1;
    }
  }
  *cur_out = cur[0];
  return FSB_OK;
}

// z = V(r) through the captured graph (FSB_MG_GRAPH=0: plain launches)
int mg_vcycle_run(fsb_ctx* c, int* cur_out)
{
  fsb_mg_state* m = c->mg;
  if (!c->mg_graph || m->vexec_failed) return mg_vcycle(c, cur_out);
  if (!m->vexec)
  {
    const long long before = c->launches;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    {
      cudaGetLastError();
      m->vexec_failed = true;
      return mg_vcycle(c, cur_out);
    }
    const int rc = mg_vcycle(c, &m->vexec_cur);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    m->vexec_launches = (int)(c->launches - before);
    c->launches = before;
    if (rc != FSB_OK || e != cudaSuccess || !g ||
        cudaGraphInstantiate(&m->vexec, g, 0) != cudaSuccess)
    {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      m->vexec = nullptr;
      m->vexec_failed = true;
      return mg_vcycle(c, cur_out);
    }
    cudaGraphDestroy(g);
  }
  FSB_CUDA(c, cudaGraphLaunch(m->vexec, c->stream));
  c->launches += m->vexec_launches;
  *cur_out = m->vexec_cur;
  return FSB_OK;
}

} // namespace

void fsb_mg_free(fsb_ctx* c)
{
  if (!c->mg) return;
  fsb_mg_state* m = c->mg;
  if (m->vexec) cudaGraphExecDestroy(m->vexec);
  for (int l = 0; l < m->n_levels; ++l)
  {
    MgLevel& L = m->lv[l];
    if (l > 0)
    {
      cudaFree(L.lab); cudaFree(L.code); cudaFree(L.b);
    }
    cudaFree(L.x[0]); cudaFree(L.x[1]); cudaFree(L.r);
  }
  cudaFree(m->scal); cudaFree(m->partials);
  if (m->scal_h) cudaFreeHost(m->scal_h);
  delete m;
  c->mg = nullptr;
}

// Called by fsb_k_pressure_solve after k_cg_build (x = 0, r = b, |b|^2 and the threshold in the CG
// scalars).  *converged = 1: x holds the solution, c->iters / c->err are set.  *converged = 0: the
// iteration broke down or hit its cap; the caller re-runs the set-up and solves with Jacobi-PCG.
int fsb_k_mg_solve(fsb_ctx* c, int* converged)
{
  *converged = 0;
  FSB_TRY(mg_build_hierarchy(c));
  fsb_mg_state* m = c->mg;
  FSB_TRY(mg_setup_labels(c));
  MgLevel& L0 = m->lv[0];
  const int blocks = (int)std::min<int64_t>(m->partials_cap,
                                            std::max<int64_t>(1, (int64_t)fsb_div_up(L0.ld, 1024) * L0.ny));
  // the cap: a healthy multigrid iteration needs tens of iterations; beyond this Jacobi is used
  const int cap = std::min(c->mg_max_iters, c->max_iters < 0 ? c->mg_max_iters : c->max_iters);
  k_mg_init_scalars<<<1, 1, 0, c->stream>>>(m->scal, c->scal, cap);
  FSB_LAUNCHED(c);
  float* p = c->cg_p[0];
  float* q = c->cg_p[1];
  int zc = 0;
  FSB_TRY(mg_vcycle_run(c, &zc));
  k_mg_dot_rz<<<blocks, 256, 0, c->stream>>>(c->cg_r, L0.x[zc], L0.ny, L0.ld, m->scal, m->partials, 1);
  FSB_LAUNCHED(c);
  FSB_CUDA(c, cudaMemcpyAsync(p, L0.x[zc], sizeof(float) * (size_t)L0.ld * L0.ny,
                              cudaMemcpyDeviceToDevice, c->stream));
  MgScalars fin;
  for (;;)
  {
    k_mg_apply_dot<<<blocks, 256, 0, c->stream>>>(p, L0.code, q, L0.nx, L0.ny, L0.ld, L0.inv_h2,
                                                  m->scal, m->partials);
    FSB_LAUNCHED(c);
    k_mg_update<<<blocks, 256, 0, c->stream>>>(c->cg_x, c->cg_r, p, q, L0.ny, L0.ld, m->scal,
                                               m->partials, cap);
    FSB_LAUNCHED(c);
    FSB_CUDA(c, cudaMemcpyAsync(m->scal_h, m->scal, sizeof(MgScalars), cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    fin = *m->scal_h;
    if (fin.fail || fin.done) break;
    FSB_TRY(mg_vcycle_run(c, &zc));
    k_mg_dot_rz<<<blocks, 256, 0, c->stream>>>(c->cg_r, L0.x[zc], L0.ny, L0.ld, m->scal, m->partials, 0);
    FSB_LAUNCHED(c);
    k_mg_direction<<<blocks, 256, 0, c->stream>>>(p, L0.x[zc], L0.ny, L0.ld, m->scal);
    FSB_LAUNCHED(c);
  }
  const bool ok = !fin.fail && (float)fin.r2 < fin.thr;
  // an explicit user cap that was reached counts as "done" exactly like the Jacobi solve
  const bool capped = !fin.fail && c->max_iters >= 0 && fin.iter >= c->max_iters;
  if (!ok && !capped) return FSB_OK;
  c->iters = fin.iter;
  c->err = (fin.rhs2 == 0.0) ? 0.0f : std::sqrt((float)fin.r2 / (float)fin.rhs2);
  *converged = 1;
  return FSB_OK;
}
