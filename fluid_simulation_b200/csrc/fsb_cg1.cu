// Single-reduction Jacobi-PCG (fsb_cg1_kernels.cuh): launcher and host loop.  Opt-in
// (FSB_CG_MODE=single), single GPU, default off.  NOT yet run on a GPU in round 1: the per-group
// arithmetic and the whole solve loop are verified through the host emulation only.
#include <algorithm>
#include <cmath>

#include "fsb_internal.cuh"
#include "fsb_cg1_kernels.cuh"

struct fsb_cg1_state
{
  float *r1 = nullptr, *s1 = nullptr, *w0 = nullptr, *w1 = nullptr;
  Cg1Scalars* scal = nullptr;
  Cg1Scalars* scal_h = nullptr;
  double* partials = nullptr;
  int blocks = 0;
  size_t cells = 0;
};

namespace {

__global__ void __launch_bounds__(256)
k_cg1_sweep(const float* __restrict__ r_old, const float* __restrict__ s_old,
            const float* __restrict__ w_old, float* __restrict__ r_new, float* __restrict__ s_new,
            float* __restrict__ w_new, float* __restrict__ p, float* __restrict__ x,
            const uint8_t* __restrict__ code, int nx, int ny, int ld, const Cg1Coef k,
            Cg1Scalars* __restrict__ s, double* __restrict__ partials)
{
  if (s->done) return;
  const float alpha = s->init ? 0.0f : s->alpha;
  const float beta = s->init ? 0.0f : s->beta;
  double ag = 0.0, ad = 0.0, ar = 0.0;
  const int segs = (ld + 1023) / 1024;
  for (int w = blockIdx.x; w < segs * ny; w += gridDim.x)
  {
    const int j = w / segs;
    const int i0 = ((w - j * segs) * 256 + threadIdx.x) * 4;
    if (i0 < ld)
      cg1_group(r_old, s_old, w_old, r_new, s_new, w_new, p, x, code, nx, ny, ld, i0, j, alpha, beta,
                k, &ag, &ad, &ar);
  }
  const double tg = block_sum(ag);
  const double td = block_sum(ad);
  const double tr = block_sum(ar);
  if (threadIdx.x == 0)
  {
    partials[blockIdx.x] = tg;
    partials[gridDim.x + blockIdx.x] = td;
    partials[2 * gridDim.x + blockIdx.x] = tr;
  }
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&s->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double fold[3];
  for (int q = 0; q < 3; ++q)
  {
    double v = 0.0;
    const volatile double* part = partials + (size_t)q * gridDim.x;
    for (int t = threadIdx.x; t < (int)gridDim.x; t += blockDim.x) v += part[t];
    fold[q] = block_sum(v);
  }
  if (threadIdx.x == 0)
  {
    cg1_advance(s, fold[0], fold[1], fold[2]);
    s->ticket = 0;
  }
}

__global__ void k_cg1_init(Cg1Scalars* s, const CgScalars* cg)
{
  s->gamma = 0.0; s->delta = 0.0; s->r2 = cg->r2; s->rhs2 = cg->rhs2;
  s->alpha = 0.0f; s->beta = 0.0f; s->thr = cg->thr;
  s->iter = 0; s->done = 0; s->max_iters = cg->max_iters; s->init = 1; s->ticket = 0;
}

template <class T>
int cg1_alloc(fsb_ctx* c, T** p, size_t count)
{
  *p = nullptr;
  if (cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)) != cudaSuccess)
    return fsb_fail(c, FSB_ERR_NOMEM, "cudaMalloc of %zu bytes failed (single-reduction CG)", sizeof(T) * count);
  return FSB_OK;
}

} // namespace

void fsb_cg1_free(fsb_ctx* c)
{
  fsb_cg1_state* st = c->cg1;
  if (!st) return;
  cudaFree(st->r1); cudaFree(st->s1); cudaFree(st->w0); cudaFree(st->w1);
  cudaFree(st->scal); cudaFree(st->partials);
  if (st->scal_h) cudaFreeHost(st->scal_h);
  delete st;
  c->cg1 = nullptr;
}

// Called after k_cg_build (x = 0, r = b, stencil codes, |b|^2 and the threshold in the CG scalars).
int fsb_k_cg1_solve(fsb_ctx* c)
{
  const size_t cells = (size_t)c->ld * c->ny;
  if (!c->cg1 || c->cg1->cells != cells)
  {
    fsb_cg1_free(c);
    fsb_cg1_state* st = new fsb_cg1_state();
    c->cg1 = st;
    st->cells = cells;
    FSB_TRY(cg1_alloc(c, &st->r1, cells));
    FSB_TRY(cg1_alloc(c, &st->s1, cells));
    FSB_TRY(cg1_alloc(c, &st->w0, cells));
    FSB_TRY(cg1_alloc(c, &st->w1, cells));
    FSB_TRY(cg1_alloc(c, &st->scal, 1));
    st->blocks = (int)std::min<int64_t>((int64_t)c->sm_count * 8,
                                        std::max<int64_t>(1, (int64_t)fsb_div_up(c->ld, 1024) * c->ny));
    FSB_TRY(cg1_alloc(c, &st->partials, (size_t)3 * st->blocks));
    if (cudaMallocHost((void**)&st->scal_h, sizeof(Cg1Scalars)) != cudaSuccess)
      return fsb_fail(c, FSB_ERR_NOMEM, "cudaMallocHost failed");
  }
  fsb_cg1_state* st = c->cg1;
  Cg1Coef k;
  {
    const double dx2 = std::pow((double)c->dx, 2);
    k.off = (float)(1 / dx2);
    for (int n = 0; n < 5; ++n)
    {
      k.diag[n] = (float)(-n / dx2);
      k.invdiag[n] = (k.diag[n] != 0.0f) ? 1.0f / k.diag[n] : 1.0f;
    }
  }
  float* r[2] = {c->cg_r, st->r1};
  float* s[2] = {c->cg_p[1], st->s1};
  float* w[2] = {st->w0, st->w1};
  float* p = c->cg_p[0];
  const size_t bytes = sizeof(float) * cells;
  FSB_CUDA(c, cudaMemsetAsync(s[0], 0, bytes, c->stream));
  FSB_CUDA(c, cudaMemsetAsync(w[0], 0, bytes, c->stream));
  FSB_CUDA(c, cudaMemsetAsync(p, 0, bytes, c->stream));
  k_cg1_init<<<1, 1, 0, c->stream>>>(st->scal, c->scal);
  FSB_LAUNCHED(c);
  int cur = 0;
  Cg1Scalars fin;
  const int check_every = 16;
  for (int launched = 0;;)
  {
    for (int q = 0; q < check_every; ++q, ++launched)
    {
      k_cg1_sweep<<<st->blocks, 256, 0, c->stream>>>(r[cur], s[cur], w[cur], r[cur ^ 1], s[cur ^ 1],
                                                     w[cur ^ 1], p, c->cg_x, c->cg_code, c->nx, c->ny,
                                                     c->ld, k, st->scal, st->partials);
      FSB_LAUNCHED(c);
      cur ^= 1;
    }
    FSB_CUDA(c, cudaMemcpyAsync(st->scal_h, st->scal, sizeof(Cg1Scalars), cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    fin = *st->scal_h;
    if (fin.done) break;
    if (launched > 4 * (fin.max_iters + 2) + 64)
      return fsb_fail(c, FSB_ERR_CUDA, "single-reduction CG did not terminate");
  }
  c->iters = fin.iter;
  c->err = (fin.rhs2 == 0.0 || (float)fin.rhs2 == 0.0f) ? 0.0f : std::sqrt((float)fin.r2 / (float)fin.rhs2);
  return FSB_OK;
}
