// C ABI of libfsb.so (include/fsb.h): context life cycle, host<->HBM state
// transfer, stage entry points and the fused step drivers.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fsb_internal.cuh"

namespace {

std::string g_create_error;

__global__ void k_iota(int* __restrict__ dst, int64_t first, int64_t n)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[first + k] = (int)(first + k);
}

__global__ void k_fill_u8(uint8_t* __restrict__ dst, uint8_t v, int64_t n)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = v;
}

} // namespace

int fsb_fail(fsb_ctx* ctx, int code, const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err_msg = buf;
  else g_create_error = buf;
  return code;
}

static int prof_drain(fsb_ctx* c)
{
  if (c->prof_used == 0) return FSB_OK;
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < c->prof_used; ++k)
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->prof_ev[k][0], c->prof_ev[k][1]) == cudaSuccess)
    {
      c->prof_ms[c->prof_stage[k]] += ms;
      c->prof_calls[c->prof_stage[k]] += 1;
    }
  }
  c->prof_used = 0;
  return FSB_OK;
}

void fsb_prof_begin(fsb_ctx* c, int stage)
{
  if (!c->profiling) return;
  if (!c->prof_made)
  {
    for (int k = 0; k < fsb_ctx::kProfPool; ++k)
    {
      cudaEventCreate(&c->prof_ev[k][0]);
      cudaEventCreate(&c->prof_ev[k][1]);
    }
    c->prof_made = true;
  }
  if (c->prof_used == fsb_ctx::kProfPool) prof_drain(c);
  c->prof_stage[c->prof_used] = stage;
  cudaEventRecord(c->prof_ev[c->prof_used][0], c->stream);
}

void fsb_prof_end(fsb_ctx* c, int stage)
{
  if (!c->profiling) return;
  (void)stage;
  cudaEventRecord(c->prof_ev[c->prof_used][1], c->stream);
  c->prof_used++;
}

// ------------------------------------------------------------ allocation --
template <class T>
static int dev_alloc(fsb_ctx* c, T** p, size_t count, int fill_zero = 1)
{
  *p = nullptr;
  cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
  if (e != cudaSuccess)
    return fsb_fail(c, FSB_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", sizeof(T) * count,
                    cudaGetErrorString(e));
  if (fill_zero) FSB_CUDA(c, cudaMemsetAsync(*p, 0, sizeof(T) * (count ? count : 1), c->stream));
  return FSB_OK;
}

static int ensure_stage(fsb_ctx* c, size_t bytes)
{
  if (bytes <= c->stage_bytes) return FSB_OK;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  c->stage_bytes = 0;
  FSB_TRY(dev_alloc(c, (char**)&c->stage, bytes, 0));
  c->stage_bytes = bytes;
  return FSB_OK;
}

static int ensure_particle_capacity(fsb_ctx* c, int64_t need)
{
  if (need <= c->cap) return FSB_OK;
  if (need >= (int64_t)2147483647)
    return fsb_fail(c, FSB_ERR_INVALID, "particle count %lld exceeds int32 indexing",
                    (long long)need);
  int64_t ncap = c->cap ? c->cap : 1024;
  while (ncap < need) ncap *= 2;
  float4* np[2];
  int* no[2];
  for (int k = 0; k < 2; ++k)
  {
    FSB_TRY(dev_alloc(c, &np[k], (size_t)ncap, 0));
    FSB_TRY(dev_alloc(c, &no[k], (size_t)ncap, 0));
  }
  if (c->n > 0)
  {
    FSB_CUDA(c, cudaMemcpyAsync(np[c->pcur], c->part[c->pcur], sizeof(float4) * c->n,
                                cudaMemcpyDeviceToDevice, c->stream));
    FSB_CUDA(c, cudaMemcpyAsync(no[c->pcur], c->orig[c->pcur], sizeof(int) * c->n,
                                cudaMemcpyDeviceToDevice, c->stream));
  }
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 2; ++k)
  {
    if (c->part[k]) cudaFree(c->part[k]);
    if (c->orig[k]) cudaFree(c->orig[k]);
    c->part[k] = np[k];
    c->orig[k] = no[k];
  }
  if (c->sort_key) cudaFree(c->sort_key);
  if (c->sort_rank) cudaFree(c->sort_rank);
  if (c->sort_idx) cudaFree(c->sort_idx);
  FSB_TRY(dev_alloc(c, &c->sort_key, (size_t)ncap, 0));
  FSB_TRY(dev_alloc(c, &c->sort_rank, (size_t)ncap, 0));
  FSB_TRY(dev_alloc(c, &c->sort_idx, (size_t)ncap, 0));
  c->cap = ncap;
  return FSB_OK;
}

static float* pick_grid(fsb_ctx* c, int which)
{
  switch (which)
  {
  case FSB_U_FRONT: return fsb_uf(c);
  case FSB_V_FRONT: return fsb_vf(c);
  case FSB_U_BACK: return fsb_ub(c);
  case FSB_V_BACK: return fsb_vb(c);
  case FSB_U_PREV: return c->u_prev;
  case FSB_V_PREV: return c->v_prev;
  case FSB_U_DIFF: return c->u_diff;
  case FSB_V_DIFF: return c->v_diff;
  }
  return nullptr;
}

// The fused FLIP / PIC-FLIP steps interpolate (front - previous) tap by tap
// and leave the diff buffer unmaterialised; it is produced on demand.
static int flush_diff(fsb_ctx* c)
{
  if (!c->diff_pending) return FSB_OK;
  c->diff_pending = false;
  return fsb_k_update_diff(c);
}

#define CHECK_CTX(ctx)                                                         \
  do                                                                           \
  {                                                                            \
    if (!(ctx)) return fsb_fail(nullptr, FSB_ERR_INVALID, "null context");     \
    cudaError_t _e = cudaSetDevice((ctx)->device);                             \
    if (_e != cudaSuccess)                                                     \
      return fsb_fail((ctx), FSB_ERR_CUDA, "cudaSetDevice(%d) failed: %s",     \
                      (ctx)->device, cudaGetErrorString(_e));                  \
  } while (0)

extern "C" {

const char* fsb_version(void) { return "fsb 0.1 (sm_100a)"; }

const char* fsb_last_error(const fsb_ctx* ctx)
{
  return ctx ? ctx->err_msg.c_str() : g_create_error.c_str();
}

int fsb_create(fsb_ctx** out, int size_x, int size_y, float length_x, float length_y,
               float density, float pic_ratio, int device)
{
  if (!out) return fsb_fail(nullptr, FSB_ERR_INVALID, "null output pointer");
  *out = nullptr;
  if (size_x < 3 || size_y < 3)
    return fsb_fail(nullptr, FSB_ERR_INVALID, "grid must be at least 3x3 (got %dx%d)", size_x,
                    size_y);
  if ((int64_t)size_x * size_y >= (int64_t)2147483647 / 2)
    return fsb_fail(nullptr, FSB_ERR_INVALID, "grid too large for int32 cell indexing");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fsb_fail(nullptr, FSB_ERR_CUDA,
                    "no CUDA device available (%s); libfsb has no CPU fallback",
                    cudaGetErrorString(e));
  if (device < 0 || device >= n_dev)
    return fsb_fail(nullptr, FSB_ERR_INVALID, "device %d out of range (0..%d)", device, n_dev - 1);
  e = cudaSetDevice(device);
  if (e != cudaSuccess)
    return fsb_fail(nullptr, FSB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess)
    return fsb_fail(nullptr, FSB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fsb_fail(nullptr, FSB_ERR_CUDA,
                    "device %d is sm_%d%d; libfsb is built for sm_100a only", device, prop.major,
                    prop.minor);

  fsb_ctx* c = new fsb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->nx = size_x;
  c->ny = size_y;
  c->ld = (size_x + 31) / 32 * 32;
  c->dx = length_x / size_x; // src/MacGrid.cpp:8
  c->dy = length_y / size_y;
  c->pool_dx = c->dx; // src/FluidSolver.cpp:56-65: the solver's pool copy passes deltaX twice
  c->pool_dy = c->dx;
  c->pool_nx = size_x;
  c->pool_ny = size_y;
  c->density = density;
  c->pic_ratio = pic_ratio;
  c->grav_x = 0.0f;
  c->grav_y = (float)-9.82; // src/FluidSolver.cpp:232
  if (const char* e = getenv("FSB_STAGE_KERNELS")) c->stage_v1 = (strcmp(e, "v1") == 0);
  if (const char* e = getenv("FSB_SL_ATOMIC")) c->sl_atomic = atoi(e) != 0;
  if (const char* e = getenv("FSB_CANON_PER")) c->canon_per = atoi(e) == 1 ? 1 : 2;
  if (const char* e = getenv("FSB_P2G_PIPE")) c->p2g_pipe = atoi(e) != 0;
  if (const char* e = getenv("FSB_SORT_PER")) { const int v = atoi(e); c->sort_per = (v == 1 || v == 4) ? v : 2; }
  if (const char* e = getenv("FSB_G2P_PER")) { const int v = atoi(e); c->g2p_per = (v == 1 || v == 4) ? v : 2; }
  if (const char* e = getenv("FSB_BUILD_BLOCKS_PER_SM")) c->build_blocks_per_sm = std::max(1, std::min(32, atoi(e)));

  if (const char* e = getenv("FSB_MG_MAX_ITERS")) c->mg_max_iters = std::max(1, atoi(e));
  if (const char* e = getenv("FSB_MG_SWEEPS")) c->mg_sweeps = std::max(1, std::min(8, atoi(e)));
  if (const char* e = getenv("FSB_MG_GRAPH")) c->mg_graph = atoi(e) != 0;
  if (const char* e = getenv("FSB_MG_STOP")) c->mg_stop = std::max(1, std::min(32, atoi(e)));
  if (const char* e = getenv("FSB_MG_RENORM")) c->mg_renorm = atoi(e) != 0;
  memset(c->prof_ms, 0, sizeof c->prof_ms);
  memset(c->prof_calls, 0, sizeof c->prof_calls);

  int rc = FSB_OK;
  auto fail = [&](int code) {
    g_create_error = c->err_msg;
    fsb_destroy(c);
    return code;
  };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
    return fail(fsb_fail(c, FSB_ERR_CUDA, "cudaStreamCreate failed"));
  c->own_stream = true;
  cudaEventCreate(&c->timer_ev[0]);
  cudaEventCreate(&c->timer_ev[1]);

  const size_t cells = (size_t)c->ld * c->ny;
  for (int k = 0; k < 2 && rc == FSB_OK; ++k)
  {
    if ((rc = dev_alloc(c, &c->u[k], cells)) != FSB_OK) break;
    if ((rc = dev_alloc(c, &c->v[k], cells)) != FSB_OK) break;
    if ((rc = dev_alloc(c, &c->mask_x[k], cells)) != FSB_OK) break;
    if ((rc = dev_alloc(c, &c->mask_y[k], cells)) != FSB_OK) break;
  }
  if (rc == FSB_OK) rc = dev_alloc(c, &c->u_prev, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->v_prev, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->u_diff, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->v_diff, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cell, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_x, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_r, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_r2, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_p[0], cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_p[1], cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cg_code, cells);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cell_start, (size_t)size_x * size_y + 1);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->cell_count, (size_t)size_x * size_y);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->scan_block, (size_t)size_x * size_y / 4096 + 2);
  if (rc == FSB_OK) rc = dev_alloc(c, &c->scal, 1);
  if (rc != FSB_OK) return fail(rc);
  if (cudaMallocHost((void**)&c->scal_h, 2 * sizeof(CgScalars)) != cudaSuccess)
    return fail(fsb_fail(c, FSB_ERR_NOMEM, "cudaMallocHost failed"));
  memset(c->scal_h, 0, 2 * sizeof(CgScalars));
  cudaEventCreateWithFlags(&c->cg_ev[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->cg_ev[1], cudaEventDisableTiming);
  // labels: pad columns SOLID, then the constructor's clearCellTypeBuffer (src/MacGrid.cpp:24)
  k_fill_u8<<<fsb_div_up((int64_t)cells, 256), 256, 0, c->stream>>>(c->cell, FSB_SOLID,
                                                                    (int64_t)cells);
  c->launches++;
  rc = fsb_k_classify(c); // n == 0: border SOLID, interior AIR
  if (rc != FSB_OK) return fail(rc);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess)
    return fail(fsb_fail(c, FSB_ERR_CUDA, "initialisation failed: %s",
                         cudaGetErrorString(cudaGetLastError())));
  *out = c;
  return FSB_OK;
}

void fsb_destroy(fsb_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int k = 0; k < 2; ++k)
  {
    cudaFree(c->u[k]); cudaFree(c->v[k]);
    cudaFree(c->mask_x[k]); cudaFree(c->mask_y[k]);
    cudaFree(c->part[k]); cudaFree(c->orig[k]);
  }
  cudaFree(c->u_prev); cudaFree(c->v_prev); cudaFree(c->u_diff); cudaFree(c->v_diff);
  cudaFree(c->cell); cudaFree(c->cg_x); cudaFree(c->cg_r); cudaFree(c->cg_r2); cudaFree(c->cg_p[0]); cudaFree(c->cg_p[1]);
  cudaFree(c->cg_code); cudaFree(c->cell_start); cudaFree(c->cell_count); cudaFree(c->scan_block);
  cudaFree(c->sort_key); cudaFree(c->sort_rank); cudaFree(c->sort_idx);
  for (int k = 0; k < c->n_ipc_opened; ++k) cudaIpcCloseMemHandle(c->ipc_opened[k]);
  cudaFree(c->peer_x_dev); cudaFree(c->mail_local);
  fsb_sl_free(c);
  fsb_mg_free(c);
  cudaFree(c->slab_buf_part); cudaFree(c->slab_buf_orig); cudaFree(c->slab_ctr);
  cudaFree(c->cg_tile_flags); cudaFree(c->cg_tile_list);
  cudaFree(c->partials); cudaFree(c->scal); cudaFree(c->stage);
  if (c->scal_h) cudaFreeHost(c->scal_h);
  if (c->cg_graph) cudaGraphExecDestroy(c->cg_graph);
  if (c->cg_ev[0]) cudaEventDestroy(c->cg_ev[0]);
  if (c->cg_ev[1]) cudaEventDestroy(c->cg_ev[1]);
  if (c->timer_ev[0]) cudaEventDestroy(c->timer_ev[0]);
  if (c->timer_ev[1]) cudaEventDestroy(c->timer_ev[1]);
  if (c->prof_made)
    for (int k = 0; k < fsb_ctx::kProfPool; ++k)
    {
      cudaEventDestroy(c->prof_ev[k][0]);
      cudaEventDestroy(c->prof_ev[k][1]);
    }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int fsb_set_stream(fsb_ctx* c, void* cuda_stream)
{
  CHECK_CTX(c);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return FSB_OK;
}

int fsb_synchronize(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}

int fsb_size_x(const fsb_ctx* c) { return c ? c->nx : 0; }
int fsb_size_y(const fsb_ctx* c) { return c ? c->ny : 0; }
float fsb_delta_x(const fsb_ctx* c) { return c ? c->dx : 0.0f; }
float fsb_delta_y(const fsb_ctx* c) { return c ? c->dy : 0.0f; }

int fsb_set_cg(fsb_ctx* c, int max_iters, float tol)
{
  CHECK_CTX(c);
  c->max_iters = max_iters;
  c->tol = tol;
  return FSB_OK;
}
int fsb_get_cg_info(const fsb_ctx* c, int* iterations, float* error)
{
  if (!c) return FSB_ERR_INVALID;
  if (iterations) *iterations = c->iters;
  if (error) *error = c->err;
  return FSB_OK;
}
int fsb_set_preconditioner(fsb_ctx* c, int kind)
{
  CHECK_CTX(c);
  if (kind != FSB_PRECOND_JACOBI && kind != FSB_PRECOND_MULTIGRID)
    return fsb_fail(c, FSB_ERR_INVALID, "unknown preconditioner %d", kind);
  c->precond = kind;
  return FSB_OK;
}
int fsb_set_pic_ratio(fsb_ctx* c, float pic_ratio)
{
  CHECK_CTX(c);
  const float t = pic_ratio < 0.0f ? 0.0f : pic_ratio; // CLAMP(pic_ratio, 0, 1)
  c->pic_ratio = t > 1.0f ? 1.0f : t;
  return FSB_OK;
}
int fsb_set_density(fsb_ctx* c, float density)
{
  CHECK_CTX(c);
  c->density = density;
  return FSB_OK;
}
int fsb_set_integrator(fsb_ctx* c, int integrator)
{
  CHECK_CTX(c);
  if (integrator != FSB_INTEGRATOR_RK3 && integrator != FSB_INTEGRATOR_EULER)
    return fsb_fail(c, FSB_ERR_INVALID, "unknown integrator %d", integrator);
  c->integrator = integrator;
  return FSB_OK;
}
int fsb_set_pool(fsb_ctx* c, int size_x, int size_y, float delta_x, float delta_y)
{
  CHECK_CTX(c);
  c->pool_nx = size_x;
  c->pool_ny = size_y;
  c->pool_dx = delta_x;
  c->pool_dy = delta_y;
  return FSB_OK;
}
int fsb_set_gravity(fsb_ctx* c, float ax, float ay)
{
  CHECK_CTX(c);
  c->grav_x = ax;
  c->grav_y = ay;
  return FSB_OK;
}

// ---------------------------------------------------------------- particles
// On a slab-partitioned context (fsb_slab_configure with world > 1) the index map holds GLOBAL
// particle ids (and -1 for retired ghosts), not a permutation of 0 .. n-1: the calls that treat it as
// one -- or that hand out ids starting at the local count -- are refused; fsb_slab_get / _add are
// their slab forms.
#define REFUSE_ON_SLAB(c, what) \
  return fsb_fail((c), FSB_ERR_INVALID, what " on a slab-partitioned context: use the fsb_slab_* calls")

int fsb_append_particles(fsb_ctx* c, const float* aos4, int64_t n)
{
  CHECK_CTX(c);
  if (c->slab_partitioned) REFUSE_ON_SLAB(c, "fsb_append_particles");
  if (n < 0 || (n > 0 && !aos4)) return fsb_fail(c, FSB_ERR_INVALID, "bad particle buffer");
  if (n == 0) return FSB_OK;
  FSB_TRY(ensure_particle_capacity(c, c->n + n));
  FSB_CUDA(c, cudaMemcpyAsync(c->part[c->pcur] + c->n, aos4, sizeof(float4) * n,
                              cudaMemcpyHostToDevice, c->stream));
  k_iota<<<fsb_div_up(n, 256), 256, 0, c->stream>>>(c->orig[c->pcur], c->n, n);
  FSB_LAUNCHED(c);
  c->n += n;
  c->sort_valid = false;
  return FSB_OK;
}
int fsb_set_particles(fsb_ctx* c, const float* aos4, int64_t n)
{
  CHECK_CTX(c);
  c->n = 0;
  c->slab_partitioned = false; // a fresh, whole set (to be distributed again if the context is a slab)
  c->sort_valid = false;
  return fsb_append_particles(c, aos4, n);
}
int64_t fsb_num_particles(const fsb_ctx* c) { return c ? c->n : 0; }
int fsb_get_particles(fsb_ctx* c, float* aos4)
{
  CHECK_CTX(c);
  if (c->slab_partitioned) REFUSE_ON_SLAB(c, "fsb_get_particles");
  if (c->n == 0) return FSB_OK;
  if (!aos4) return fsb_fail(c, FSB_ERR_INVALID, "null particle buffer");
  // part[pcur^1] is scratch outside the sort
  FSB_TRY(fsb_k_unpermute(c, c->part[c->pcur ^ 1]));
  FSB_CUDA(c, cudaMemcpyAsync(aos4, c->part[c->pcur ^ 1], sizeof(float4) * c->n,
                              cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}

int fsb_emit_source(fsb_ctx* c, float x_min, float x_max, float y_min, float y_max, float delta_x,
                    float delta_y, float vel_x, float vel_y, int64_t* n_added)
{
  CHECK_CTX(c);
  if (n_added) *n_added = 0;
  if (c->slab_partitioned) REFUSE_ON_SLAB(c, "fsb_emit_source");
  // src/FluidDomain.cpp:37-41: increments are formed in double and rounded;
  // the coordinates are accumulated by repeated float addition
  const float x_incr = (float)(delta_x / 2.5);
  const float y_incr = (float)(delta_y / 2.5);
  if (!(x_incr > 0.0f) || !(y_incr > 0.0f))
    return fsb_fail(c, FSB_ERR_INVALID, "source spacing must be positive");
  std::vector<float> xs, ys;
  for (float y = y_min; y < y_max; y += y_incr) ys.push_back(y);
  for (float x = x_min; x < x_max; x += x_incr) xs.push_back(x);
  const int64_t total = (int64_t)xs.size() * (int64_t)ys.size();
  if (total == 0) return FSB_OK;
  FSB_TRY(ensure_particle_capacity(c, c->n + total));
  FSB_TRY(ensure_stage(c, sizeof(float) * (xs.size() + ys.size())));
  float* xs_d = c->stage;
  float* ys_d = c->stage + xs.size();
  FSB_CUDA(c, cudaMemcpyAsync(xs_d, xs.data(), sizeof(float) * xs.size(), cudaMemcpyHostToDevice,
                              c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(ys_d, ys.data(), sizeof(float) * ys.size(), cudaMemcpyHostToDevice,
                              c->stream));
  FSB_TRY(fsb_k_emit_source_dev(c, c->n, xs_d, ys_d, (int64_t)xs.size(), (int64_t)ys.size(), vel_x,
                                vel_y));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream)); // xs/ys are host temporaries
  c->n += total;
  c->sort_valid = false;
  if (n_added) *n_added = total;
  return FSB_OK;
}

// -------------------------------------------------------------------- grids
int fsb_set_grid(fsb_ctx* c, int which, const float* src)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  float* g = pick_grid(c, which);
  if (!g || !src) return fsb_fail(c, FSB_ERR_INVALID, "bad grid selector %d or null buffer", which);
  FSB_CUDA(c, cudaMemcpy2DAsync(g, sizeof(float) * c->ld, src, sizeof(float) * c->nx,
                                sizeof(float) * c->nx, c->ny, cudaMemcpyHostToDevice, c->stream));
  return FSB_OK;
}
int fsb_get_grid(fsb_ctx* c, int which, float* dst)
{
  CHECK_CTX(c);
  if (which == FSB_U_DIFF || which == FSB_V_DIFF) FSB_TRY(flush_diff(c));
  float* g = pick_grid(c, which);
  if (!g || !dst) return fsb_fail(c, FSB_ERR_INVALID, "bad grid selector %d or null buffer", which);
  FSB_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(float) * c->nx, g, sizeof(float) * c->ld,
                                sizeof(float) * c->nx, c->ny, cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
int fsb_set_cell_types(fsb_ctx* c, const uint8_t* src)
{
  CHECK_CTX(c);
  if (!src) return fsb_fail(c, FSB_ERR_INVALID, "null buffer");
  FSB_CUDA(c, cudaMemcpy2DAsync(c->cell, c->ld, src, c->nx, c->nx, c->ny, cudaMemcpyHostToDevice,
                                c->stream));
  return FSB_OK;
}
int fsb_get_cell_types(fsb_ctx* c, uint8_t* dst)
{
  CHECK_CTX(c);
  if (!dst) return fsb_fail(c, FSB_ERR_INVALID, "null buffer");
  FSB_CUDA(c, cudaMemcpy2DAsync(dst, c->nx, c->cell, c->ld, c->nx, c->ny, cudaMemcpyDeviceToHost,
                                c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
int fsb_get_pressure(fsb_ctx* c, float* dst)
{
  CHECK_CTX(c);
  if (!dst) return fsb_fail(c, FSB_ERR_INVALID, "null buffer");
  if (!c->pressure_valid)
  {
    memset(dst, 0, sizeof(float) * (size_t)c->nx * c->ny);
    return FSB_OK;
  }
  FSB_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(float) * c->nx, c->cg_x, sizeof(float) * c->ld,
                                sizeof(float) * c->nx, c->ny, cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}

// ------------------------------------------------------------------- stages
static int square_cells(fsb_ctx* c)
{
  // the particle-to-grid gather window assumes the reference's own validity
  // condition (src/FluidSolver.cpp:89-97)
  if (!(c->nx == c->pool_nx && c->ny == c->pool_ny && std::fabs(c->dx - c->pool_dx) < 0.0000001 &&
        std::fabs(c->dy - c->pool_dy) < 0.0000001))
    return fsb_fail(c, FSB_ERR_INVALID,
                    "Memory pool and fluid domain does not match. Initialize the fluid solver "
                    "with a memory pool corresponding to the right fluid domain!");
  return FSB_OK;
}

int fsb_classify_cells(fsb_ctx* c)
{
  CHECK_CTX(c);
  return fsb_k_classify(c);
}
int fsb_clear_cell_types(fsb_ctx* c)
{
  CHECK_CTX(c);
  return fsb_k_clear_labels(c);
}
int fsb_swap_velocity_buffers(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  c->front ^= 1;
  return FSB_OK;
}
int fsb_p2g_spread(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_TRY(square_cells(c));
  FSB_TRY(flush_diff(c));
  return fsb_k_p2g(c);
}
int fsb_save_previous(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  return fsb_k_save_previous(c);
}
int fsb_add_acceleration(fsb_ctx* c, float ax, float ay, float dt)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  return fsb_k_add_acceleration(c, ax, ay, dt);
}
int fsb_enforce_dirichlet(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  return fsb_k_enforce_dirichlet(c);
}
int fsb_extend_velocity(fsb_ctx* c, int n_iterations)
{
  CHECK_CTX(c);
  if (n_iterations < 0) return fsb_fail(c, FSB_ERR_INVALID, "negative iteration count");
  FSB_TRY(flush_diff(c));
  return fsb_k_extend_velocity(c, n_iterations);
}
int fsb_extend_velocity_averaging(fsb_ctx* c, int n_iterations)
{
  CHECK_CTX(c);
  if (n_iterations < 0) return fsb_fail(c, FSB_ERR_INVALID, "negative iteration count");
  FSB_TRY(flush_diff(c));
  return fsb_k_extend_velocity_avg(c, n_iterations);
}
int fsb_pressure_solve(fsb_ctx* c, float density, float dt)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  return fsb_k_pressure_solve(c, density, dt);
}
int fsb_update_diff(fsb_ctx* c)
{
  CHECK_CTX(c);
  c->diff_pending = false;
  return fsb_k_update_diff(c);
}
int fsb_g2p(fsb_ctx* c, int mode, float pic_ratio)
{
  CHECK_CTX(c);
  if (mode < FSB_G2P_PIC || mode > FSB_G2P_PICFLIP)
    return fsb_fail(c, FSB_ERR_INVALID, "unknown g2p mode %d", mode);
  FSB_TRY(flush_diff(c));
  return fsb_k_g2p(c, mode, pic_ratio);
}
int fsb_advect_particles(fsb_ctx* c, float dt, int ensure_outside_obstacles)
{
  CHECK_CTX(c);
  return fsb_k_advect_particles(c, dt, ensure_outside_obstacles);
}
int fsb_advect_velocity_sl(fsb_ctx* c, float dt)
{
  CHECK_CTX(c);
  return fsb_k_advect_velocity_sl(c, dt);
}
int fsb_advect_particles_grid(fsb_ctx* c, float dt)
{
  CHECK_CTX(c);
  return fsb_k_advect_particles_grid(c, dt);
}

int fsb_add_external_force(fsb_ctx* c, float fx, float fy, float dt)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  // :266-271: vel + F / density * dt, i.e. an acceleration of F / density (one fp32 division)
  return fsb_k_add_acceleration(c, fx / c->density, fy / c->density, dt);
}
int fsb_p2g_gather(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_TRY(flush_diff(c));
  return fsb_k_p2g_gather(c);
}

// -------------------------------------------------------------------- steps
int fsb_step(fsb_ctx* c, int kind, float dt)
{
  CHECK_CTX(c);
  if (kind < FSB_STEP_SEMILAGRANGIAN || kind > FSB_STEP_PICFLIP)
    return fsb_fail(c, FSB_ERR_INVALID, "unknown step kind %d", kind);
  FSB_TRY(square_cells(c)); // validate(), src/FluidSolver.cpp:89-107
  const float gx = c->grav_x, gy = c->grav_y;
  if (kind == FSB_STEP_SEMILAGRANGIAN)
  {
    // src/FluidSolver.cpp:99-134
    FSB_TRY(flush_diff(c));
    FSB_TRY(fsb_k_classify(c));
    FSB_TRY(fsb_k_advect_velocity_sl(c, dt));
    if (c->stage_v1)
    {
      FSB_TRY(fsb_k_add_acceleration(c, gx, gy, dt));
      FSB_TRY(fsb_k_enforce_dirichlet(c));
      FSB_TRY(fsb_k_pressure_solve(c, c->density, dt));
      FSB_TRY(fsb_k_enforce_dirichlet(c));
    }
    else
    {
      // gravity + walls in one pass; the walls after the projection ride on the velocity patch
      FSB_TRY(fsb_k_prev_gravity_dirichlet(c, gx, gy, dt, 0));
      FSB_TRY(fsb_k_pressure_solve(c, c->density, dt, true));
    }
    FSB_TRY(fsb_k_advect_particles_grid(c, dt));
    return FSB_OK;
  }
  // src/FluidSolver.cpp:136-251
  if (kind == FSB_STEP_PIC) FSB_TRY(flush_diff(c));
  if (c->stage_v1 || c->sort_valid || c->n == 0) FSB_TRY(fsb_k_classify(c));
  else
  {
    // classification rides on the cell sort's counting pass: the particle set is read once
    FSB_TRY(fsb_k_classify_reset(c));
    FSB_TRY(fsb_k_sort_particles(c, true));
  }
  FSB_TRY(fsb_k_p2g(c));
  // updatePreviousVelocityBuffer + addExternalAcceleration + enforceDirichlet in one pass
  FSB_TRY(fsb_k_prev_gravity_dirichlet(c, gx, gy, dt, kind != FSB_STEP_PIC));
  FSB_TRY(fsb_k_extend_velocity(c, 2));
  if (c->stage_v1)
  {
    FSB_TRY(fsb_k_pressure_solve(c, c->density, dt));
    FSB_TRY(fsb_k_enforce_dirichlet(c));
  }
  else FSB_TRY(fsb_k_pressure_solve(c, c->density, dt, true));
  if (kind == FSB_STEP_PIC)
  {
    FSB_TRY(fsb_k_g2p_advect(c, FSB_G2P_PIC, 0.0f, dt, 0));
  }
  else
  {
    // updateVelocityDiffBuffer is folded into the transfer (front - previous
    // per tap); the buffer itself is produced if somebody asks for it
    c->diff_pending = true;
    if (kind == FSB_STEP_FLIP) FSB_TRY(fsb_k_g2p_advect(c, FSB_G2P_FLIP, 0.0f, dt, 0));
    else FSB_TRY(fsb_k_g2p_advect(c, FSB_G2P_PICFLIP, c->pic_ratio, dt, 1));
  }
  return FSB_OK;
}

// ----------------------------------------------------------- particle slabs
// One process per GPU, full grids on every rank, particles partitioned by row slab with one
// ghost row from each neighbour.  A step is three device phases with two exchanges in between
// (the transport -- NCCL in fluid_simulation_b200/sharding.py, plain copies in the in-process
// tests -- is the caller's):
//   ghosts in -> fsb_slab_step_a (labels + P2G of the own rows) -> exchange label / u / v rows ->
//   fsb_slab_step_b (grid passes, pressure solve) -> fsb_slab_step_c (G2P, advection) -> migrate
int fsb_slab_configure(fsb_ctx* c, int rank, int world)
{
  CHECK_CTX(c);
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || c->ny < 4 * world)
    return fsb_fail(c, FSB_ERR_INVALID, "bad slab arguments (rank %d of %d)", rank, world);
  c->slab_world = world;
  c->slab_rank = rank;
  c->slab_lo = (int)((int64_t)c->ny * rank / world);
  c->slab_hi = (int)((int64_t)c->ny * (rank + 1) / world);
  c->slab_grouped = false;
  if (!c->slab_ctr) FSB_TRY(dev_alloc(c, &c->slab_ctr, (size_t)2 * kMaxRanks));
  return FSB_OK;
}
int fsb_slab_rows(const fsb_ctx* c, int* row_lo, int* row_hi)
{
  if (!c) return FSB_ERR_INVALID;
  if (row_lo) *row_lo = c->slab_world > 1 ? c->slab_lo : 0;
  if (row_hi) *row_hi = c->slab_world > 1 ? c->slab_hi : c->ny;
  return FSB_OK;
}
// appends particles with their global ids (host or device pointers)
int fsb_slab_add(fsb_ctx* c, const float* aos4, const int32_t* ids, int64_t n)
{
  CHECK_CTX(c);
  if (n < 0 || (n > 0 && (!aos4 || !ids))) return fsb_fail(c, FSB_ERR_INVALID, "bad particle buffers");
  if (n == 0) return FSB_OK;
  FSB_TRY(ensure_particle_capacity(c, c->n + n));
  FSB_CUDA(c, cudaMemcpyAsync(c->part[c->pcur] + c->n, aos4, sizeof(float4) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(c->orig[c->pcur] + c->n, ids, sizeof(int) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream)); // the caller may reuse its buffers
  c->n += n;
  c->sort_valid = false;
  c->slab_grouped = false;
  c->slab_partitioned = true; // ids given by the caller: global ids
  return FSB_OK;
}
int fsb_slab_sort_out(fsb_ctx* c, int64_t* counts)
{
  CHECK_CTX(c);
  if (!counts) return fsb_fail(c, FSB_ERR_INVALID, "null counts");
  if (!c->slab_ctr) return fsb_fail(c, FSB_ERR_INVALID, "call fsb_slab_configure first");
  return fsb_k_slab_sort_out(c, counts);
}
int fsb_slab_take(fsb_ctx* c, int dest, float* aos4, int32_t* ids)
{
  CHECK_CTX(c);
  if (!c->slab_grouped || dest < 0 || dest >= c->slab_world)
    return fsb_fail(c, FSB_ERR_INVALID, "fsb_slab_take without a preceding fsb_slab_sort_out");
  int64_t off = 0;
  for (int q = 0; q < dest; ++q) off += c->slab_count[q];
  const int64_t n = c->slab_count[dest];
  if (n == 0) return FSB_OK;
  if (!aos4 || !ids) return fsb_fail(c, FSB_ERR_INVALID, "null buffers");
  FSB_CUDA(c, cudaMemcpyAsync(aos4, c->part[c->pcur] + off, sizeof(float4) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(ids, c->orig[c->pcur] + off, sizeof(int) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
int fsb_slab_keep_own(fsb_ctx* c)
{
  CHECK_CTX(c);
  if (!c->slab_grouped) return fsb_fail(c, FSB_ERR_INVALID, "fsb_slab_keep_own without fsb_slab_sort_out");
  int64_t off = 0;
  for (int q = 0; q < c->slab_rank; ++q) off += c->slab_count[q];
  const int64_t n = c->slab_count[c->slab_rank];
  if (n > 0 && off > 0)
  {
    // the other buffer is scratch outside the sort
    FSB_CUDA(c, cudaMemcpyAsync(c->part[c->pcur ^ 1], c->part[c->pcur] + off, sizeof(float4) * n,
                                cudaMemcpyDeviceToDevice, c->stream));
    FSB_CUDA(c, cudaMemcpyAsync(c->orig[c->pcur ^ 1], c->orig[c->pcur] + off, sizeof(int) * n,
                                cudaMemcpyDeviceToDevice, c->stream));
    c->pcur ^= 1;
  }
  c->n = n;
  c->sort_valid = false;
  c->slab_grouped = false;
  if (c->slab_world > 1) c->slab_partitioned = true;
  return FSB_OK;
}
int fsb_slab_boundary(fsb_ctx* c, int side, int64_t* n)
{
  CHECK_CTX(c);
  if (!n || (side != 0 && side != 1)) return fsb_fail(c, FSB_ERR_INVALID, "bad arguments");
  if (!c->slab_ctr) return fsb_fail(c, FSB_ERR_INVALID, "call fsb_slab_configure first");
  FSB_TRY(fsb_k_slab_row_select(c, side == 0 ? c->slab_lo : c->slab_hi - 1, n));
  c->slab_sel = *n;
  return FSB_OK;
}
int fsb_slab_boundary_take(fsb_ctx* c, float* aos4, int32_t* ids)
{
  CHECK_CTX(c);
  const int64_t n = c->slab_sel;
  if (n == 0) return FSB_OK;
  if (!aos4 || !ids) return fsb_fail(c, FSB_ERR_INVALID, "null buffers");
  FSB_CUDA(c, cudaMemcpyAsync(aos4, c->slab_buf_part, sizeof(float4) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(ids, c->slab_buf_orig, sizeof(int) * n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
// this rank's particles with their global ids, in device order
int fsb_slab_get(fsb_ctx* c, float* aos4, int32_t* ids)
{
  CHECK_CTX(c);
  if (c->n == 0) return FSB_OK;
  if (!aos4 || !ids) return fsb_fail(c, FSB_ERR_INVALID, "null buffers");
  FSB_CUDA(c, cudaMemcpyAsync(aos4, c->part[c->pcur], sizeof(float4) * c->n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(ids, c->orig[c->pcur], sizeof(int) * c->n, cudaMemcpyDefault, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
// rows [row_lo, row_hi) of a MacGrid buffer (which = FSB_U_FRONT ..) or of the labels
// (which = FSB_ROWS_LABELS), dense, host or device pointer
int fsb_get_rows(fsb_ctx* c, int which, int row_lo, int row_hi, void* dst)
{
  CHECK_CTX(c);
  if (row_lo < 0 || row_hi > c->ny || row_lo > row_hi || !dst) return fsb_fail(c, FSB_ERR_INVALID, "bad row range");
  if (row_lo == row_hi) return FSB_OK;
  if (which == FSB_ROWS_LABELS)
  {
    FSB_CUDA(c, cudaMemcpy2DAsync(dst, c->nx, c->cell + (size_t)row_lo * c->ld, c->ld, c->nx,
                                  row_hi - row_lo, cudaMemcpyDefault, c->stream));
  }
  else
  {
    if (which == FSB_U_DIFF || which == FSB_V_DIFF) FSB_TRY(flush_diff(c));
    float* g = pick_grid(c, which);
    if (!g) return fsb_fail(c, FSB_ERR_INVALID, "bad grid selector %d", which);
    FSB_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(float) * c->nx, g + (size_t)row_lo * c->ld,
                                  sizeof(float) * c->ld, sizeof(float) * c->nx, row_hi - row_lo,
                                  cudaMemcpyDefault, c->stream));
  }
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}
int fsb_set_rows(fsb_ctx* c, int which, int row_lo, int row_hi, const void* src)
{
  CHECK_CTX(c);
  if (row_lo < 0 || row_hi > c->ny || row_lo > row_hi || !src) return fsb_fail(c, FSB_ERR_INVALID, "bad row range");
  if (row_lo == row_hi) return FSB_OK;
  if (which == FSB_ROWS_LABELS)
  {
    FSB_CUDA(c, cudaMemcpy2DAsync(c->cell + (size_t)row_lo * c->ld, c->ld, src, c->nx, c->nx,
                                  row_hi - row_lo, cudaMemcpyDefault, c->stream));
  }
  else
  {
    FSB_TRY(flush_diff(c));
    float* g = pick_grid(c, which);
    if (!g) return fsb_fail(c, FSB_ERR_INVALID, "bad grid selector %d", which);
    FSB_CUDA(c, cudaMemcpy2DAsync(g + (size_t)row_lo * c->ld, sizeof(float) * c->ld, src,
                                  sizeof(float) * c->nx, sizeof(float) * c->nx, row_hi - row_lo,
                                  cudaMemcpyDefault, c->stream));
  }
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FSB_OK;
}

// phase A of a PIC / FLIP / PIC-FLIP step: classification from the local particles (own + ghost
// rows) and P2G of this rank's rows
int fsb_slab_step_a(fsb_ctx* c, int kind)
{
  CHECK_CTX(c);
  if (kind < FSB_STEP_SEMILAGRANGIAN || kind > FSB_STEP_PICFLIP) return fsb_fail(c, FSB_ERR_INVALID, "bad step kind %d", kind);
  FSB_TRY(square_cells(c));
  if (kind == FSB_STEP_SEMILAGRANGIAN)
  {
    // src/FluidSolver.cpp:99-134: the particles only mark cells here (no ghosts needed: a particle marks
    // its own cell), the velocity lives on the grid, identical on every rank
    FSB_TRY(flush_diff(c));
    return fsb_k_classify(c);
  }
  if (kind == FSB_STEP_PIC) FSB_TRY(flush_diff(c));
  if (c->stage_v1 || c->sort_valid || c->n == 0) FSB_TRY(fsb_k_classify(c));
  else
  {
    FSB_TRY(fsb_k_classify_reset(c));
    FSB_TRY(fsb_k_sort_particles(c, true));
  }
  return fsb_k_p2g(c);
}
// phase B: everything on the grid (identical on every rank once the rows are exchanged; the CG may
// additionally be sharded with fsb_shard_connect)
int fsb_slab_step_b(fsb_ctx* c, int kind, float dt)
{
  CHECK_CTX(c);
  if (kind < FSB_STEP_SEMILAGRANGIAN || kind > FSB_STEP_PICFLIP) return fsb_fail(c, FSB_ERR_INVALID, "bad step kind %d", kind);
  if (kind == FSB_STEP_SEMILAGRANGIAN)
  {
    // the deterministic gather form of the velocity advection (fsb_sl.cu) gives every rank the same bits
    if (c->sl_atomic || c->stage_v1)
      return fsb_fail(c, FSB_ERR_INVALID, "the float-atomics velocity advection (FSB_SL_ATOMIC / FSB_STAGE_V1: "
                                          "order-dependent sums) cannot run on slabs");
    FSB_TRY(fsb_k_advect_velocity_sl(c, dt));
    FSB_TRY(fsb_k_prev_gravity_dirichlet(c, c->grav_x, c->grav_y, dt, 0));
    return fsb_k_pressure_solve(c, c->density, dt, true);
  }
  FSB_TRY(fsb_k_prev_gravity_dirichlet(c, c->grav_x, c->grav_y, dt, kind != FSB_STEP_PIC));
  FSB_TRY(fsb_k_extend_velocity(c, 2));
  if (c->stage_v1)
  {
    FSB_TRY(fsb_k_pressure_solve(c, c->density, dt));
    return fsb_k_enforce_dirichlet(c);
  }
  return fsb_k_pressure_solve(c, c->density, dt, true);
}
// phase C: the ghosts are retired, then G2P + blend + advection of this rank's particles
int fsb_slab_step_c(fsb_ctx* c, int kind, float dt)
{
  CHECK_CTX(c);
  if (kind < FSB_STEP_SEMILAGRANGIAN || kind > FSB_STEP_PICFLIP) return fsb_fail(c, FSB_ERR_INVALID, "bad step kind %d", kind);
  if (c->slab_world > 1) FSB_TRY(fsb_k_slab_mark_ghosts(c));
  if (kind == FSB_STEP_SEMILAGRANGIAN) return fsb_k_advect_particles_grid(c, dt);
  if (kind == FSB_STEP_PIC) return fsb_k_g2p_advect(c, FSB_G2P_PIC, 0.0f, dt, 0);
  c->diff_pending = true;
  if (kind == FSB_STEP_FLIP) return fsb_k_g2p_advect(c, FSB_G2P_FLIP, 0.0f, dt, 0);
  return fsb_k_g2p_advect(c, FSB_G2P_PICFLIP, c->pic_ratio, dt, 1);
}

// --------------------------------------------------------------- state files
namespace {
struct StateHeader
{
  char magic[8];
  uint32_t version;
  int32_t nx, ny;
  float dx, dy, density, pic_ratio, grav_x, grav_y;
  int32_t integrator, max_iters;
  float tol, pool_dx, pool_dy;
  int32_t pool_nx, pool_ny;
  int64_t n_particles;
  uint8_t reserved[16];
};
static_assert(sizeof(StateHeader) == 96, "state header layout");
const char kStateMagic[8] = {'F', 'S', 'B', 'S', 'T', 'A', 'T', 'E'};
// closes the file on every exit path (the CUDA error macros return early)
struct FileCloser
{
  FILE* f;
  explicit FileCloser(FILE* file) : f(file) {}
  ~FileCloser() { if (f) fclose(f); }
  int close() { FILE* t = f; f = nullptr; return t ? fclose(t) : 0; }
};
} // namespace

int fsb_save_state(fsb_ctx* c, const char* path)
{
  CHECK_CTX(c);
  if (!path) return fsb_fail(c, FSB_ERR_INVALID, "null path");
  if (c->slab_partitioned) REFUSE_ON_SLAB(c, "fsb_save_state");
  FILE* f = fopen(path, "wb");
  if (!f) return fsb_fail(c, FSB_ERR_INVALID, "cannot open %s for writing", path);
  FileCloser guard(f);
  StateHeader hd;
  memset(&hd, 0, sizeof hd);
  memcpy(hd.magic, kStateMagic, 8);
  hd.version = 1;
  hd.nx = c->nx; hd.ny = c->ny; hd.dx = c->dx; hd.dy = c->dy;
  hd.density = c->density; hd.pic_ratio = c->pic_ratio; hd.grav_x = c->grav_x; hd.grav_y = c->grav_y;
  hd.integrator = c->integrator; hd.max_iters = c->max_iters; hd.tol = c->tol;
  hd.pool_dx = c->pool_dx; hd.pool_dy = c->pool_dy; hd.pool_nx = c->pool_nx; hd.pool_ny = c->pool_ny;
  hd.n_particles = c->n;
  const size_t cells = (size_t)c->nx * c->ny;
  std::vector<float> buf(std::max(cells, (size_t)c->n * 4));
  int rc = FSB_OK;
  bool ok = fwrite(&hd, sizeof hd, 1, f) == 1;
  if (ok)
  {
    std::vector<uint8_t> lab(cells);
    rc = fsb_get_cell_types(c, lab.data());
    ok = rc == FSB_OK && fwrite(lab.data(), 1, cells, f) == cells;
  }
  for (int w = FSB_U_FRONT; ok && w <= FSB_V_DIFF; ++w)
  {
    rc = fsb_get_grid(c, w, buf.data());
    ok = rc == FSB_OK && fwrite(buf.data(), sizeof(float), cells, f) == cells;
  }
  if (ok && c->n > 0)
  {
    // the device order (cell-sorted by the last step) and the map back to the caller's indices;
    // the next cell sort orders each cell by original index, so the continuation is bit-identical
    // whatever order the set is stored in -- the device order merely saves an un-permute pass
    FSB_CUDA(c, cudaMemcpyAsync(buf.data(), c->part[c->pcur], sizeof(float4) * c->n,
                                cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
    ok = fwrite(buf.data(), 4 * sizeof(float), (size_t)c->n, f) == (size_t)c->n;
    if (ok)
    {
      FSB_CUDA(c, cudaMemcpyAsync(buf.data(), c->orig[c->pcur], sizeof(int) * c->n,
                                  cudaMemcpyDeviceToHost, c->stream));
      FSB_CUDA(c, cudaStreamSynchronize(c->stream));
      ok = fwrite(buf.data(), sizeof(int), (size_t)c->n, f) == (size_t)c->n;
    }
  }
  ok = (guard.close() == 0) && ok;
  if (rc != FSB_OK) return rc;
  if (!ok) return fsb_fail(c, FSB_ERR_INVALID, "short write to %s", path);
  return FSB_OK;
}

// Transactional: the whole file is read into host memory and validated (header fields, sizes against
// the file length, the particle index map) BEFORE the context is touched; a rejected or truncated
// file leaves the running simulation exactly as it was.  No C++ exception crosses the C ABI.
static int load_state_impl(fsb_ctx* c, const char* path)
{
  FILE* f = fopen(path, "rb");
  if (!f) return fsb_fail(c, FSB_ERR_INVALID, "cannot open %s", path);
  FileCloser guard(f);
  StateHeader hd;
  if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, kStateMagic, 8) != 0 || hd.version != 1)
    return fsb_fail(c, FSB_ERR_INVALID, "%s is not a version-1 FSBSTATE file", path);
  if (hd.nx != c->nx || hd.ny != c->ny)
    return fsb_fail(c, FSB_ERR_INVALID, "%s holds a %dx%d domain, the context is %dx%d", path, hd.nx,
                    hd.ny, c->nx, c->ny);
  auto pos_finite = [](float v) { return std::isfinite(v) && v > 0.0f; };
  if (!pos_finite(hd.dx) || !pos_finite(hd.dy) || !pos_finite(hd.pool_dx) || !pos_finite(hd.pool_dy) ||
      !std::isfinite(hd.density) || !std::isfinite(hd.pic_ratio) || !std::isfinite(hd.grav_x) ||
      !std::isfinite(hd.grav_y) || !std::isfinite(hd.tol) || hd.pool_nx <= 0 || hd.pool_ny <= 0 ||
      (hd.integrator != FSB_INTEGRATOR_EULER && hd.integrator != FSB_INTEGRATOR_RK3))
    return fsb_fail(c, FSB_ERR_INVALID, "%s: header fields out of range", path);
  const size_t cells = (size_t)c->nx * c->ny;
  // the file length fixes how many particles it can hold: no allocation from an unchecked count
  if (fseek(f, 0, SEEK_END) != 0) return fsb_fail(c, FSB_ERR_INVALID, "%s: cannot seek", path);
  const long long file_len = ftell(f);
  const long long fixed = (long long)sizeof hd + (long long)cells * (1 + 8 * (long long)sizeof(float));
  if (hd.n_particles < 0 || hd.n_particles > INT32_MAX || file_len < fixed ||
      (file_len - fixed) != hd.n_particles * (long long)(4 * sizeof(float) + sizeof(int)))
    return fsb_fail(c, FSB_ERR_INVALID, "%s is truncated or its particle count (%lld) does not match its length",
                    path, (long long)hd.n_particles);
  if (fseek(f, (long)sizeof hd, SEEK_SET) != 0) return fsb_fail(c, FSB_ERR_INVALID, "%s: cannot seek", path);
  const int64_t np = hd.n_particles;
  std::vector<uint8_t> lab(cells);
  std::vector<float> grids(8 * cells), parts((size_t)np * 4);
  std::vector<int> orig((size_t)np);
  bool ok = fread(lab.data(), 1, cells, f) == cells &&
            fread(grids.data(), sizeof(float), 8 * cells, f) == 8 * cells &&
            (np == 0 || (fread(parts.data(), 4 * sizeof(float), (size_t)np, f) == (size_t)np &&
                         fread(orig.data(), sizeof(int), (size_t)np, f) == (size_t)np));
  guard.close();
  if (!ok) return fsb_fail(c, FSB_ERR_INVALID, "%s is truncated", path);
  for (size_t k = 0; k < cells; ++k)
    if (lab[k] > FSB_SOLID) return fsb_fail(c, FSB_ERR_INVALID, "%s: cell label out of range", path);
  {
    // the map must be a permutation of 0 .. n-1 (it indexes the caller-order buffer on read-back)
    std::vector<uint8_t> seen((size_t)np, 0);
    for (int64_t k = 0; k < np; ++k)
    {
      if (orig[(size_t)k] < 0 || orig[(size_t)k] >= np || seen[(size_t)orig[(size_t)k]])
        return fsb_fail(c, FSB_ERR_INVALID, "%s: the particle index map is not a permutation", path);
      seen[(size_t)orig[(size_t)k]] = 1;
    }
  }
  // ---- everything checked: now the context changes
  if (np > 0) FSB_TRY(ensure_particle_capacity(c, np));
  FSB_TRY(fsb_set_cell_types(c, lab.data()));
  c->diff_pending = false;
  for (int w = FSB_U_FRONT; w <= FSB_V_DIFF; ++w) FSB_TRY(fsb_set_grid(c, w, grids.data() + (size_t)w * cells));
  c->n = 0;
  c->sort_valid = false;
  if (np > 0)
  {
    FSB_CUDA(c, cudaMemcpyAsync(c->part[c->pcur], parts.data(), sizeof(float4) * np, cudaMemcpyHostToDevice,
                                c->stream));
    FSB_CUDA(c, cudaMemcpyAsync(c->orig[c->pcur], orig.data(), sizeof(int) * np, cudaMemcpyHostToDevice,
                                c->stream));
  }
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n = np;
  c->dx = hd.dx; c->dy = hd.dy; c->density = hd.density; c->pic_ratio = hd.pic_ratio;
  c->grav_x = hd.grav_x; c->grav_y = hd.grav_y; c->integrator = hd.integrator;
  c->max_iters = hd.max_iters; c->tol = hd.tol;
  c->pool_dx = hd.pool_dx; c->pool_dy = hd.pool_dy; c->pool_nx = hd.pool_nx; c->pool_ny = hd.pool_ny;
  c->pressure_valid = false;
  return FSB_OK;
}

int fsb_load_state(fsb_ctx* c, const char* path)
{
  CHECK_CTX(c);
  if (!path) return fsb_fail(c, FSB_ERR_INVALID, "null path");
  if (c->slab_partitioned) REFUSE_ON_SLAB(c, "fsb_load_state");
  try
  {
    return load_state_impl(c, path);
  }
  catch (const std::bad_alloc&)
  {
    return fsb_fail(c, FSB_ERR_NOMEM, "out of host memory while reading %s", path);
  }
  catch (const std::exception& e)
  {
    return fsb_fail(c, FSB_ERR_INVALID, "reading %s failed: %s", path, e.what());
  }
}

// ------------------------------------------------------------------ sharding
namespace {
struct ShardBlob
{
  uint32_t magic;
  int nx, ny, ld;
  cudaIpcMemHandle_t h[6]; // r, p[0], p[1], r2, x, mailbox
};
static_assert(sizeof(ShardBlob) <= FSB_SHARD_BLOB_BYTES, "blob too large");
constexpr uint32_t kBlobMagic = 0x46534232u; // "FSB2"
} // namespace

int fsb_shard_export(fsb_ctx* c, void* blob)
{
  CHECK_CTX(c);
  if (!blob) return fsb_fail(c, FSB_ERR_INVALID, "null blob");
  if (!c->mail_local)
  {
    FSB_TRY(dev_alloc(c, &c->mail_local, (size_t)kMailTypes * kMaxRanks));
    FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  ShardBlob b;
  memset(&b, 0, sizeof b);
  b.magic = kBlobMagic;
  b.nx = c->nx; b.ny = c->ny; b.ld = c->ld;
  void* ptrs[6] = {c->cg_r, c->cg_p[0], c->cg_p[1], c->cg_r2, c->cg_x, c->mail_local};
  for (int k = 0; k < 6; ++k) FSB_CUDA(c, cudaIpcGetMemHandle(&b.h[k], ptrs[k]));
  memset(blob, 0, FSB_SHARD_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  return FSB_OK;
}

int fsb_shard_disconnect(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < c->n_ipc_opened; ++k) cudaIpcCloseMemHandle(c->ipc_opened[k]);
  c->n_ipc_opened = 0;
  if (c->peer_x_dev) cudaFree(c->peer_x_dev);
  c->peer_x_dev = nullptr;
  for (int q = 0; q < kMaxRanks; ++q)
  {
    c->peer_r[q] = c->peer_r2[q] = c->peer_p[0][q] = c->peer_p[1][q] = c->peer_x[q] = nullptr;
    c->shard.mail[q] = nullptr;
  }
  c->shard.world = 1;
  c->shard.rank = 0;
  fsb_cg_reconfigure(c);
  return FSB_OK;
}

int fsb_shard_connect(fsb_ctx* c, int rank, int world, const void* all_blobs)
{
  CHECK_CTX(c);
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !all_blobs)
    return fsb_fail(c, FSB_ERR_INVALID, "bad shard arguments (rank %d of %d)", rank, world);
  if (!c->mail_local) return fsb_fail(c, FSB_ERR_INVALID, "call fsb_shard_export first");
  if (c->ny < 4 * world) return fsb_fail(c, FSB_ERR_INVALID, "grid too small for %d slabs", world);
  FSB_TRY(fsb_shard_disconnect(c));
  if (world == 1) return FSB_OK;
  const char* base = (const char*)all_blobs;
  float* peers_x[kMaxRanks];
  int n_peers = 0;
  for (int q = 0; q < world; ++q)
  {
    if (q == rank)
    {
      c->shard.mail[q] = c->mail_local;
      continue;
    }
    ShardBlob b;
    memcpy(&b, base + (size_t)q * FSB_SHARD_BLOB_BYTES, sizeof b);
    if (b.magic != kBlobMagic || b.nx != c->nx || b.ny != c->ny || b.ld != c->ld)
      return fsb_fail(c, FSB_ERR_INVALID, "rank %d exported a different domain", q);
    void* opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const bool neighbour = (q == rank - 1 || q == rank + 1);
    for (int k = 0; k < 6; ++k)
    {
      if (k < 4 && !neighbour) continue; // r and p are only needed from the adjacent slabs
      cudaError_t e = cudaIpcOpenMemHandle(&opened[k], b.h[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fsb_fail(c, FSB_ERR_COMM, "cudaIpcOpenMemHandle(rank %d, buffer %d): %s", q, k,
                        cudaGetErrorString(e));
      c->ipc_opened[c->n_ipc_opened++] = opened[k];
    }
    c->peer_r[q] = (float*)opened[0];
    c->peer_p[0][q] = (float*)opened[1];
    c->peer_p[1][q] = (float*)opened[2];
    c->peer_r2[q] = (float*)opened[3];
    c->peer_x[q] = (float*)opened[4];
    c->shard.mail[q] = (MailSlot*)opened[5];
    peers_x[n_peers++] = c->peer_x[q];
  }
  FSB_TRY(dev_alloc(c, &c->peer_x_dev, (size_t)kMaxRanks, 0));
  FSB_CUDA(c, cudaMemcpyAsync(c->peer_x_dev, peers_x, sizeof(float*) * n_peers,
                              cudaMemcpyHostToDevice, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->shard.world = world;
  c->shard.rank = rank;
  // even split of the rows; the slabs need not be multiples of the tile height
  c->shard.row_lo = (int)((int64_t)c->ny * rank / world);
  c->shard.row_hi = (int)((int64_t)c->ny * (rank + 1) / world);
  fsb_cg_reconfigure(c);
  return FSB_OK;
}

int fsb_shard_rows(const fsb_ctx* c, int* row_lo, int* row_hi)
{
  if (!c) return FSB_ERR_INVALID;
  if (row_lo) *row_lo = c->shard.world > 1 ? c->shard.row_lo : 0;
  if (row_hi) *row_hi = c->shard.world > 1 ? c->shard.row_hi : c->ny;
  return FSB_OK;
}

// -------------------------------------------------------------- measurement
int fsb_profile_enable(fsb_ctx* c, int on)
{
  CHECK_CTX(c);
  FSB_TRY(prof_drain(c));
  c->profiling = on != 0;
  return FSB_OK;
}
int fsb_profile_read(fsb_ctx* c, float* ms, int* calls)
{
  CHECK_CTX(c);
  FSB_TRY(prof_drain(c));
  for (int k = 0; k < FSB_PROF_COUNT; ++k)
  {
    if (ms) ms[k] = c->prof_ms[k];
    if (calls) calls[k] = c->prof_calls[k];
    c->prof_ms[k] = 0;
    c->prof_calls[k] = 0;
  }
  return FSB_OK;
}
int64_t fsb_launch_count(const fsb_ctx* c) { return c ? c->launches : 0; }
int64_t fsb_cg_swept_cells(const fsb_ctx* c)
{
  if (!c || c->cg_tile_rows == 0) return 0;
  const int rows = (c->shard.world > 1 ? c->shard.row_hi - c->shard.row_lo : c->ny);
  const int64_t tiles_x = (c->ld + 127) / 128;
  const int64_t n_tiles = tiles_x * ((rows + c->cg_tile_rows - 1) / c->cg_tile_rows);
  const int64_t n = (c->cg_skip_tiles && c->scal_h) ? (int64_t)c->cg_last_active_tiles : n_tiles;
  return (n > 0 ? n : n_tiles) * 128 * c->cg_tile_rows;
}

int fsb_cg_launch_mode(const fsb_ctx* c)
{
  if (!c) return 0;
  if (c->last_solve_mg) return 3;
  if (c->cg_tile_rows == 0) return 0;
  if (c->last_solve_one) return 4;
  return c->cg_fused ? 2 : 1;
}
int fsb_timer_start(fsb_ctx* c)
{
  CHECK_CTX(c);
  FSB_CUDA(c, cudaEventRecord(c->timer_ev[0], c->stream));
  return FSB_OK;
}
int fsb_timer_stop(fsb_ctx* c, float* elapsed_ms)
{
  CHECK_CTX(c);
  FSB_CUDA(c, cudaEventRecord(c->timer_ev[1], c->stream));
  FSB_CUDA(c, cudaEventSynchronize(c->timer_ev[1]));
  float ms = 0;
  FSB_CUDA(c, cudaEventElapsedTime(&ms, c->timer_ev[0], c->timer_ev[1]));
  if (elapsed_ms) *elapsed_ms = ms;
  return FSB_OK;
}

} // extern "C"
