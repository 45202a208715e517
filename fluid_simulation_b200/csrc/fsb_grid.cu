// Grid-side stages of the step: cell classification, previous/diff buffers,
// external acceleration, Dirichlet walls, velocity extension, semi-Lagrangian
// velocity advection.  One thread per cell, rows padded to `ld`.
#include "fsb_device.cuh"
#include "fsb_internal.cuh"
#define FSB_VEC_WANT_GRID
#include "fsb_vec_kernels.cuh"

namespace {

constexpr int kBlock = 256;

// one thread per cell: blockIdx.y is the row, so no integer division is needed (runtime
// div / mod issue on the slow XU pipe)
__device__ __forceinline__ bool cell_of_thread(const GridDims d, int* i, int* j)
{
  *i = blockIdx.x * blockDim.x + threadIdx.x;
  *j = blockIdx.y;
  return *i < d.nx && *j < d.ny;
}

inline dim3 cell_grid(const fsb_ctx* c) { return dim3(fsb_div_up(c->ld, kBlock), c->ny); }
// one thread per float4 group / per 16 labels
inline dim3 vec4_grid(const fsb_ctx* c) { return dim3(fsb_div_up(c->ld, 4 * kBlock), c->ny); }
inline dim3 vec16_grid(const fsb_ctx* c) { return dim3(fsb_div_up(c->ld, 16 * kBlock), c->ny); }
inline GridDims dims(const fsb_ctx* c) { return make_grid_dims(c->nx, c->ny, c->ld, c->dx, c->dy); }

// src/MacGrid.cpp:32-50 clearCellTypeBuffer + the border reset of
// src/FluidDomain.cpp:169-179 (the border is SOLID before and after marking).
__global__ void k_fill_labels(uint8_t* __restrict__ cell, const GridDims d)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const bool border = (i == 0 || j == 0 || i == d.nx - 1 || j == d.ny - 1);
  cell[i + (size_t)j * d.ld] = border ? FSB_SOLID : FSB_AIR;
}

// src/FluidDomain.cpp:157-167: cell = (int)((pos / length) * size), clamped.
// lengthX() is recomputed as size * delta in float (include/Grid.h:54-55).
// Marking a border cell is undone by the border reset, so it is skipped; the
// store is idempotent, hence race-free without atomics.
// `len` carries lengthX / lengthY in its dx / dy slots (with the power-of-two fast path).
__global__ void k_mark_liquid(const float4* __restrict__ part, int64_t n,
                              uint8_t* __restrict__ cell, const GridDims d, const GridDims len)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float4 p = part[k];
  int x = (int)(div_dx(len, p.x) * (float)d.nx);
  int y = (int)(div_dy(len, p.y) * (float)d.ny);
  x = clampi(x, 0, d.nx - 1);
  y = clampi(y, 0, d.ny - 1);
  if (x == 0 || y == 0 || x == d.nx - 1 || y == d.ny - 1) return;
  cell[x + (size_t)y * d.ld] = FSB_LIQUID;
}

// src/MacGrid.cpp:58-70
__global__ void k_update_diff(const float* __restrict__ uf, const float* __restrict__ vf,
                              const float* __restrict__ up, const float* __restrict__ vp,
                              float* __restrict__ ud, float* __restrict__ vd, const GridDims d)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const size_t k = i + (size_t)j * d.ld;
  ud[k] = uf[k] - up[k];
  vd[k] = vf[k] - vp[k];
}

// src/FluidSolver.cpp:276-295: only the left and bottom faces of LIQUID cells
__global__ void k_add_acceleration(float* __restrict__ uf, float* __restrict__ vf,
                                   const uint8_t* __restrict__ cell, const GridDims d, float ax,
                                   float ay, float dt)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const size_t k = i + (size_t)j * d.ld;
  if (cell[k] == FSB_LIQUID)
  {
    uf[k] = uf[k] + ax * dt;
    vf[k] = vf[k] + ay * dt;
  }
}

// src/FluidSolver.cpp:297-321: one-sided wall condition
__global__ void k_enforce_dirichlet(float* __restrict__ uf, float* __restrict__ vf,
                                    const uint8_t* __restrict__ cell, const GridDims d)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const size_t k = i + (size_t)j * d.ld;
  const int im1 = clampi(i - 1, 0, d.nx - 1);
  const int jm1 = clampi(j - 1, 0, d.ny - 1);
  const int c = cell[k];
  const int cw = cell[im1 + (size_t)j * d.ld];
  const int cs = cell[i + (size_t)jm1 * d.ld];
  const float u = uf[k], v = vf[k];
  if ((cw == FSB_SOLID && u < 0.0f) || (c == FSB_SOLID && u > 0.0f)) uf[k] = 0.0f;
  if ((cs == FSB_SOLID && v < 0.0f) || (c == FSB_SOLID && v > 0.0f)) vf[k] = 0.0f;
}

// Fused form of the three passes that follow P2G in stepFLIP / stepPICFLIP
// (src/FluidSolver.cpp:190-193,230-233): previous = front (src/MacGrid.cpp:52-56), gravity on the
// left / bottom faces of LIQUID cells (:276-295), one-sided Dirichlet (:297-321).  Four cells per
// thread (16-byte accesses); 9 B read + 16 B written per cell instead of three passes.  Each
// face's arithmetic is the reference's, so the result is bit-identical to the separate stages.
__global__ void k_prev_gravity_dirichlet(float* __restrict__ uf, float* __restrict__ vf,
                                         float* __restrict__ up, float* __restrict__ vp,
                                         const uint8_t* __restrict__ cell, const GridDims d,
                                         float ax, float ay, float dt, int save_prev)
{
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i >= d.ld) return;
  const size_t k = i + (size_t)j * d.ld;
  float4 u = *reinterpret_cast<const float4*>(uf + k);
  float4 v = *reinterpret_cast<const float4*>(vf + k);
  const uchar4 c = *reinterpret_cast<const uchar4*>(cell + k);
  const uchar4 s = *reinterpret_cast<const uchar4*>(cell + i + (size_t)max(j - 1, 0) * d.ld);
  const uint8_t w = cell[max(i - 1, 0) + (size_t)j * d.ld];
  if (save_prev)
  {
    *reinterpret_cast<float4*>(up + k) = u;
    *reinterpret_cast<float4*>(vp + k) = v;
  }
  const float gu = ax * dt, gv = ay * dt;
  if (c.x == FSB_LIQUID) { u.x = u.x + gu; v.x = v.x + gv; }
  if (c.y == FSB_LIQUID) { u.y = u.y + gu; v.y = v.y + gv; }
  if (c.z == FSB_LIQUID) { u.z = u.z + gu; v.z = v.z + gv; }
  if (c.w == FSB_LIQUID) { u.w = u.w + gu; v.w = v.w + gv; }
#define FSB_WALL(val, lower, here) \
  if (((lower) == FSB_SOLID && (val) < 0.0f) || ((here) == FSB_SOLID && (val) > 0.0f)) (val) = 0.0f
  FSB_WALL(u.x, w, c.x);   FSB_WALL(u.y, c.x, c.y); FSB_WALL(u.z, c.y, c.z); FSB_WALL(u.w, c.z, c.w);
  FSB_WALL(v.x, s.x, c.x); FSB_WALL(v.y, s.y, c.y); FSB_WALL(v.z, s.z, c.z); FSB_WALL(v.w, s.w, c.w);
#undef FSB_WALL
  *reinterpret_cast<float4*>(uf + k) = u;
  *reinterpret_cast<float4*>(vf + k) = v;
}

// src/FluidSolver.cpp:490-530: validity masks and the front->back copy,
// including the line-527 typo (an invalid v-face zeroes the front U).
__global__ void k_extend_init(float* __restrict__ uf, const float* __restrict__ vf,
                              float* __restrict__ ub, float* __restrict__ vb,
                              uint8_t* __restrict__ mxf, uint8_t* __restrict__ mxb,
                              uint8_t* __restrict__ myf, uint8_t* __restrict__ myb,
                              const uint8_t* __restrict__ cell, const GridDims d)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const size_t k = i + (size_t)j * d.ld;
  const bool liq = cell[k] == FSB_LIQUID;
  const bool u_valid = liq || cell_type(cell, d, i - 1, j) == FSB_LIQUID;
  const bool v_valid = liq || cell_type(cell, d, i, j - 1) == FSB_LIQUID;
  const float u = uf[k];
  mxf[k] = u_valid;
  mxb[k] = u_valid;
  ub[k] = u_valid ? u : 0.0f;
  myf[k] = v_valid;
  myb[k] = v_valid;
  vb[k] = v_valid ? vf[k] : 0.0f;
  if (!u_valid || !v_valid) uf[k] = 0.0f;
}

// src/FluidSolver.cpp:533-620, one sweep.  Reads only faces whose FRONT mask
// is 1 and writes only faces whose front mask is 0, so updating the back
// velocity buffer in place is race-free and order-independent; the four-term
// sum keeps the reference's order (i-1,j),(i,j-1),(i,j+1),(i+1,j).
__global__ void k_extend_sweep(float* __restrict__ ub, float* __restrict__ vb,
                               const uint8_t* __restrict__ mxf, uint8_t* __restrict__ mxb,
                               const uint8_t* __restrict__ myf, uint8_t* __restrict__ myb,
                               const uint8_t* __restrict__ cell, const GridDims d)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  // a face that passes the tests below has a non-SOLID cell on both sides, so
  // 1 <= i,j <= size-2 there and the four neighbours exist
  if (i < 1 || j < 1 || i > d.nx - 2 || j > d.ny - 2) return;
  const size_t k = i + (size_t)j * d.ld;
  const size_t kw = k - 1, ke = k + 1, ks = k - d.ld, kn = k + d.ld;
  const bool not_solid = cell[k] != FSB_SOLID;
  if (mxf[k] == 0 && not_solid && cell[kw] != FSB_SOLID)
  {
    float nv = 0.0f;
    int n = 0;
    if (mxf[kw] == 1) { nv += ub[kw]; n++; }
    if (mxf[ks] == 1) { nv += ub[ks]; n++; }
    if (mxf[kn] == 1) { nv += ub[kn]; n++; }
    if (mxf[ke] == 1) { nv += ub[ke]; n++; }
    if (n > 0)
    {
      ub[k] = nv / (float)n;
      mxb[k] = 1;
    }
  }
  if (myf[k] == 0 && not_solid && cell[ks] != FSB_SOLID)
  {
    float nv = 0.0f;
    int n = 0;
    if (myf[kw] == 1) { nv += vb[kw]; n++; }
    if (myf[ks] == 1) { nv += vb[ks]; n++; }
    if (myf[kn] == 1) { nv += vb[kn]; n++; }
    if (myf[ke] == 1) { nv += vb[ke]; n++; }
    if (n > 0)
    {
      vb[k] = nv / (float)n;
      myb[k] = 1;
    }
  }
}

// include/Grid.h:152-184 as a scatter with float atomics (order of the sums
// differs from the reference's face order: within 1e-5 of the field maximum).
template <class D>
__device__ __forceinline__ void grid_splat_atomic(float* __restrict__ g, const D d, float x, float y,
                                                  float value)
{
  const float xd = div_dx(d, x);
  const float yd = div_dy(d, y);
  int i = (int)xd;
  int j = (int)yd;
  int i1 = i + 1;
  int j1 = j + 1;
  const float fi = xd - (float)i;
  const float fj = yd - (float)j;
  i = clampi(i, 0, d.nx - 1);
  j = clampi(j, 0, d.ny - 1);
  i1 = clampi(i1, 0, d.nx - 1);
  j1 = clampi(j1, 0, d.ny - 1);
  const float v0 = (1.0f - fj) * value;
  const float v1 = fj * value;
  atomicAdd(g + i + (size_t)j * d.ld, (1.0f - fi) * v0);
  atomicAdd(g + i1 + (size_t)j * d.ld, fi * v0);
  atomicAdd(g + i + (size_t)j1 * d.ld, (1.0f - fi) * v1);
  atomicAdd(g + i1 + (size_t)j1 * d.ld, fi * v1);
}

// src/FluidSolver.cpp:721-771: forward splat of each liquid-adjacent face value
// to its back-traced position, into the zeroed BACK buffer; no swap.
template <class D>
__global__ void k_advect_velocity_sl(const float* __restrict__ uf, const float* __restrict__ vf,
                                     float* __restrict__ ub, float* __restrict__ vb,
                                     const uint8_t* __restrict__ cell, const D d, float dt,
                                     int integrator)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const bool liq = cell[i + (size_t)j * d.ld] == FSB_LIQUID;
  if (liq || cell_type(cell, d, i - 1, j) == FSB_LIQUID)
  {
    const float x_pos = (float)i * d.dx;
    const float y_pos = ((float)j + 0.5f) * d.dy;
    float xq, yq;
    advected_position(uf, vf, d, integrator, x_pos, y_pos, -dt, &xq, &yq);
    const float val = vel_x_interp(uf, d, x_pos, y_pos);
    grid_splat_atomic(ub, d, xq, yq - 0.5f * d.dy, val);
  }
  if (liq || cell_type(cell, d, i, j - 1) == FSB_LIQUID)
  {
    const float x_pos = ((float)i + 0.5f) * d.dx;
    const float y_pos = (float)j * d.dy;
    float xq, yq;
    advected_position(uf, vf, d, integrator, x_pos, y_pos, -dt, &xq, &yq);
    const float val = vel_y_interp(vf, d, x_pos, y_pos);
    grid_splat_atomic(vb, d, xq - 0.5f * d.dx, yq, val);
  }
}


} // namespace

static int fill_labels(fsb_ctx* c)
{
  if (c->stage_v1) k_fill_labels<<<cell_grid(c), kBlock, 0, c->stream>>>(c->cell, dims(c));
  else k_fill_labels16<<<vec16_grid(c), kBlock, 0, c->stream>>>(c->cell, dims(c));
  FSB_LAUNCHED(c);
  return FSB_OK;
}

// the border / AIR reset only; the LIQUID marks come from the cell sort's counting pass, which
// reads every particle anyway (fsb_k_sort_particles with mark_labels)
int fsb_k_classify_reset(fsb_ctx* c)
{
  fsb_prof_begin(c, FSB_PROF_CLASSIFY);
  FSB_TRY(fill_labels(c));
  fsb_prof_end(c, FSB_PROF_CLASSIFY);
  return FSB_OK;
}

int fsb_k_classify(fsb_ctx* c)
{
  fsb_prof_begin(c, FSB_PROF_CLASSIFY);
  FSB_TRY(fill_labels(c));
  if (c->n > 0)
  {
    // lengthX() is recomputed as size * delta in float (include/Grid.h:54-55)
    const GridDims len =
        make_grid_dims(c->nx, c->ny, c->ld, (float)c->nx * c->dx, (float)c->ny * c->dy);
    k_mark_liquid<<<fsb_div_up(c->n, kBlock), kBlock, 0, c->stream>>>(c->part[c->pcur], c->n,
                                                                       c->cell, dims(c), len);
    FSB_LAUNCHED(c);
  }
  fsb_prof_end(c, FSB_PROF_CLASSIFY);
  return FSB_OK;
}

int fsb_k_clear_labels(fsb_ctx* c) { return fill_labels(c); }

int fsb_k_save_previous(fsb_ctx* c)
{
  const size_t bytes = (size_t)c->ld * c->ny * sizeof(float);
  fsb_prof_begin(c, FSB_PROF_GRID_PRE);
  FSB_CUDA(c, cudaMemcpyAsync(c->u_prev, fsb_uf(c), bytes, cudaMemcpyDeviceToDevice, c->stream));
  FSB_CUDA(c, cudaMemcpyAsync(c->v_prev, fsb_vf(c), bytes, cudaMemcpyDeviceToDevice, c->stream));
  fsb_prof_end(c, FSB_PROF_GRID_PRE);
  return FSB_OK;
}

int fsb_k_update_diff(fsb_ctx* c)
{
  k_update_diff<<<cell_grid(c), kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->u_prev, c->v_prev,
                                                        c->u_diff, c->v_diff, dims(c));
  FSB_LAUNCHED(c);
  return FSB_OK;
}

int fsb_k_add_acceleration(fsb_ctx* c, float ax, float ay, float dt)
{
  fsb_prof_begin(c, FSB_PROF_GRID_PRE);
  k_add_acceleration<<<cell_grid(c), kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell, dims(c),
                                                             ax, ay, dt);
  FSB_LAUNCHED(c);
  fsb_prof_end(c, FSB_PROF_GRID_PRE);
  return FSB_OK;
}

int fsb_k_prev_gravity_dirichlet(fsb_ctx* c, float ax, float ay, float dt, int save_prev)
{
  fsb_prof_begin(c, FSB_PROF_GRID_PRE);
  k_prev_gravity_dirichlet<<<dim3(fsb_div_up(c->ld, 4 * kBlock), c->ny), kBlock, 0, c->stream>>>(
      fsb_uf(c), fsb_vf(c), c->u_prev, c->v_prev, c->cell, dims(c), ax, ay, dt, save_prev);
  FSB_LAUNCHED(c);
  fsb_prof_end(c, FSB_PROF_GRID_PRE);
  return FSB_OK;
}

int fsb_k_enforce_dirichlet(fsb_ctx* c)
{
  fsb_prof_begin(c, FSB_PROF_GRID_PRE);
  if (c->stage_v1)
    k_enforce_dirichlet<<<cell_grid(c), kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell,
                                                                dims(c));
  else
    k_enforce_dirichlet4<<<vec4_grid(c), kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell,
                                                                 dims(c));
  FSB_LAUNCHED(c);
  fsb_prof_end(c, FSB_PROF_GRID_PRE);
  return FSB_OK;
}

int fsb_k_extend_velocity(fsb_ctx* c, int n_iter)
{
  fsb_prof_begin(c, FSB_PROF_EXTEND);
  if (n_iter == 2 && !c->stage_v1)
  {
    // two passes, masks recomputed from the labels (see k_extend2_a); mask_x[0] holds the packed
    // validity byte between the passes.  The stored mask buffers are scratch of this stage (every
    // call rebuilds them), so leaving them untouched is not observable.
    uint8_t* m1 = c->mask_x[0];
    const dim3 grid(fsb_div_up(c->ld, 4 * kBlock), fsb_div_up(c->ny, kExtendRows));
    k_extend2_a<<<grid, kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), m1,
                                                c->cell, dims(c));
    FSB_LAUNCHED(c);
    k_extend2_b<<<grid, kBlock, 0, c->stream>>>(fsb_uf(c), fsb_ub(c), fsb_vb(c), m1, c->cell,
                                                dims(c));
    FSB_LAUNCHED(c);
    c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:621
    fsb_prof_end(c, FSB_PROF_EXTEND);
    return FSB_OK;
  }
  k_extend_init<<<cell_grid(c), kBlock, 0, c->stream>>>(
      fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), c->mask_x[c->mask_front],
      c->mask_x[c->mask_front ^ 1], c->mask_y[c->mask_front], c->mask_y[c->mask_front ^ 1], c->cell,
      dims(c));
  FSB_LAUNCHED(c);
  for (int it = 0; it < n_iter; ++it)
  {
    k_extend_sweep<<<cell_grid(c), kBlock, 0, c->stream>>>(
        fsb_ub(c), fsb_vb(c), c->mask_x[c->mask_front], c->mask_x[c->mask_front ^ 1],
        c->mask_y[c->mask_front], c->mask_y[c->mask_front ^ 1], c->cell, dims(c));
    FSB_LAUNCHED(c);
    c->mask_front ^= 1; // swapValidMaskBuffer, src/FluidSolver.cpp:619
  }
  c->front ^= 1; // swapVelocityBuffers, src/FluidSolver.cpp:621
  fsb_prof_end(c, FSB_PROF_EXTEND);
  return FSB_OK;
}

// ----------------------------------------------- extendVelocityAvarageing --
// src/FluidSolver.cpp:625-707 (no step calls it).  A non-SOLID cell without the validity mark takes the
// mean of the cell-centred back-buffer velocities of its marked neighbours and writes it to BOTH of its
// faces per component (include/MacGrid.h:124-133) -- faces it shares with its neighbours -- so within a
// sweep a cell sees what the cells scanned before it wrote: the result depends on the row-major scan
// order.  The dependences are local, though: cell (i, j) reads faces written by (i-2..i-1, j),
// (i-1..i+1, j-1) and (i, j-2) among the cells scanned before it, and cells scanned after it that write
// a face it reads are (i+1..i+2, j), (i-1..i+1, j+1), (i, j+2).  With t = i + 2 j every such earlier
// cell has a smaller t and every later one a larger t, so the cells of one t are mutually independent:
// the sweep runs as nx + 2 ny - 2 wavefronts of one CTA, bit-identical to the sequential scan.
namespace {

__global__ void k_extend_avg_init(const float* __restrict__ uf, const float* __restrict__ vf,
                                  float* __restrict__ ub, float* __restrict__ vb,
                                  const uint8_t* __restrict__ cell, uint8_t* __restrict__ m0,
                                  uint8_t* __restrict__ m1, const GridDims d, int* __restrict__ bad_border)
{
  int i, j;
  if (!cell_of_thread(d, &i, &j)) return;
  const size_t k = i + (size_t)j * d.ld;
  const uint8_t ct = cell[k];
  const uint8_t m = ct == FSB_LIQUID ? 1 : 0;
  m0[k] = m;
  m1[k] = m;
  ub[k] = uf[k];
  vb[k] = vf[k];
  // the reference indexes (i +- 1, j +- 1) and (i + 2, j) / (i, j + 2) of non-SOLID cells without a
  // clamp (it asserts): a classified grid has a SOLID border, anything else is refused
  if ((i == 0 || j == 0 || i == d.nx - 1 || j == d.ny - 1) && ct != FSB_SOLID) *bad_border = 1;
}

__global__ void __launch_bounds__(1024)
k_extend_avg_sweep(float* __restrict__ ub, float* __restrict__ vb, const uint8_t* __restrict__ cell,
                   const uint8_t* __restrict__ mf, uint8_t* __restrict__ mb, const GridDims d)
{
  const int nx = d.nx, ny = d.ny, ld = d.ld;
  auto ubc = [&](int i, int j) { return (ub[i + (size_t)j * ld] + ub[i + 1 + (size_t)j * ld]) / 2.0f; };
  auto vbc = [&](int i, int j) { return (vb[i + (size_t)j * ld] + vb[i + (size_t)(j + 1) * ld]) / 2.0f; };
  for (int t = 0; t <= nx - 1 + 2 * (ny - 1); ++t)
  {
    const int j_lo = max(0, (t - (nx - 1) + 1) / 2), j_hi = min(ny - 1, t / 2);
    for (int j = j_lo + (int)threadIdx.x; j <= j_hi; j += (int)blockDim.x)
    {
      const int i = t - 2 * j;
      const size_t k = i + (size_t)j * ld;
      if (mf[k] == 0 && cell[k] != FSB_SOLID)
      {
        float nvx = 0.0f, nvy = 0.0f;
        int n = 0;
        if (mf[k - 1] == 1) { nvx += ubc(i - 1, j); nvy += vbc(i - 1, j); n++; }
        if (mf[k - ld] == 1) { nvx += ubc(i, j - 1); nvy += vbc(i, j - 1); n++; }
        if (mf[k + ld] == 1) { nvx += ubc(i, j + 1); nvy += vbc(i, j + 1); n++; }
        if (mf[k + 1] == 1) { nvx += ubc(i + 1, j); nvy += vbc(i + 1, j); n++; }
        if (n > 0)
        {
          nvx /= (float)n;
          nvy /= (float)n;
          ub[k] = nvx; ub[k + 1] = nvx;
          vb[k] = nvy; vb[k + ld] = nvy;
          mb[k] = 1;
        }
      }
    }
    __syncthreads();
  }
}

} // namespace

int fsb_k_extend_velocity_avg(fsb_ctx* c, int n_iter)
{
  const GridDims d = dims(c);
  int* flag = nullptr;
  FSB_CUDA(c, cudaMalloc(&flag, sizeof(int)));
  FSB_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  k_extend_avg_init<<<cell_grid(c), kBlock, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c),
                                                              c->cell, c->mask_x[c->mask_front],
                                                              c->mask_x[c->mask_front ^ 1], d, flag);
  FSB_LAUNCHED(c);
  int bad = 0;
  FSB_CUDA(c, cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(flag);
  if (bad)
    return fsb_fail(c, FSB_ERR_INVALID, "extendVelocityAvarageing needs SOLID border cells (the reference "
                                        "indexes past the border of a non-SOLID border cell)");
  for (int it = 0; it < n_iter; ++it)
  {
    k_extend_avg_sweep<<<1, 1024, 0, c->stream>>>(fsb_ub(c), fsb_vb(c), c->cell, c->mask_x[c->mask_front],
                                                   c->mask_x[c->mask_front ^ 1], d);
    FSB_LAUNCHED(c);
    c->mask_front ^= 1; // swapValidMaskBuffer, :703
  }
  c->front ^= 1; // swapVelocityBuffers, :706
  return FSB_OK;
}

int fsb_k_advect_velocity_sl(fsb_ctx* c, float dt)
{
  const size_t bytes = (size_t)c->ld * c->ny * sizeof(float);
  fsb_prof_begin(c, FSB_PROF_ADVECT_SL);
  if (!c->stage_v1 && !c->sl_atomic)
  {
    // the deterministic gather (fsb_sl.cu): bit-identical to the reference's face order
    int done = 0;
    FSB_TRY(fsb_k_advect_velocity_sl_gather(c, dt, &done));
    if (done)
    {
      fsb_prof_end(c, FSB_PROF_ADVECT_SL);
      return FSB_OK;
    }
  }
  FSB_CUDA(c, cudaMemsetAsync(fsb_ub(c), 0, bytes, c->stream));
  FSB_CUDA(c, cudaMemsetAsync(fsb_vb(c), 0, bytes, c->stream));
  const GridDims d = dims(c);
  if (d.pow2 == 3)
    k_advect_velocity_sl<GridDimsP2><<<cell_grid(c), kBlock, 0, c->stream>>>(
        fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), c->cell, as_pow2(d), dt, c->integrator);
  else
    k_advect_velocity_sl<GridDims><<<cell_grid(c), kBlock, 0, c->stream>>>(
        fsb_uf(c), fsb_vf(c), fsb_ub(c), fsb_vb(c), c->cell, d, dt, c->integrator);
  FSB_LAUNCHED(c);
  fsb_prof_end(c, FSB_PROF_ADVECT_SL);
  return FSB_OK;
}
