// Semi-Lagrangian velocity advection, src/FluidSolver.cpp:709-772, as a
// deterministic gather.
//
// The reference walks the faces row-major; every face next to a LIQUID cell is traced back with
// the configured integrator (RK3: include/OdeSolver.h:102-113), and its own interpolated value is
// SPLAT bilinearly at the back-traced position into the zeroed BACK buffer (include/Grid.h:152-184);
// no normalisation, no swap.  A node of the back buffer is therefore a sum of contributions in the
// order in which the reference visits their source faces.
//
// Two kernels:
//   k_sl_trace   one thread per cell: both faces' back-traced positions with the three RK stages
//                fused (bilinear taps straight from the front buffers), reduced to a 16-byte record
//                per face {fi, fj, value, packed (di, dj, valid)} -- (di, dj) is the base node of the
//                splat RELATIVE to the source face -- and the largest |di|, |dj| of the launch.
//   k_sl_gather  one thread per destination node, tiles of 32 x 8 nodes: the records of the tile and
//                of a halo as wide as the largest displacement are staged in shared memory; every
//                node runs through its candidate sources in the reference's row-major source order and
//                adds, with the reference's expression order, the parts of their splats that land on
//                it (clamped duplicates included).  Same additions in the same order as the reference:
//                bit-identical to the CPU path, independent of any scheduling, hence reproducible
//                across ranks (a slab needs a halo of `reach` source rows).
// The previous form (one thread per face, float atomics into the back buffer) is kept as the fallback
// for displacements beyond 127 cells per step, which no stable simulation produces.
#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"

namespace {

constexpr int kTW = 32, kTH = 8; // destination tile of the gather
constexpr int kMaxPack = 127;    // |di|, |dj| representable in a record

__device__ __forceinline__ uint32_t pack_rec(int di, int dj)
{
  return 0x80000000u | (uint32_t)((di + 128) & 0xff) | ((uint32_t)((dj + 128) & 0xff) << 8);
}

// back-traced splat of one face: base node (unclamped, include/Grid.h:152-160) and fractions
template <class D>
__device__ __forceinline__ float4 trace_face(const float* __restrict__ uf, const float* __restrict__ vf,
                                             const D d, int integrator, float x_pos, float y_pos,
                                             float dt, bool is_u, int is, int js, int* maxd)
{
  float xq, yq;
  advected_position(uf, vf, d, integrator, x_pos, y_pos, -dt, &xq, &yq);
  const float val = is_u ? vel_x_interp(uf, d, x_pos, y_pos) : vel_y_interp(vf, d, x_pos, y_pos);
  // addToVelXInterpolated(x, y) = U.splat(x, y - dy/2); addToVelYInterpolated = V.splat(x - dx/2, y)
  const float sx = is_u ? xq : xq - 0.5f * d.dx;
  const float sy = is_u ? yq - 0.5f * d.dy : yq;
  const float xd = div_dx(d, sx);
  const float yd = div_dy(d, sy);
  const int i = (int)xd, j = (int)yd;
  const float fi = xd - (float)i, fj = yd - (float)j;
  int di = i - is, dj = j - js;
  // non-finite positions or absurd displacements: flagged, the launcher falls back
  if (!(di >= -kMaxPack && di <= kMaxPack && dj >= -kMaxPack && dj <= kMaxPack))
  {
    *maxd = 1 << 20;
    di = dj = 0;
  }
  *maxd = max(*maxd, max(abs(di), abs(dj)));
  return make_float4(fi, fj, val, __uint_as_float(pack_rec(di, dj)));
}

template <class D>
__global__ void __launch_bounds__(256)
k_sl_trace(const float* __restrict__ uf, const float* __restrict__ vf, const uint8_t* __restrict__ cell,
           float4* __restrict__ rec_u, float4* __restrict__ rec_v, const D d, float dt, int integrator,
           int* __restrict__ maxd_out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  int maxd = 0;
  if (i < d.nx)
  {
    const size_t t = i + (size_t)j * d.nx; // records are dense (nx per row)
    const bool liq = cell[i + (size_t)j * d.ld] == FSB_LIQUID;
    float4 ru = make_float4(0.f, 0.f, 0.f, 0.f), rv = ru;
    if (liq || cell_type(cell, d, i - 1, j) == FSB_LIQUID) // :723-747
      ru = trace_face(uf, vf, d, integrator, (float)i * d.dx, ((float)j + 0.5f) * d.dy, dt, true, i, j, &maxd);
    if (liq || cell_type(cell, d, i, j - 1) == FSB_LIQUID) // :749-768
      rv = trace_face(uf, vf, d, integrator, ((float)i + 0.5f) * d.dx, (float)j * d.dy, dt, false, i, j, &maxd);
    rec_u[t] = ru;
    rec_v[t] = rv;
  }
  // one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
  if ((threadIdx.x & 31) == 0 && maxd > 0) atomicMax(maxd_out, maxd);
}

// what one source record adds to node (a, b): include/Grid.h:166-183, write order
// (i, j), (i1, j), (i, j1), (i1, j1), every index clamped on its own
__device__ __forceinline__ void add_source(float& acc, const float4 r, int sx, int sy, int a, int b, int nx,
                                           int ny)
{
  const uint32_t w = __float_as_uint(r.w);
  if (!(w & 0x80000000u)) return;
  const int i = sx + (int)(w & 0xff) - 128, j = sy + (int)((w >> 8) & 0xff) - 128;
  const int i0 = clampi(i, 0, nx - 1), i1 = clampi(i + 1, 0, nx - 1);
  if (i0 != a && i1 != a) return;
  const int j0 = clampi(j, 0, ny - 1), j1 = clampi(j + 1, 0, ny - 1);
  if (j0 != b && j1 != b) return;
  const float fi = r.x, fj = r.y, val = r.z;
  const float v0 = (1.0f - fj) * val;
  const float v1 = fj * val;
  if (i0 == a && j0 == b) acc = acc + (1.0f - fi) * v0;
  if (i1 == a && j0 == b) acc = acc + fi * v0;
  if (i0 == a && j1 == b) acc = acc + (1.0f - fi) * v1;
  if (i1 == a && j1 == b) acc = acc + fi * v1;
}

// W: reach of the candidate window in source cells (largest |di|, |dj| + 1: a splat also touches
// base + 1).  Component z = 0: u, 1: v.
template <int W>
__global__ void __launch_bounds__(kTW* kTH)
k_sl_gather(const float4* __restrict__ rec_u, const float4* __restrict__ rec_v, float* __restrict__ ub,
            float* __restrict__ vb, int nx, int ny, int ld)
{
  constexpr int SW = kTW + 2 * W, SH = kTH + 2 * W;
  __shared__ float4 s_rec[SH][SW];
  const float4* __restrict__ rec = blockIdx.z == 0 ? rec_u : rec_v;
  float* __restrict__ out = blockIdx.z == 0 ? ub : vb;
  const int a0 = blockIdx.x * kTW, b0 = blockIdx.y * kTH;
  int any_valid = 0;
  for (int t = threadIdx.x; t < SW * SH; t += kTW * kTH)
  {
    const int lx = t % SW, ly = t / SW;
    const int sx = a0 - W + lx, sy = b0 - W + ly;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sx >= 0 && sx < nx && sy >= 0 && sy < ny) r = rec[sx + (size_t)sy * nx];
    s_rec[ly][lx] = r;
    any_valid |= (__float_as_uint(r.w) & 0x80000000u) != 0u;
  }
  // a window without a single traced face (air, solid): the nodes keep the zero of the cleared buffer
  any_valid = __syncthreads_or(any_valid);
  const int tx = threadIdx.x % kTW, ty = threadIdx.x / kTW;
  const int a = a0 + tx, b = b0 + ty;
  if (a >= nx || b >= ny) return;
  if (!any_valid)
  {
    out[a + (size_t)b * ld] = 0.0f;
    return;
  }
  float acc = 0.0f; // the zeroed back buffer, :712-719
  // the reference's source order: rows ascending, columns ascending within a row
#pragma unroll 1
  for (int ly = 0; ly <= 2 * W; ++ly)
  {
#pragma unroll
    for (int lx = 0; lx <= 2 * W; ++lx)
      add_source(acc, s_rec[ty + ly][tx + lx], a - W + lx, b - W + ly, a, b, nx, ny);
  }
  out[a + (size_t)b * ld] = acc;
}

// any reach: candidates straight from global memory (L1 / L2 serve the overlap)
__global__ void __launch_bounds__(256)
k_sl_gather_any(const float4* __restrict__ rec_u, const float4* __restrict__ rec_v, float* __restrict__ ub,
                float* __restrict__ vb, int nx, int ny, int ld, int w)
{
  const float4* __restrict__ rec = blockIdx.z == 0 ? rec_u : rec_v;
  float* __restrict__ out = blockIdx.z == 0 ? ub : vb;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= nx) return;
  float acc = 0.0f;
  for (int sy = max(0, b - w); sy <= min(ny - 1, b + w); ++sy)
    for (int sx = max(0, a - w); sx <= min(nx - 1, a + w); ++sx)
      add_source(acc, __ldg(rec + sx + (size_t)sy * nx), sx, sy, a, b, nx, ny);
  out[a + (size_t)b * ld] = acc;
}

} // namespace

// Returns FSB_OK with *done = 0 when the displacement does not fit a record: the caller falls back.
int fsb_k_advect_velocity_sl_gather(fsb_ctx* c, float dt, int* done)
{
  *done = 0;
  const size_t cells = (size_t)c->nx * c->ny;
  if (c->sl_cells != cells)
  {
    if (c->sl_rec_u) cudaFree(c->sl_rec_u);
    if (c->sl_rec_v) cudaFree(c->sl_rec_v);
    c->sl_rec_u = c->sl_rec_v = nullptr;
    c->sl_cells = 0;
    if (cudaMalloc(&c->sl_rec_u, sizeof(float4) * cells) != cudaSuccess ||
        cudaMalloc(&c->sl_rec_v, sizeof(float4) * cells) != cudaSuccess)
    {
      cudaGetLastError();
      if (c->sl_rec_u) cudaFree(c->sl_rec_u);
      c->sl_rec_u = c->sl_rec_v = nullptr;
      return FSB_OK; // no room for the records: the fallback needs none
    }
    if (!c->sl_maxd && cudaMalloc(&c->sl_maxd, sizeof(int)) != cudaSuccess)
      return fsb_fail(c, FSB_ERR_NOMEM, "cudaMalloc failed");
    c->sl_cells = cells;
  }
  const GridDims d = make_grid_dims(c->nx, c->ny, c->ld, c->dx, c->dy);
  FSB_CUDA(c, cudaMemsetAsync(c->sl_maxd, 0, sizeof(int), c->stream));
  const dim3 tgrid(fsb_div_up(c->nx, 256), c->ny);
  if (d.pow2 == 3)
    k_sl_trace<GridDimsP2><<<tgrid, 256, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell, (float4*)c->sl_rec_u,
                                                        (float4*)c->sl_rec_v, as_pow2(d), dt, c->integrator,
                                                        c->sl_maxd);
  else
    k_sl_trace<GridDims><<<tgrid, 256, 0, c->stream>>>(fsb_uf(c), fsb_vf(c), c->cell, (float4*)c->sl_rec_u,
                                                      (float4*)c->sl_rec_v, d, dt, c->integrator, c->sl_maxd);
  FSB_LAUNCHED(c);
  int maxd = 0;
  FSB_CUDA(c, cudaMemcpyAsync(&maxd, c->sl_maxd, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FSB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (maxd > kMaxPack) return FSB_OK;
  const int w = maxd + 1;
  c->sl_reach = w;
  const dim3 ggrid(fsb_div_up(c->nx, kTW), fsb_div_up(c->ny, kTH), 2);
  const float4 *ru = (const float4*)c->sl_rec_u, *rv = (const float4*)c->sl_rec_v;
#define FSB_SL_GATHER(W) \
  k_sl_gather<W><<<ggrid, kTW * kTH, 0, c->stream>>>(ru, rv, fsb_ub(c), fsb_vb(c), c->nx, c->ny, c->ld)
  if (w == 1) FSB_SL_GATHER(1);
  else if (w == 2) FSB_SL_GATHER(2);
  else if (w == 3) FSB_SL_GATHER(3);
  else if (w == 4) FSB_SL_GATHER(4);
  else
    k_sl_gather_any<<<dim3(fsb_div_up(c->nx, 256), c->ny, 2), 256, 0, c->stream>>>(
        ru, rv, fsb_ub(c), fsb_vb(c), c->nx, c->ny, c->ld, w);
#undef FSB_SL_GATHER
  FSB_LAUNCHED(c);
  *done = 1;
  return FSB_OK;
}

void fsb_sl_free(fsb_ctx* c)
{
  if (c->sl_rec_u) cudaFree(c->sl_rec_u);
  if (c->sl_rec_v) cudaFree(c->sl_rec_v);
  if (c->sl_maxd) cudaFree(c->sl_maxd);
  c->sl_rec_u = c->sl_rec_v = nullptr;
  c->sl_maxd = nullptr;
  c->sl_cells = 0;
}
