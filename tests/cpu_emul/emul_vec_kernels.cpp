// TEST INFRASTRUCTURE ONLY.  Host emulation of the vectorised stage kernels of
// fluid_simulation_b200/csrc/fsb_vec_kernels.cuh: the SAME source is compiled by g++ with the
// CUDA execution-space keywords defined away, and every (block, thread) of a launch is run as one
// loop iteration.  The kernels use no shared memory, shuffles or barriers, and their in-place
// passes are race-free by construction (they write only faces whose mask bit is 0 and read only
// faces whose bit is 1), so a sequential sweep is a valid schedule.  tests/test_cpu_emul.py
// compares the results with the oracle; nothing here is linked into libfsb.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h> // float4 / uint4 / make_float4 for the host compiler

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#ifndef __restrict__
#define __restrict__
#endif

using std::fmaf;
using std::max;
using std::min;

struct EmulIdx
{
  unsigned x, y, z;
};
static EmulIdx blockIdx, blockDim, threadIdx, gridDim;

static inline unsigned __vcmpeq4(unsigned a, unsigned b)
{
  unsigned r = 0;
  for (int t = 0; t < 4; ++t)
    if (((a >> (8 * t)) & 0xffu) == ((b >> (8 * t)) & 0xffu)) r |= 0xffu << (8 * t);
  return r;
}
template <class T>
static inline T __ldg(const T* p)
{
  return *p;
}

#define FSB_VEC_WANT_GRID
#define FSB_VEC_WANT_CG
#include "fsb_vec_kernels.cuh"
#include "fsb_mg_kernels.cuh"
#include "fsb_cg_math.cuh"
#include "fsb_cg_one_scalars.h"
#include <cmath>
#include <vector>

namespace {
template <class F>
void launch(unsigned gx, unsigned gy, unsigned threads, F body)
{
  gridDim = {gx, gy, 1};
  blockDim = {threads, 1, 1};
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx)
      for (unsigned tx = 0; tx < threads; ++tx)
      {
        blockIdx = {bx, by, 0};
        threadIdx = {tx, 0, 0};
        body();
      }
}
unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }
} // namespace

extern "C" {

// all grids are pitched: ld = nx rounded up to 32, pad columns zero (labels: SOLID)
void emul_fill_labels(uint8_t* cell, int nx, int ny, int ld, float dx, float dy)
{
  const GridDims d = make_grid_dims(nx, ny, ld, dx, dy);
  launch(div_up(ld, 16 * 256), ny, 256, [&] { k_fill_labels16(cell, d); });
}

void emul_enforce_dirichlet(float* uf, float* vf, const uint8_t* cell, int nx, int ny, int ld,
                            float dx, float dy)
{
  const GridDims d = make_grid_dims(nx, ny, ld, dx, dy);
  launch(div_up(ld, 1024), ny, 256, [&] { k_enforce_dirichlet4(uf, vf, cell, d); });
}

// extendVelocityIndividual with two sweeps: front (uf, vf) -> back (ub, vb); uf is modified
void emul_extend2(float* uf, const float* vf, float* ub, float* vb, uint8_t* m1,
                  const uint8_t* cell, int nx, int ny, int ld, float dx, float dy)
{
  const GridDims d = make_grid_dims(nx, ny, ld, dx, dy);
  launch(div_up(ld, 1024), div_up(ny, kExtendRows), 256,
         [&] { k_extend2_a(uf, vf, ub, vb, m1, cell, d); });
  launch(div_up(ld, 1024), div_up(ny, kExtendRows), 256,
         [&] { k_extend2_b(uf, ub, vb, m1, cell, d); });
}

void emul_pressure_patch(const float* uf, const float* vf, float* ub, float* vb, const float* x,
                         const uint8_t* cell, int nx, int ny, int ld, float dx, float dy, float dt,
                         float density, int dirichlet)
{
  const GridDims d = make_grid_dims(nx, ny, ld, dx, dy);
  // the library picks the compile-time power-of-two form exactly like this
  if (d.pow2 == 3)
  {
    const GridDimsP2 p = as_pow2(d);
    if (dirichlet)
      launch(div_up(ld, 1024), ny, 256,
             [&] { k_pressure_patch4<true, GridDimsP2>(uf, vf, ub, vb, x, cell, p, dt, density); });
    else
      launch(div_up(ld, 1024), ny, 256,
             [&] { k_pressure_patch4<false, GridDimsP2>(uf, vf, ub, vb, x, cell, p, dt, density); });
  }
  else if (dirichlet)
    launch(div_up(ld, 1024), ny, 256,
           [&] { k_pressure_patch4<true, GridDims>(uf, vf, ub, vb, x, cell, d, dt, density); });
  else
    launch(div_up(ld, 1024), ny, 256,
           [&] { k_pressure_patch4<false, GridDims>(uf, vf, ub, vb, x, cell, d, dt, density); });
}

// pressure system set-up; sums[3] = |b|^2, b.z, liquid count
void emul_cg_build(const float* uf, const float* vf, const uint8_t* cell, uint8_t* code, float* r,
                   const float* invdiag5, int nx, int ny, int ld, float dx, float dy, double* sums)
{
  const GridDims d = make_grid_dims(nx, ny, ld, dx, dy);
  sums[0] = sums[1] = sums[2] = 0.0;
  for (int j = 0; j < ny; ++j)
    for (int i0 = 0; i0 < ld; i0 += 4)
    {
      uint32_t cd;
      float4 b;
      if (d.pow2 == 3)
        cg_build_group(uf, vf, cell, as_pow2(d), invdiag5, i0, j, &cd, &b, &sums[0], &sums[1], &sums[2]);
      else
        cg_build_group(uf, vf, cell, d, invdiag5, i0, j, &cd, &b, &sums[0], &sums[1], &sums[2]);
      std::memcpy(code + i0 + (size_t)j * ld, &cd, 4);
      std::memcpy(r + i0 + (size_t)j * ld, &b, 16);
    }
}

// One V-cycle z = V(r) of the multigrid preconditioner, the launch sequence of mg_vcycle() in
// fsb_mg.cu (`sweeps` + `sweeps` damped-Jacobi sweeps, restriction, 40 sweeps on the coarsest level, prolongation;
// coarsening until both sides are <= `stop`; `renorm`: wall-conservative transfer weights) on host arrays.  lab / code0 / r are pitched (ld = nx rounded up to 32); z receives the result.
// Returns the number of levels.
int emul_mg_vcycle(const uint8_t* lab0, const uint8_t* code0, const float* r0, int nx0, int ny0,
                   float inv_h2_0, float* z, int sweeps, int stop, int renorm)
{
  struct Lv
  {
    int nx, ny, ld;
    float inv_h2;
    std::vector<uint8_t> lab, code;
    std::vector<float> x[2], b, r;
  };
  std::vector<Lv> lv;
  int nx = nx0, ny = ny0;
  float inv_h2 = inv_h2_0;
  for (int l = 0; l < 16; ++l)
  {
    Lv L;
    L.nx = nx; L.ny = ny; L.ld = (nx + 31) / 32 * 32; L.inv_h2 = inv_h2;
    const size_t cells = (size_t)L.ld * ny;
    L.lab.assign(cells, FSB_SOLID); L.code.assign(cells, 0);
    L.x[0].assign(cells, 0.f); L.x[1].assign(cells, 0.f); L.b.assign(cells, 0.f); L.r.assign(cells, 0.f);
    lv.push_back(std::move(L));
    if (nx <= stop && ny <= stop) break;
    nx = (nx + 1) / 2; ny = (ny + 1) / 2; inv_h2 *= 0.25f;
  }
  {
    Lv& L = lv[0];
    std::memcpy(L.lab.data(), lab0, L.lab.size());
    std::memcpy(L.code.data(), code0, L.code.size());
    std::memcpy(L.b.data(), r0, L.b.size() * sizeof(float));
  }
  for (size_t l = 1; l < lv.size(); ++l)
  {
    Lv& F = lv[l - 1];
    Lv& C = lv[l];
    launch(div_up(C.ld, 256), C.ny, 256, [&] {
      k_mg_coarsen_labels(F.lab.data(), F.nx, F.ny, F.ld, C.lab.data(), C.nx, C.ny, C.ld);
    });
    launch(div_up(C.ld, 256), C.ny, 256, [&] { k_mg_codes(C.lab.data(), C.code.data(), C.nx, C.ny, C.ld); });
  }
  auto coef = [](float ih2) {
    MgCoef k;
    k.inv_h2 = ih2;
    k.wdinv[0] = 0.0f;
    for (int n = 1; n < 5; ++n) k.wdinv[n] = kMgOmega * (-1.0f / ((float)n * ih2));
    return k;
  };
  const int last = (int)lv.size() - 1;
  std::vector<int> cur(lv.size(), 0);
  auto smooth_first = [&](Lv& L) {
    const MgCoef kf = coef(L.inv_h2);
    launch(div_up(L.ld, 1024), L.ny, 256, [&] {
      k_mg_smooth<true>(nullptr, L.b.data(), L.code.data(), L.x[0].data(), L.nx, L.ny, L.ld, kf);
    });
  };
  auto smooth = [&](Lv& L, int& c) {
    const MgCoef kf = coef(L.inv_h2);
    launch(div_up(L.ld, 1024), L.ny, 256, [&] {
      k_mg_smooth<false>(L.x[c].data(), L.b.data(), L.code.data(), L.x[c ^ 1].data(), L.nx, L.ny, L.ld, kf);
    });
    c ^= 1;
  };
  for (int l = 0; l < last; ++l)
  {
    Lv& L = lv[l];
    cur[l] = 0;
    smooth_first(L);
    for (int s = 1; s < sweeps; ++s) smooth(L, cur[l]); // pre-smoothing
    launch(div_up(L.ld, 1024), L.ny, 256, [&] {
      k_mg_residual(L.x[cur[l]].data(), L.b.data(), L.code.data(), L.r.data(), L.nx, L.ny, L.ld, L.inv_h2);
    });
    Lv& C = lv[l + 1];
    launch(div_up(C.ld, 256), C.ny, 256, [&] {
      k_mg_restrict(L.r.data(), L.nx, L.ny, L.ld, C.code.data(), C.b.data(), C.nx, C.ny, C.ld,
                    renorm ? C.lab.data() : nullptr);
    });
  }
  {
    Lv& L = lv[last];
    cur[last] = 0;
    smooth_first(L);
    for (int s = 1; s < 40; ++s) smooth(L, cur[last]); // k_mg_coarse_solve: 40 sweeps from zero
  }
  for (int l = last - 1; l >= 0; --l)
  {
    Lv& L = lv[l];
    Lv& C = lv[l + 1];
    launch(div_up(L.ld, 1024), L.ny, 256, [&] {
      k_mg_prolong_add(L.x[cur[l]].data(), L.code.data(), L.nx, L.ny, L.ld, C.x[cur[l + 1]].data(), C.nx,
                       C.ny, C.ld, renorm ? C.lab.data() : nullptr);
    });
    for (int s = 0; s < sweeps; ++s) smooth(L, cur[l]); // post-smoothing
  }
  std::memcpy(z, lv[0].x[cur[0]].data(), lv[0].x[cur[0]].size() * sizeof(float));
  return (int)lv.size();
}

// The one-sweep Jacobi-PCG of fsb_cg_one.cu on host arrays: the same per-float4 arithmetic
// (fsb_cg_math.cuh: apply_a4, direction4, dot4, fma4), the same scalar step (one_advance,
// fsb_cg_one_scalars.h), the same sweep structure -- every float4 group recomputes the update of
// iteration k for its four neighbours from the OLD vectors (the kernel's halo recomputation),
// r and p are ping-ponged, x is updated on odd iterations only (two updates back to back) with a
// final pass when the solve ends on an even one.  Partial sums are added in group order.
// code / b pitched (from emul_cg_build); x receives the solution.  Returns the iteration count
// (Eigen's convention); *relres = |r| / |b|.
int emul_cg_one_solve(const uint8_t* code, const float* b, int nx, int ny, int ld, float dx, float tol,
                      int max_iters, float* x, float* relres)
{
  const size_t cells = (size_t)ld * ny;
  std::vector<float> r[2] = {std::vector<float>(b, b + cells), std::vector<float>(cells, 0.f)};
  std::vector<float> p[2] = {std::vector<float>(cells, 0.f), std::vector<float>(cells, 0.f)};
  std::memset(x, 0, cells * sizeof(float));
  const double dx2 = std::pow((double)dx, 2);
  float4 lut[8];
  for (int t = 0; t < 8; ++t) lut[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n = 0; n < 5; ++n)
  {
    const float diag = (float)(-n / dx2);
    lut[n + 1] = make_float4(diag != 0.0f ? 1.0f / diag : 1.0f, diag, (float)(1 / dx2), 0.f);
  }
  const float inv5 = lut[5].x, diag5 = lut[5].y, off = lut[5].z;
  double rhs2 = 0.0;
  for (size_t q = 0; q < cells; ++q) rhs2 += (double)b[q] * (double)b[q];
  OneState ss;
  std::memset(&ss, 0, sizeof ss);
  ss.r2 = rhs2;
  float thr = tol * tol * (float)rhs2; // Eigen: max(tol^2 |b|^2, FLT_MIN)
  if (thr < 1.17549435e-38f) thr = 1.17549435e-38f;
  ss.thr = thr; ss.max_iters = max_iters; ss.sweep = -1;
  if ((float)rhs2 == 0.0f || (float)rhs2 < thr || max_iters <= 0) { *relres = 0.0f; return 0; }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ld4 = [&](const std::vector<float>& v, int i0, int j) -> float4 {
    if (j < 0 || j >= ny || i0 < 0 || i0 >= ld) return zero4; // TMA zero fill
    return *reinterpret_cast<const float4*>(&v[i0 + (size_t)j * ld]);
  };
  auto ld1 = [&](const std::vector<float>& v, int i, int j) -> float {
    if (j < 0 || j >= ny || i < 0 || i >= ld) return 0.0f;
    return v[i + (size_t)j * ld];
  };
  auto cd4 = [&](int i0, int j) -> uint32_t {
    if (j < 0 || j >= ny || i0 < 0 || i0 >= ld) return 0u;
    return *reinterpret_cast<const uint32_t*>(code + i0 + (size_t)j * ld);
  };
  auto cd1 = [&](int i, int j) -> uint32_t {
    if (j < 0 || j >= ny || i < 0 || i >= ld) return 0u;
    return code[i + (size_t)j * ld];
  };
  int cur = 0;
  float alpha_prev = 0.0f;
  bool pending = false;
  const std::vector<float>* last_p = nullptr;
  for (int sweep = -1; !ss.done; ++sweep, cur ^= 1)
  {
    const float alpha = ss.alpha, beta = ss.beta, nalpha = -alpha;
    const std::vector<float>&ro = r[cur], &po = p[cur];
    std::vector<float>&rnw = r[cur ^ 1], &pnw = p[cur ^ 1];
    const bool with_x = sweep >= 0 && (sweep & 1);
    // the update of iteration k for one float4 group, from the old vectors only
    auto upd4 = [&](int i0, int j, float4* rn, float4* pn) {
      const float4 pc = ld4(po, i0, j);
      const uint32_t c4 = cd4(i0, j);
      const float4 q = apply_a4(pc, ld1(po, i0 - 1, j), ld1(po, i0 + 4, j), ld4(po, i0, j - 1),
                                ld4(po, i0, j + 1), c4, lut, diag5, off);
      *rn = fma4(nalpha, q, ld4(ro, i0, j));
      *pn = direction4(*rn, pc, c4, lut, inv5, beta);
    };
    auto upd1 = [&](int i, int j) -> float { // the same for one halo-column cell
      const float pc = ld1(po, i, j);
      const float4 kf = lut[cd1(i, j)];
      const float q = fmaf(kf.y, pc, kf.z * ((ld1(po, i - 1, j) + ld1(po, i + 1, j)) +
                                             (ld1(po, i, j - 1) + ld1(po, i, j + 1))));
      const float rr = fmaf(nalpha, q, ld1(ro, i, j));
      return fmaf(beta, pc, kf.x * rr);
    };
    double acc[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < ny; ++j)
      for (int i0 = 0; i0 < ld; i0 += 4)
      {
        float4 rn, pn, rs, ps, rnn, pnn;
        upd4(i0, j, &rn, &pn);
        upd4(i0, j - 1, &rs, &ps);
        upd4(i0, j + 1, &rnn, &pnn);
        const uint32_t c4 = cd4(i0, j);
        const float4 q2 = apply_a4(pn, upd1(i0 - 1, j), upd1(i0 + 4, j), ps, pnn, c4, lut, diag5, off);
        const size_t o = i0 + (size_t)j * ld;
        if (with_x)
        {
          float4 xn = *reinterpret_cast<float4*>(x + o);
          xn = fma4(alpha_prev, *reinterpret_cast<const float4*>(&pnw[o]), xn); // p_{k-1}
          xn = fma4(alpha, ld4(po, i0, j), xn);
          *reinterpret_cast<float4*>(x + o) = xn;
        }
        *reinterpret_cast<float4*>(&pnw[o]) = pn;
        *reinterpret_cast<float4*>(&rnw[o]) = rn;
        float4 iv = make_float4(inv5, inv5, inv5, inv5);
        if (c4 != kInterior4)
          iv = make_float4(lut[c4 & 0xff].x, lut[(c4 >> 8) & 0xff].x, lut[(c4 >> 16) & 0xff].x, lut[c4 >> 24].x);
        const float4 z = make_float4(iv.x * rn.x, iv.y * rn.y, iv.z * rn.z, iv.w * rn.w);
        const float4 mq = make_float4(iv.x * q2.x, iv.y * q2.y, iv.z * q2.z, iv.w * q2.w);
        acc[0] += (double)dot4(pn, q2);
        acc[1] += (double)dot4(rn, z);
        acc[2] += (double)dot4(rn, rn);
        acc[3] += (double)dot4(z, q2);
        acc[4] += (double)dot4(q2, mq);
      }
    one_advance(&ss, acc[0], acc[1], acc[2], acc[3], acc[4]);
    if (sweep >= 0)
    {
      alpha_prev = alpha;
      pending = !with_x;
      last_p = &p[cur];
    }
  }
  if (pending)
    for (size_t q = 0; q < cells; ++q) x[q] = fmaf(alpha_prev, (*last_p)[q], x[q]);
  *relres = (float)std::sqrt((float)ss.r2 / (float)rhs2);
  return ss.iter;
}

} // extern "C"
