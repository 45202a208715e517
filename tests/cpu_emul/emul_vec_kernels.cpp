// TEST INFRASTRUCTURE ONLY.  Host emulation of the vectorised stage kernels of
// fluid_simulation_b200/csrc/fsb_vec_kernels.cuh: the SAME source is compiled by g++ with the
// CUDA execution-space keywords defined away, and every (block, thread) of a launch is run as one
// loop iteration.  The kernels use no shared memory, shuffles or barriers, and their in-place
// passes are race-free by construction (they write only faces whose mask bit is 0 and read only
// faces whose bit is 1), so a sequential sweep is a valid schedule.  tests/test_cpu_emul.py
// compares the results with the oracle; nothing here is linked into libfsb.so.
#include <algorithm>
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

} // extern "C"
