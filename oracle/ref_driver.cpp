// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// C-ABI driver around the reference's OWN classes.  It is linked with the
// reference's five simulation sources compiled unchanged from /root/reference
// (see oracle/Makefile) and gives the tests stage-level access to
// FluidSolver's private methods (src/FluidSolver.cpp:253-963) and to
// MacGrid's private buffers.  The access-specifier override below only
// affects THIS translation unit; the reference translation units are
// compiled as they are, and GCC lays members out in declaration order
// regardless of access, so the object layout is the same on both sides.
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include <Eigen/SparseCore>
#include <Eigen/IterativeLinearSolvers>

#define private public
#define protected public
#include <FluidSolver.h>
#include <FluidDomain.h>
#include <Renderer.h>
#undef private
#undef protected

#define FSX(name) fsr_##name
#include "fluid_oracle_api.h"

namespace {

struct RefCtx
{
  FluidDomain domain;
  FluidSolver solver;
  bool solved; // pressureSolve ran past its early return at least once
  RefCtx(int sx, int sy, float lx, float ly, float density, float pic_ratio)
      : domain(sx, sy, lx, ly, density, pic_ratio),
        solver(FluidSolverMemoryPool(domain)), // examples/simple.cpp:23-24
        solved(false)
  {
  }
};

RefCtx* C(void* h) { return static_cast<RefCtx*>(h); }

Grid<MyFloat>* pickGrid(MacGrid& g, int which)
{
  switch (which)
  {
  case FSX_U_FRONT: return g._vel_x_front_buffer.get();
  case FSX_V_FRONT: return g._vel_y_front_buffer.get();
  case FSX_U_BACK: return g._vel_x_back_buffer.get();
  case FSX_V_BACK: return g._vel_y_back_buffer.get();
  case FSX_U_PREV: return &g._vel_x_previous;
  case FSX_V_PREV: return &g._vel_y_previous;
  case FSX_U_DIFF: return &g._vel_x_diff;
  case FSX_V_DIFF: return &g._vel_y_diff;
  }
  return nullptr;
}

} // namespace

extern "C" {

void* fsr_create(int size_x, int size_y, float length_x, float length_y,
                 float density, float pic_ratio)
{
  return new RefCtx(size_x, size_y, length_x, length_y, density, pic_ratio);
}
void fsr_destroy(void* h) { delete C(h); }
float fsr_delta_x(void* h) { return C(h)->domain.deltaX(); }
float fsr_delta_y(void* h) { return C(h)->domain.deltaY(); }

void fsr_set_cg(void* h, int max_iters, float tol)
{
  C(h)->solver._cg_solver.setMaxIterations(max_iters);
  C(h)->solver._cg_solver.setTolerance(tol);
}

void fsr_append_particles(void* h, const float* a, int64_t n)
{
  MarkerParticleSet& ps = C(h)->domain.markerParticleSet();
  for (int64_t k = 0; k < n; ++k)
    ps.addParticle(MarkerParticle(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]));
}
void fsr_set_particles(void* h, const float* a, int64_t n)
{
  C(h)->domain.resetParticleSet();
  fsr_append_particles(h, a, n);
}
int64_t fsr_num_particles(void* h) { return C(h)->domain.markerParticleSet().size(); }
void fsr_get_particles(void* h, float* a)
{
  MarkerParticleSet& ps = C(h)->domain.markerParticleSet();
  int64_t k = 0;
  for (auto it = ps.begin(); it != ps.end(); ++it, ++k)
  {
    a[4 * k] = it->posX();
    a[4 * k + 1] = it->posY();
    a[4 * k + 2] = it->velX();
    a[4 * k + 3] = it->velY();
  }
}
int64_t fsr_emit_source(void* h, float x_min, float x_max, float y_min, float y_max,
                        float delta_x, float delta_y, float vel_x, float vel_y)
{
  MarkerParticleSet& ps = C(h)->domain.markerParticleSet();
  const int64_t before = ps.size();
  // time_step 0, one spawn: fires on the first update (src/FluidDomain.cpp:34)
  FluidSource src({x_min, x_max, y_min, y_max}, delta_x, delta_y, vel_x, vel_y, 0.0, 1);
  src.update(ps, 0.0);
  return ps.size() - before;
}

void fsr_set_grid(void* h, int which, const float* src)
{
  Grid<MyFloat>* g = pickGrid(C(h)->domain.macGrid(), which);
  std::memcpy(g->data.data(), src, sizeof(float) * g->data.size());
}
void fsr_get_grid(void* h, int which, float* dst)
{
  Grid<MyFloat>* g = pickGrid(C(h)->domain.macGrid(), which);
  std::memcpy(dst, g->data.data(), sizeof(float) * g->data.size());
}
void fsr_set_cell_types(void* h, const uint8_t* src)
{
  MacGrid& g = C(h)->domain.macGrid();
  for (size_t k = 0; k < g._cell_type_buffer.data.size(); ++k)
    g._cell_type_buffer.data[k] = (CellType)src[k];
}
void fsr_get_cell_types(void* h, uint8_t* dst)
{
  MacGrid& g = C(h)->domain.macGrid();
  for (size_t k = 0; k < g._cell_type_buffer.data.size(); ++k)
    dst[k] = (uint8_t)g._cell_type_buffer.data[k];
}

void fsr_classify_cells(void* h)
{
  C(h)->domain.classifyCells(C(h)->domain.markerParticleSet());
}
void fsr_p2g_spread(void* h)
{
  C(h)->solver.transferVelocityToGridSpread(C(h)->domain.markerParticleSet(),
                                            C(h)->domain.macGrid());
}
void fsr_save_previous(void* h) { C(h)->domain.macGrid().updatePreviousVelocityBuffer(); }
void fsr_add_acceleration(void* h, float ax, float ay, float dt)
{
  C(h)->solver.addExternalAcceleration(C(h)->domain.macGrid(), ax, ay, dt);
}
void fsr_enforce_dirichlet(void* h) { C(h)->solver.enforceDirichlet(C(h)->domain.macGrid()); }
void fsr_extend_velocity(void* h, int n_iter)
{
  C(h)->solver.extendVelocityIndividual(C(h)->domain.macGrid(), n_iter);
}
void fsr_pressure_solve(void* h, float density, float dt)
{
  RefCtx* c = C(h);
  c->solver.pressureSolve(c->domain.macGrid(), c->domain.markerParticleSet(), density, dt);
}
void fsr_get_pressure(void* h, float* dst)
{
  RefCtx* c = C(h);
  const Grid<int>& idx = c->solver._mem_pool.fluid_indices;
  const Eigen::VectorXf& x = c->solver._cg_solver.last_solution;
  const size_t n = idx.data.size();
  for (size_t k = 0; k < n; ++k)
  {
    const int id = idx.data[k];
    dst[k] = (id >= 0 && id < (int)x.size()) ? x[id] : 0.0f;
  }
}
int fsr_cg_iterations(void* h) { return (int)C(h)->solver._cg_solver.iterations(); }
float fsr_cg_error(void* h) { return C(h)->solver._cg_solver.error(); }
void fsr_update_diff(void* h) { C(h)->domain.macGrid().updateVelocityDiffBuffer(); }
void fsr_g2p(void* h, int mode, float pic_ratio)
{
  RefCtx* c = C(h);
  MacGrid& g = c->domain.macGrid();
  MarkerParticleSet& ps = c->domain.markerParticleSet();
  if (mode == FSX_G2P_PIC) c->solver.transferVelocityToParticlesPIC(g, ps);
  else if (mode == FSX_G2P_FLIP) c->solver.transferVelocityToParticlesFLIP(g, ps);
  else c->solver.transferVelocityToParticlesPICFLIP(g, ps, pic_ratio);
}
void fsr_advect_particles(void* h, float dt, int ensure_outside)
{
  RefCtx* c = C(h);
  if (ensure_outside)
    c->domain.markerParticleSet().advectAndEnsureOutsideObstacles(dt, c->domain.macGrid());
  else
    c->domain.markerParticleSet().advect(dt);
}
void fsr_advect_velocity_sl(void* h, float dt)
{
  C(h)->solver.advectVelocitySemiLagrangian(C(h)->domain.macGrid(), dt);
}
void fsr_advect_particles_grid(void* h, float dt)
{
  RefCtx* c = C(h);
  c->solver.advectParticlesWithGrid(c->domain.markerParticleSet(), c->domain.macGrid(), dt);
}
void fsr_add_external_force(void* h, float fx, float fy, float dt)
{
  C(h)->solver.addExternalForce(C(h)->domain, fx, fy, dt);
}
void fsr_p2g_gather(void* h)
{
  C(h)->solver.transferVelocityToGridGather(C(h)->domain.markerParticleSet(),
                                            C(h)->domain.macGrid());
}
void fsr_extend_velocity_avg(void* h, int n_iterations)
{
  C(h)->solver.extendVelocityAvarageing(C(h)->domain.macGrid(), n_iterations);
}
void fsr_render_rgb(void* h, int width, int height, float x_min, float x_max, float y_min,
                    float y_max, uint8_t* rgb)
{
  // the reference's own frame path, including its PPM writer (read back from a temporary file)
  RefCtx* c = C(h);
  Canvas canvas(width, height);
  Renderer renderer(BBox<MyFloat>{x_min, x_max, y_min, y_max});
  renderer.clearCanvas(canvas);
  renderer.renderGridCellsToCanvas(c->domain.macGrid(), canvas);
  renderer.renderParticlesToCanvas(c->domain.markerParticleSet(), canvas);
  char path[] = "/tmp/fsr_frame_XXXXXX";
  const int fd = mkstemp(path);
  if (fd >= 0) close(fd);
  renderer.writeCanvasToPpm(path, canvas);
  FILE* f = fopen(path, "rb");
  if (f)
  {
    int w = 0, hh = 0, mx = 0;
    if (fscanf(f, "P6\n%d %d\n%d", &w, &hh, &mx) == 3 && w == width && hh == height)
    {
      fgetc(f); // the single whitespace after the maxval
      if (fread(rgb, 1, (size_t)width * height * 3, f) != (size_t)width * height * 3) memset(rgb, 0, 3);
    }
    fclose(f);
  }
  remove(path);
}
int fsr_step(void* h, int kind, float dt)
{
  RefCtx* c = C(h);
  try
  {
    switch (kind)
    {
    case FSX_STEP_SEMILAGRANGIAN: c->solver.stepSemiLagrangian(c->domain, dt); break;
    case FSX_STEP_PIC: c->solver.stepPIC(c->domain, dt); break;
    case FSX_STEP_FLIP: c->solver.stepFLIP(c->domain, dt); break;
    default: c->solver.stepPICFLIP(c->domain, dt); break;
    }
  }
  catch (const std::runtime_error&)
  {
    return 1;
  }
  return 0;
}

} // extern "C"
