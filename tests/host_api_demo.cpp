// Drives the host classes of include/fsb/ the way the reference's examples/simple.cpp:20-83 drives
// fluidsim_lib (same scene, same frame loop, no rendering) and prints one line per step that
// tests/test_host_api.py compares with the CPU checker.  Also exercises the accessors a renderer
// uses and the reference's error behaviour.
//
//   host_api_demo [n_steps] [grid] [kind: picflip|flip|pic|sl]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include <FluidSolver.h>

static const MyFloat WORLD_X_SIZE = 1.0;
static const MyFloat WORLD_Y_SIZE = 1.0;

static void print_state(int step, FluidDomain& fluid_domain, const FluidSolver& fluid_solver)
{
  const MacGrid& grid = fluid_domain.macGrid();
  const MarkerParticleSet& particles = fluid_domain.markerParticleSet();
  int n_liquid = 0, n_solid = 0;
  for (int j = 0; j < grid.sizeY(); ++j)
    for (int i = 0; i < grid.sizeX(); ++i)
    {
      const CellType t = grid.cellType(i, j);
      n_liquid += (t == LIQUID);
      n_solid += (t == SOLID);
    }
  double sx = 0, sy = 0, su = 0, sv = 0;
  for (MarkerParticleSet::const_iterator it = particles.begin(); it != particles.end(); ++it)
  {
    sx += it->posX();
    sy += it->posY();
    su += it->velX();
    sv += it->velY();
  }
  const int n = particles.size();
  // cell-centred velocity of the middle cell, the way Renderer::renderGridVelocitiesToCanvas reads it
  const int ci = grid.sizeX() / 4, cj = grid.sizeY() / 4;
  std::printf("STEP %d particles %d liquid %d solid %d mean %.17g %.17g %.17g %.17g cg %d %.9g probe %.9g %.9g\n",
              step, n, n_liquid, n_solid, sx / n, sy / n, su / n, sv / n, fluid_solver.iterations(),
              (double)fluid_solver.error(), (double)grid.velX(ci, cj), (double)grid.velY(ci, cj));
}

int main(int argc, char** argv)
{
  const int n_steps = argc > 1 ? std::atoi(argv[1]) : 10;
  const int GRID = argc > 2 ? std::atoi(argv[2]) : 64;
  const char* kind = argc > 3 ? argv[3] : "picflip";
  const MyFloat DELTA_X = WORLD_X_SIZE / GRID, DELTA_Y = WORLD_Y_SIZE / GRID;

  try
  {
    FluidDomain fluid_domain(GRID, GRID, WORLD_X_SIZE, WORLD_Y_SIZE, 0.01, 0.05);
    FluidSolverMemoryPool mem_pool(fluid_domain);
    FluidSolver fluid_solver(mem_pool);
    fluid_domain.addFluidSource(FluidSource(
        {(MyFloat)(2.0 / GRID), (MyFloat)0.35, (MyFloat)(2.0 / GRID), (MyFloat)(1 - 2.0 / GRID)},
        DELTA_X, DELTA_Y, 0.0, 0.0, 0.0, 1));

    // the reference's frame loop: two clamped sub-steps of 0.01 per 0.02 s frame
    const MyFloat seconds_per_frame = 0.02;
    int step = 0;
    while (step < n_steps)
    {
      MyFloat dt;
      for (MyFloat frame_time = 0; frame_time < seconds_per_frame && step < n_steps; frame_time += dt)
      {
        dt = 0.01;
        dt = CLAMP(dt, 0, seconds_per_frame - frame_time);
        fluid_domain.update(dt);
        if (!std::strcmp(kind, "picflip")) fluid_solver.stepPICFLIP(fluid_domain, dt);
        else if (!std::strcmp(kind, "flip")) fluid_solver.stepFLIP(fluid_domain, dt);
        else if (!std::strcmp(kind, "pic")) fluid_solver.stepPIC(fluid_domain, dt);
        else fluid_solver.stepSemiLagrangian(fluid_domain, dt);
        print_state(step, fluid_domain, fluid_solver);
        ++step;
      }
    }

    // host edits between steps (FluidInteractionHandler.cpp:28-33,52-64 does this): add a
    // particle, edit one through the iterator, change the ratio, poke a face and a label
    MarkerParticleSet& set = fluid_domain.markerParticleSet();
    const int before = set.size();
    set.addParticle(MarkerParticle(0.5, 0.5, 0.25, -0.5));
    fluid_domain.setPicRatio(3.0); // clamped to 1
    fluid_domain.macGrid().setVelXHalfIndexed(5, 5, 1.5f);
    fluid_domain.macGrid().setCellType(GRID / 2, GRID / 2, SOLID);
    fluid_solver.stepPICFLIP(fluid_domain, 0.01);
    std::printf("EDIT particles %d -> %d picratio %.3g last %.9g %.9g\n", before, set.size(),
                (double)fluid_domain.picRatio(), (double)(set.end() - 1)->posX(),
                (double)(set.end() - 1)->posY());
    set.begin()->setPosition(0.25f, 0.75f); // write through the non-const iterator
    fluid_solver.stepPICFLIP(fluid_domain, 0.01);
    std::printf("EDIT2 first %.9g %.9g\n", (double)set.cbegin()->posX(), (double)set.cbegin()->posY());
    fluid_domain.resetParticleSet();
    fluid_solver.stepPICFLIP(fluid_domain, 0.01);
    std::printf("RESET particles %d liquid0 %d\n", set.size(), (int)(fluid_domain.macGrid().cellType(5, 5) == LIQUID));
  }
  catch (const std::runtime_error& e)
  {
    std::cout << "UNEXPECTED " << e.what() << std::endl;
    return EXIT_FAILURE;
  }

  // the reference throws from every step* when the pool does not match the domain
  // (src/FluidSolver.cpp:89-107): non-square cells always fail because of the pool copy
  try
  {
    FluidDomain tall(32, 16, 1.0, 1.0, 0.01, 0.05);
    FluidSolverMemoryPool pool(tall);
    FluidSolver solver(pool);
    solver.stepPICFLIP(tall, 0.01);
    std::printf("ERROR-CASE no exception\n");
    return EXIT_FAILURE;
  }
  catch (const std::runtime_error& e)
  {
    std::printf("ERROR-CASE %s\n", e.what());
  }
  return EXIT_SUCCESS;
}
