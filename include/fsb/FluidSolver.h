// FluidSolverMemoryPool and FluidSolver of the reference (include/FluidSolver.h:14-146,
// src/FluidSolver.cpp) as host handles: every step* call is ONE call into libfsb (fsb_step), which
// runs classification, P2G, the grid passes, the matrix-free CG pressure solve, G2P and particle
// advection as CUDA kernels on the domain's stream with no host work in between.
#ifndef FSB_FLUID_SOLVER_H
#define FSB_FLUID_SOLVER_H

#include <cmath>
#include <stdexcept>

#include "FluidDomain.h"

// The scratch grids of the reference's pool (liquid indices, particle counts, validity masks, P2G
// accumulators) are the device workspace inside the fsb context; on the host the pool is only the
// shape it was built for, which is what FluidSolver::validate compares.
class FluidSolverMemoryPool : public GridInterface
{
public:
  FluidSolverMemoryPool(int size_x, int size_y, MyFloat delta_x, MyFloat delta_y)
      : GridInterface(size_x, size_y, delta_x, delta_y) {}
  FluidSolverMemoryPool(const FluidDomain& fluid_domain)
      : GridInterface(fluid_domain.sizeX(), fluid_domain.sizeY(), fluid_domain.deltaX(),
                      fluid_domain.deltaY()) {}
  // src/FluidSolver.cpp:56-65: the reference's copy constructor passes deltaX for BOTH deltas,
  // which is why a solver only ever validates against domains with square cells.  Kept.
  FluidSolverMemoryPool(const FluidSolverMemoryPool& other)
      : GridInterface(other.sizeX(), other.sizeY(), other.deltaX(), other.deltaX()) {}
};

class FluidSolver
{
public:
  FluidSolver(FluidSolverMemoryPool mem_pool) // by value, then copied again, as upstream (:78-82)
      : _mem_pool(mem_pool), _max_iterations(100), _tolerance(1.1920929e-7f) {}

  void stepSemiLagrangian(FluidDomain& fluid_domain, MyFloat dt) { step(fluid_domain, FSB_STEP_SEMILAGRANGIAN, dt); }
  void stepPIC(FluidDomain& fluid_domain, MyFloat dt) { step(fluid_domain, FSB_STEP_PIC, dt); }
  void stepFLIP(FluidDomain& fluid_domain, MyFloat dt) { step(fluid_domain, FSB_STEP_FLIP, dt); }
  void stepPICFLIP(FluidDomain& fluid_domain, MyFloat dt) { step(fluid_domain, FSB_STEP_PICFLIP, dt); }

  // ---- additions (the reference hard-codes these: src/FluidSolver.cpp:81, Eigen's default tol)
  void setMaxIterations(int max_iterations) { _max_iterations = max_iterations; }
  void setTolerance(MyFloat tolerance) { _tolerance = tolerance; }
  int iterations() const { return _last_iterations; }
  MyFloat error() const { return _last_error; }
  // opt-in geometric-multigrid preconditioner for the pressure CG (same system and stopping rule,
  // tens of iterations instead of thousands; the default is the reference's diagonal one)
  void setMultigridPreconditioner(bool on) { _multigrid = on; }

private:
  void step(FluidDomain& fluid_domain, int kind, MyFloat dt)
  {
    const fsb::ContextPtr& dev = fluid_domain.device();
    // validate() of the reference runs inside fsb_step against the pool registered here and
    // fails with the reference's message, rethrown as std::runtime_error
    dev->check(fsb_set_pool(dev->get(), _mem_pool.sizeX(), _mem_pool.sizeY(), _mem_pool.deltaX(),
                            _mem_pool.deltaY()));
    dev->check(fsb_set_cg(dev->get(), _max_iterations, _tolerance));
    dev->check(fsb_set_preconditioner(dev->get(), _multigrid ? FSB_PRECOND_MULTIGRID : FSB_PRECOND_JACOBI));
    fluid_domain.sync_to_device();
    dev->check(fsb_step(dev->get(), kind, dt));
    fluid_domain.device_changed();
    dev->check(fsb_get_cg_info(dev->get(), &_last_iterations, &_last_error));
  }

  FluidSolverMemoryPool _mem_pool;
  int _max_iterations;
  MyFloat _tolerance;
  int _last_iterations = 0;
  MyFloat _last_error = 0;
  bool _multigrid = false;
};

#endif
