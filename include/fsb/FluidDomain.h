// FluidSource and FluidDomain of the reference (include/FluidDomain.h:10-76, src/FluidDomain.cpp)
// on top of one libfsb device context.  The domain owns the context; its MacGrid and
// MarkerParticleSet members are views of it.  The LevelSet of the reference is not part of any
// step* path (README "Not Yet Implemented") and is not provided.
#ifndef FSB_FLUID_DOMAIN_H
#define FSB_FLUID_DOMAIN_H

#include <vector>

#include "DeviceContext.h"
#include "MacGrid.h"
#include "MarkerParticleSet.h"
#include "MathDefinitions.h"

class FluidSource
{
public:
  FluidSource(BBox<MyFloat> area, MyFloat delta_x, MyFloat delta_y, MyFloat x_velocity,
              MyFloat y_velocity, MyFloat time_step, int max_spawns)
      : _area(area), _x_velocity(x_velocity), _y_velocity(y_velocity), _delta_x(delta_x),
        _delta_y(delta_y), _time_step(time_step), _time_since_last(0.0), _max_spawns(max_spawns),
        _n_spawns(0) {}

  // src/FluidDomain.cpp:29-52.  The lattice is generated ON THE DEVICE (fsb_emit_source), in the
  // reference's order, after whatever the caller queued with addParticle.
  void update(MarkerParticleSet& particle_set, MyFloat dt)
  {
    if (isFinished()) return;
    if (_time_since_last >= _time_step)
    {
      const fsb::ContextPtr& dev = particle_set.device();
      if (!dev) throw std::runtime_error("FluidSource::update: the set is not part of a FluidDomain");
      particle_set.sync_to_device();
      dev->check(fsb_emit_source(dev->get(), _area.x_min, _area.x_max, _area.y_min, _area.y_max,
                                 _delta_x, _delta_y, _x_velocity, _y_velocity, nullptr));
      particle_set.device_appended();
      _time_since_last = 0;
      _n_spawns++;
    }
    _time_since_last += dt;
  }
  bool isFinished() { return _max_spawns != -1 && _n_spawns >= _max_spawns; }
  void resetSpawns() { _n_spawns = 0; }

private:
  BBox<MyFloat> _area;
  MyFloat _x_velocity, _y_velocity, _delta_x, _delta_y, _time_step, _time_since_last;
  int _max_spawns, _n_spawns;
};

class FluidDomain : public GridInterface
{
public:
  // `device`: CUDA ordinal (addition; defaults to fsb::DeviceContext::defaultDevice()).
  FluidDomain(int size_x, int size_y, MyFloat length_x, MyFloat length_y, MyFloat density,
              MyFloat pic_ratio, int device = -1)
      : GridInterface(size_x, size_y, length_x / size_x, length_y / size_y),
        _density(density), _pic_ratio(pic_ratio),
        _dev(std::make_shared<fsb::DeviceContext>(
            size_x, size_y, length_x, length_y, density, pic_ratio,
            device < 0 ? fsb::DeviceContext::defaultDevice() : device)),
        _mac_grid(_dev, size_x, size_y, length_x, length_y)
  {
    _particle_set.bind(_dev);
  }
  // the context is shared by the member views; copying a domain would alias device state
  FluidDomain(const FluidDomain&) = delete;
  FluidDomain& operator=(const FluidDomain&) = delete;

  void addFluidSource(FluidSource fluid_source) { _fluid_sources.push_back(fluid_source); }
  void update(MyFloat dt) // src/FluidDomain.cpp:80-87
  {
    for (size_t i = 0; i < _fluid_sources.size(); ++i) _fluid_sources[i].update(_particle_set, dt);
  }
  void clearFluidSources() { _fluid_sources.clear(); }
  void resetParticleSet() { _particle_set.clear(); }
  void setPicRatio(MyFloat pic_ratio) // src/FluidDomain.cpp:99-102
  {
    _pic_ratio = CLAMP(pic_ratio, 0, 1);
    _dev->check(fsb_set_pic_ratio(_dev->get(), _pic_ratio));
  }

  MacGrid& macGrid() { return _mac_grid; }
  MarkerParticleSet& markerParticleSet() { return _particle_set; }
  const MacGrid& macGrid() const { return _mac_grid; }
  const MarkerParticleSet& markerParticleSet() const { return _particle_set; }
  const MyFloat density() const { return _density; }
  const MyFloat picRatio() const { return _pic_ratio; }

  // src/FluidDomain.cpp:150-180.  Like the reference, the argument is ignored and the domain's
  // own set is classified (:157 iterates _particle_set).
  void classifyCells(MarkerParticleSet&)
  {
    sync_to_device();
    _dev->check(fsb_classify_cells(_dev->get()));
    _mac_grid.device_changed();
  }

  // ---- device coherence (used by FluidSolver)
  const fsb::ContextPtr& device() const { return _dev; }
  void sync_to_device()
  {
    _mac_grid.sync_to_device();
    _particle_set.sync_to_device();
  }
  void device_changed()
  {
    _mac_grid.device_changed();
    _particle_set.device_changed();
  }

private:
  MyFloat _density, _pic_ratio;
  fsb::ContextPtr _dev;
  MacGrid _mac_grid;
  MarkerParticleSet _particle_set;
  std::vector<FluidSource> _fluid_sources;
};

#endif
