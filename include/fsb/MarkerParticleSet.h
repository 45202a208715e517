// MarkerParticle / MarkerParticleSet of the reference (include/MarkerParticleSet.h:9-78) over a
// particle set that lives in HBM.
//
// The device keeps the particles cell-sorted; the set the caller sees is always in the caller's
// own order (the order of addParticle calls), exactly like the reference's std::vector.
// Coherence rules:
//   * addParticle queues the particle on the host; the queue is appended on the device before
//     the next device stage (one H2D copy for all queued particles);
//   * begin()/end() make the host mirror current (one D2H copy after device work) -- the const
//     overloads only read, the non-const overloads also mark the set as possibly edited, so it is
//     uploaded again before the next device stage;
//   * advect / advectAndEnsureOutsideObstacles run on the device.
// A set that is not bound to a FluidDomain is a plain host container; its advect functions throw
// because there is no CPU implementation of any stage.
#ifndef FSB_MARKER_PARTICLE_SET_H
#define FSB_MARKER_PARTICLE_SET_H

#include <cstddef>
#include <vector>

#include "DeviceContext.h"
#include "MacGrid.h"
#include "MathDefinitions.h"

class MarkerParticle
{
public:
  MarkerParticle() : _pos_x(0), _pos_y(0), _vel_x(0), _vel_y(0) {}
  MarkerParticle(MyFloat pos_x, MyFloat pos_y, MyFloat vel_x = 0, MyFloat vel_y = 0)
      : _pos_x(pos_x), _pos_y(pos_y), _vel_x(vel_x), _vel_y(vel_y) {}

  MyFloat posX() const { return _pos_x; }
  MyFloat posY() const { return _pos_y; }
  MyFloat velX() const { return _vel_x; }
  MyFloat velY() const { return _vel_y; }
  void setPosition(MyFloat pos_x, MyFloat pos_y) { _pos_x = pos_x; _pos_y = pos_y; }
  void setVelocity(MyFloat vel_x, MyFloat vel_y) { _vel_x = vel_x; _vel_y = vel_y; }
  // single-particle convenience kept from the reference (include/MarkerParticleSet.h:37-41);
  // bulk advection is MarkerParticleSet::advect on the device
  void advect(MyFloat dt) { _pos_x += _vel_x * dt; _pos_y += _vel_y * dt; }

private:
  MyFloat _pos_x, _pos_y, _vel_x, _vel_y; // the 16-byte AoS record libfsb exchanges
};
static_assert(sizeof(MarkerParticle) == 4 * sizeof(float), "MarkerParticle must be the AoS float4 record");

class MarkerParticleSet
{
public:
  typedef std::vector<MarkerParticle>::iterator iterator;
  typedef std::vector<MarkerParticle>::const_iterator const_iterator;

  MarkerParticleSet(int size = 0) : _host(size), _host_valid(true), _host_edited(size > 0) {}

  void addParticle(MarkerParticle p)
  {
    if (_host_valid) _host.push_back(p); // mirror stays current
    if (_dev && !_host_edited) _queued.push_back(p);
  }
  void reserve(int particle_count) { _host.reserve(particle_count); }
  void clear()
  {
    _host.clear();
    _queued.clear();
    _host_valid = true;
    _host_edited = true; // the device set is replaced (by an empty one) at the next sync
  }
  void advect(MyFloat dt) // src/MarkerParticleSet.cpp:40-46
  {
    require_device("MarkerParticleSet::advect");
    sync_to_device();
    _dev->check(fsb_advect_particles(_dev->get(), dt, 0));
    device_changed();
  }
  // src/MarkerParticleSet.cpp:48-62.  `mac_grid` must be the grid of the domain this set belongs
  // to (the labels are read on the device).
  void advectAndEnsureOutsideObstacles(MyFloat dt, const MacGrid& mac_grid)
  {
    require_device("MarkerParticleSet::advectAndEnsureOutsideObstacles");
    if (mac_grid.device().get() != _dev.get())
      throw std::runtime_error("advectAndEnsureOutsideObstacles: the MacGrid belongs to another domain");
    const_cast<MacGrid&>(mac_grid).sync_to_device(); // pending label edits
    sync_to_device();
    _dev->check(fsb_advect_particles(_dev->get(), dt, 1));
    device_changed();
  }

  int size() const
  {
    if (_host_valid || !_dev) return (int)_host.size();
    return (int)(fsb_num_particles(_dev->get()) + (int64_t)_queued.size());
  }

  iterator begin() { fetch(); _host_edited = true; _queued.clear(); return _host.begin(); }
  iterator end() { fetch(); _host_edited = true; _queued.clear(); return _host.end(); }
  const_iterator begin() const { fetch(); return _host.begin(); }
  const_iterator end() const { fetch(); return _host.end(); }
  const_iterator cbegin() const { return begin(); }
  const_iterator cend() const { return end(); }

  // ---- device coherence (used by FluidDomain / FluidSolver)
  void bind(const fsb::ContextPtr& dev) { _dev = dev; }
  const fsb::ContextPtr& device() const { return _dev; }
  void sync_to_device()
  {
    if (!_dev) return;
    if (_host_edited)
    {
      _dev->check(fsb_set_particles(_dev->get(), reinterpret_cast<const float*>(_host.data()),
                                    (int64_t)_host.size()));
      _host_edited = false;
    }
    else if (!_queued.empty())
    {
      _dev->check(fsb_append_particles(_dev->get(), reinterpret_cast<const float*>(_queued.data()),
                                       (int64_t)_queued.size()));
    }
    _queued.clear();
  }
  void device_changed() { _host_valid = false; }
  // particles appended on the device itself (FluidSource): the mirror is stale
  void device_appended() { _host_valid = false; }

private:
  void require_device(const char* what) const
  {
    if (!_dev)
      throw std::runtime_error(std::string(what) +
                               ": the set is not part of a FluidDomain; there is no CPU path");
  }
  void fetch() const
  {
    if (_host_valid || !_dev) return;
    const_cast<MarkerParticleSet*>(this)->sync_to_device(); // queued particles first
    _host.resize((size_t)fsb_num_particles(_dev->get()));
    _dev->check(fsb_get_particles(_dev->get(), reinterpret_cast<float*>(_host.data())));
    _host_valid = true;
  }

  fsb::ContextPtr _dev;
  mutable std::vector<MarkerParticle> _host; // mirror in the caller's order
  std::vector<MarkerParticle> _queued;       // added since the last sync, not yet on the device
  mutable bool _host_valid;                  // mirror == device set (+ queue)
  bool _host_edited;                         // mirror may have been written through an iterator
};

#endif
