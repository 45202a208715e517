// MacGrid of the reference (include/MacGrid.h:19-177, src/MacGrid.cpp) as a VIEW of device state.
//
// The eight velocity grids and the label grid live in HBM inside the fsb context; this class
// keeps lazily filled host mirrors so that the reference's per-element accessors keep working
// for callers such as the renderer (src/Renderer.cpp:36-55,110-137):
//   * a getter downloads the one grid it needs the first time it is used after device work
//     (one cudaMemcpy2D of the whole grid, then plain array reads);
//   * a setter writes the mirror and marks it dirty; dirty mirrors are uploaded before the next
//     device stage (sync_to_device), and all mirrors are dropped after it (device_changed).
// Whole-grid stages (clearCellTypeBuffer, updatePreviousVelocityBuffer, updateVelocityDiffBuffer,
// swapVelocityBuffers) run on the device.
#ifndef FSB_MAC_GRID_H
#define FSB_MAC_GRID_H

#include <cstdint>
#include <vector>

#include "DeviceContext.h"
#include "Grid.h"

enum CellType
{
  LIQUID, // = FSB_LIQUID
  AIR,    // = FSB_AIR
  SOLID   // = FSB_SOLID
};

class MacGrid : public GridInterface
{
public:
  // Stand-alone grid with its own device context (include/MacGrid.h:22).
  MacGrid(int size_x, int size_y, MyFloat length_x, MyFloat length_y)
      : GridInterface(size_x, size_y, length_x / size_x, length_y / size_y),
        _dev(std::make_shared<fsb::DeviceContext>(size_x, size_y, length_x, length_y, 1.0f, 0.0f,
                                                   fsb::DeviceContext::defaultDevice()))
  {
    init_mirrors();
  }
  // View of an existing context (used by FluidDomain).
  MacGrid(fsb::ContextPtr dev, int size_x, int size_y, MyFloat length_x, MyFloat length_y)
      : GridInterface(size_x, size_y, length_x / size_x, length_y / size_y), _dev(dev)
  {
    init_mirrors();
  }

  // ---- whole-grid stages, on the device
  void clearCellTypeBuffer() // src/MacGrid.cpp:32-50
  {
    sync_to_device();
    _dev->check(fsb_clear_cell_types(_dev->get()));
    device_changed();
  }
  void updatePreviousVelocityBuffer() // src/MacGrid.cpp:52-56
  {
    sync_to_device();
    _dev->check(fsb_save_previous(_dev->get()));
    device_changed();
  }
  void updateVelocityDiffBuffer() // src/MacGrid.cpp:58-70
  {
    sync_to_device();
    _dev->check(fsb_update_diff(_dev->get()));
    device_changed();
  }
  void swapVelocityBuffers() // src/MacGrid.cpp:89-93
  {
    sync_to_device();
    _dev->check(fsb_swap_velocity_buffers(_dev->get()));
    device_changed();
  }

  // ---- getters (include/MacGrid.h:30-111)
  MyFloat velX(int i, int j) const { return (at(FSB_U_FRONT, i, j) + at(FSB_U_FRONT, i + 1, j)) / 2; }
  MyFloat velY(int i, int j) const { return (at(FSB_V_FRONT, i, j) + at(FSB_V_FRONT, i, j + 1)) / 2; }
  MyFloat velXHalfIndexed(int i, int j) const { return at(FSB_U_FRONT, i, j); }
  MyFloat velYHalfIndexed(int i, int j) const { return at(FSB_V_FRONT, i, j); }
  MyFloat velXBackBufferHalfIndexed(int i, int j) const { return at(FSB_U_BACK, i, j); }
  MyFloat velYBackBufferHalfIndexed(int i, int j) const { return at(FSB_V_BACK, i, j); }
  MyFloat velXBackBuffer(int i, int j) const { return (at(FSB_U_BACK, i, j) + at(FSB_U_BACK, i + 1, j)) / 2; }
  MyFloat velYBackBuffer(int i, int j) const { return (at(FSB_V_BACK, i, j) + at(FSB_V_BACK, i, j + 1)) / 2; }
  // point queries with the MAC half-cell shift (include/MacGrid.h:66-91); rounding as upstream:
  // the shift is formed in double and rounded once when passed on
  MyFloat velXInterpolated(MyFloat x, MyFloat y) const { return sample(FSB_U_FRONT, x, (MyFloat)(y - _DELTA_Y * 0.5)); }
  MyFloat velYInterpolated(MyFloat x, MyFloat y) const { return sample(FSB_V_FRONT, (MyFloat)(x - _DELTA_X * 0.5), y); }
  MyFloat velXDiffInterpolated(MyFloat x, MyFloat y) const { return sample(FSB_U_DIFF, x, (MyFloat)(y - _DELTA_Y * 0.5)); }
  MyFloat velYDiffInterpolated(MyFloat x, MyFloat y) const { return sample(FSB_V_DIFF, (MyFloat)(x - _DELTA_X * 0.5), y); }
  CellType cellType(int i, int j) const
  {
    i = (int)CLAMP(i, 0, _SIZE_X - 1);
    j = (int)CLAMP(j, 0, _SIZE_Y - 1);
    fetch_labels();
    return (CellType)_labels[twoDToLinear(i, j)];
  }
  MyFloat divVelX(int i, int j) const { return (at(FSB_U_FRONT, i + 1, j) - at(FSB_U_FRONT, i, j)) / _DELTA_X; }
  MyFloat divVelY(int i, int j) const { return (at(FSB_V_FRONT, i, j + 1) - at(FSB_V_FRONT, i, j)) / _DELTA_Y; }

  // ---- setters (include/MacGrid.h:116-157)
  void setVelXHalfIndexed(int i, int j, MyFloat vel_x) { ref(FSB_U_FRONT, i, j) = vel_x; }
  void setVelYHalfIndexed(int i, int j, MyFloat vel_y) { ref(FSB_V_FRONT, i, j) = vel_y; }
  void setVelXBackBuffer(int i, int j, MyFloat vel_x)
  {
    ref(FSB_U_BACK, i, j) = vel_x;
    ref(FSB_U_BACK, i + 1, j) = vel_x;
  }
  void setVelYBackBuffer(int i, int j, MyFloat vel_y)
  {
    ref(FSB_V_BACK, i, j) = vel_y;
    ref(FSB_V_BACK, i, j + 1) = vel_y;
  }
  void setVelXBackBufferHalfIndexed(int i, int j, MyFloat vel_x) { ref(FSB_U_BACK, i, j) = vel_x; }
  void setVelYBackBufferHalfIndexed(int i, int j, MyFloat vel_y) { ref(FSB_V_BACK, i, j) = vel_y; }
  void setCellType(int i, int j, CellType cell_type)
  {
    fetch_labels();
    _labels[twoDToLinear(i, j)] = (uint8_t)cell_type;
    _labels_dirty = true;
  }
  void addToVelXInterpolated(MyFloat x, MyFloat y, MyFloat vel_x) { splat(FSB_U_BACK, x, (MyFloat)(y - 0.5 * _DELTA_Y), vel_x); }
  void addToVelYInterpolated(MyFloat x, MyFloat y, MyFloat vel_y) { splat(FSB_V_BACK, (MyFloat)(x - 0.5 * _DELTA_X), y, vel_y); }

  // ---- bulk access (additions; not in the reference): whole grids without per-element calls
  const std::vector<MyFloat>& hostGrid(int which) const { fetch(which); return _grid[which]; }
  const std::vector<uint8_t>& hostCellTypes() const { fetch_labels(); return _labels; }

  // ---- device coherence (used by FluidDomain / FluidSolver / MarkerParticleSet)
  const fsb::ContextPtr& device() const { return _dev; }
  void sync_to_device()
  {
    for (int w = 0; w < 8; ++w)
      if (_dirty[w])
      {
        _dev->check(fsb_set_grid(_dev->get(), w, _grid[w].data()));
        _dirty[w] = false;
      }
    if (_labels_dirty)
    {
      _dev->check(fsb_set_cell_types(_dev->get(), _labels.data()));
      _labels_dirty = false;
    }
  }
  void device_changed()
  {
    for (int w = 0; w < 8; ++w) _valid[w] = false;
    _labels_valid = false;
  }

private:
  void init_mirrors()
  {
    for (int w = 0; w < 8; ++w) _valid[w] = _dirty[w] = false;
    _labels_valid = _labels_dirty = false;
  }
  void fetch(int which) const
  {
    if (_valid[which]) return;
    _grid[which].resize((size_t)_SIZE_X * _SIZE_Y);
    _dev->check(fsb_get_grid(_dev->get(), which, _grid[which].data()));
    _valid[which] = true;
  }
  void fetch_labels() const
  {
    if (_labels_valid) return;
    _labels.resize((size_t)_SIZE_X * _SIZE_Y);
    _dev->check(fsb_get_cell_types(_dev->get(), _labels.data()));
    _labels_valid = true;
  }
  MyFloat at(int which, int i, int j) const
  {
    fetch(which);
    return _grid[which][twoDToLinear(i, j)];
  }
  MyFloat& ref(int which, int i, int j)
  {
    fetch(which);
    _dirty[which] = true;
    return _grid[which][twoDToLinear(i, j)];
  }
  // include/Grid.h:117-144 on the mirror (an accessor for single points, not a simulation stage)
  MyFloat sample(int which, MyFloat x, MyFloat y) const
  {
    fetch(which);
    const std::vector<MyFloat>& g = _grid[which];
    const MyFloat xd = x / _DELTA_X, yd = y / _DELTA_Y;
    int i = (int)xd, j = (int)yd;
    const MyFloat fi = xd - (MyFloat)i, fj = yd - (MyFloat)j;
    i = (int)CLAMP(i, 0, _SIZE_X - 1);
    j = (int)CLAMP(j, 0, _SIZE_Y - 1);
    const int i1 = (int)CLAMP(i + 1, 0, _SIZE_X - 1), j1 = (int)CLAMP(j + 1, 0, _SIZE_Y - 1);
    const MyFloat v0 = (1 - fi) * g[i + (size_t)j * _SIZE_X] + fi * g[i1 + (size_t)j * _SIZE_X];
    const MyFloat v1 = (1 - fi) * g[i + (size_t)j1 * _SIZE_X] + fi * g[i1 + (size_t)j1 * _SIZE_X];
    return (1 - fj) * v0 + fj * v1;
  }
  // include/Grid.h:152-184 on the mirror
  void splat(int which, MyFloat x, MyFloat y, MyFloat value)
  {
    fetch(which);
    _dirty[which] = true;
    std::vector<MyFloat>& g = _grid[which];
    const MyFloat xd = x / _DELTA_X, yd = y / _DELTA_Y;
    int i = (int)xd, j = (int)yd;
    int i1 = i + 1, j1 = j + 1;
    const MyFloat fi = xd - (MyFloat)i, fj = yd - (MyFloat)j;
    i = (int)CLAMP(i, 0, _SIZE_X - 1);
    j = (int)CLAMP(j, 0, _SIZE_Y - 1);
    i1 = (int)CLAMP(i1, 0, _SIZE_X - 1);
    j1 = (int)CLAMP(j1, 0, _SIZE_Y - 1);
    const MyFloat v0 = (1 - fj) * value, v1 = fj * value;
    g[i + (size_t)j * _SIZE_X] += (1 - fi) * v0;
    g[i1 + (size_t)j * _SIZE_X] += fi * v0;
    g[i + (size_t)j1 * _SIZE_X] += (1 - fi) * v1;
    g[i1 + (size_t)j1 * _SIZE_X] += fi * v1;
  }

  fsb::ContextPtr _dev;
  mutable std::vector<MyFloat> _grid[8];
  mutable bool _valid[8];
  bool _dirty[8];
  mutable std::vector<uint8_t> _labels;
  mutable bool _labels_valid;
  bool _labels_dirty;
};

#endif
