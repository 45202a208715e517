// Shared ownership of one libfsb context (= one simulation domain resident in the HBM of one
// B200) for the host classes of this directory.  Not part of the reference's API: the reference
// keeps its state in host std::vectors; here MacGrid, MarkerParticleSet and FluidDomain are views
// of the same device context.  Every libfsb error becomes a std::runtime_error, the exception
// type the reference's step* functions throw (src/FluidSolver.cpp:101-107).
#ifndef FSB_DEVICE_CONTEXT_H
#define FSB_DEVICE_CONTEXT_H

#include <memory>
#include <stdexcept>
#include <string>

#include "../fsb.h"

namespace fsb {

class DeviceContext
{
public:
  DeviceContext(int size_x, int size_y, float length_x, float length_y, float density,
                float pic_ratio, int device)
  {
    const int rc =
        fsb_create(&_ctx, size_x, size_y, length_x, length_y, density, pic_ratio, device);
    if (rc != FSB_OK)
      throw std::runtime_error(std::string("fsb_create failed: ") + fsb_last_error(nullptr));
  }
  ~DeviceContext() { fsb_destroy(_ctx); }
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext& operator=(const DeviceContext&) = delete;

  fsb_ctx* get() const { return _ctx; }
  void check(int rc) const
  {
    if (rc != FSB_OK) throw std::runtime_error(fsb_last_error(_ctx));
  }

  // CUDA ordinal used by domains constructed without an explicit device (process-wide).
  static int& defaultDevice()
  {
    static int device = 0;
    return device;
  }

private:
  fsb_ctx* _ctx = nullptr;
};

typedef std::shared_ptr<DeviceContext> ContextPtr;

} // namespace fsb

#endif
