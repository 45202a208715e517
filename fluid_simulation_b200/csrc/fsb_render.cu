// Device-side frame rasteriser: the frame examples/simple.cpp:73-82 draws through the reference's
// software renderer -- Renderer::clearCanvas, renderGridCellsToCanvas, renderParticlesToCanvas
// (src/Renderer.cpp:14-56,141-162) on src/Canvas.cpp's fillRectangle / drawPoint (:62-92), then the
// float -> byte conversion of writeCanvasToPpm (:217-248) -- produced from the state in HBM, so a
// run can be compared with the reference frame by frame without copying the particle set to the
// host.  Byte-identical to the reference's PPM payload.
//
// The reference paints cell rectangles in scan order (j outer, i inner), borders in the line
// colour (white), interiors in the cell's colour, later rectangles overwriting earlier ones; then
// one 3x3 square per particle.  Per pixel that is: the LAST rectangle that covers it decides, and
// since coverage is a product of a column range and a row range, the last one is (largest covering
// j, largest covering i).  The rectangle bounds are monotone in the cell index, so the largest
// covering index is found by bisection on the reference's own float expression.
#include <cstdio>
#include <vector>

#include "fsb_device.cuh"
#include "fsb_internal.cuh"

namespace {

struct RenderMap
{
  int W, H, nx, ny, ld;
  float scale_x, scale_y, translate_x, translate_y, cell_x, cell_y;
};

// static_cast<int>(-translate + k * cell_size), clamped like fillRectangle (src/Canvas.cpp:74-77)
__device__ __forceinline__ int rect_edge(float translate, float cell, int k, int hi)
{
  const int v = (int)(-translate + (float)k * cell);
  return min(max(v, 0), hi);
}

// largest k in [0, n-1] whose lower rectangle edge is <= pixel; -1 if none
__device__ __forceinline__ int last_covering(float translate, float cell, int n, int hi, int pixel)
{
  if (rect_edge(translate, cell, 0, hi) > pixel) return -1;
  int lo = 0, up = n - 1; // invariant: edge(lo) <= pixel
  while (lo < up)
  {
    const int mid = (lo + up + 1) >> 1;
    if (rect_edge(translate, cell, mid, hi) <= pixel) lo = mid;
    else up = mid - 1;
  }
  return lo;
}

__global__ void k_render_cells(uint8_t* __restrict__ rgb, const uint8_t* __restrict__ cell,
                               const RenderMap m)
{
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  if (px >= m.W || py >= m.H) return;
  uint8_t r = 255, g = 255, b = 255; // clearCanvas: white
  const int i = last_covering(m.translate_x, m.cell_x, m.nx, m.W - 1, px);
  const int j = last_covering(m.translate_y, m.cell_y, m.ny, m.H - 1, py);
  if (i >= 0 && j >= 0)
  {
    const int x0 = rect_edge(m.translate_x, m.cell_x, i, m.W - 1);
    const int x1 = rect_edge(m.translate_x, m.cell_x, i + 1, m.W - 1);
    const int y0 = rect_edge(m.translate_y, m.cell_y, j, m.H - 1);
    const int y1 = rect_edge(m.translate_y, m.cell_y, j + 1, m.H - 1);
    if (px <= x1 && py <= y1 && !(px == x0 || px == x1 || py == y0 || py == y1))
    {
      // interior: Color(0.7,0.7,1) / Color(1,1,1) / Color(0.5,0.5,0.5) through
      // static_cast<unsigned char>(CLAMP(c, 0, 1) * 255) in fp32
      const int t = cell[i + (size_t)j * m.ld];
      if (t == FSB_LIQUID) { r = 178; g = 178; b = 255; }
      else if (t == FSB_SOLID) { r = 127; g = 127; b = 127; }
    }
  }
  uint8_t* o = rgb + ((size_t)px + (size_t)py * m.W) * 3;
  o[0] = r; o[1] = g; o[2] = b;
}

// drawPoint(pos, 3): the 3x3 square, every edge clamped on its own (a particle outside the canvas
// still paints the border pixels it is clamped onto); all particles share one colour, so the
// byte stores are idempotent and need no ordering
__global__ void k_render_particles(uint8_t* __restrict__ rgb, const float4* __restrict__ part,
                                   int64_t n, const RenderMap m)
{
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float4 p = part[k];
  const int pos_x = (int)(-m.translate_x + m.scale_x * p.x);
  const int pos_y = (int)(-m.translate_y + m.scale_y * p.y);
  const int x0 = min(max(pos_x - 1, 0), m.W - 1), x1 = min(max(pos_x + 1, 0), m.W - 1);
  const int y0 = min(max(pos_y - 1, 0), m.H - 1), y1 = min(max(pos_y + 1, 0), m.H - 1);
  for (int y = y0; y <= y1; ++y)
    for (int x = x0; x <= x1; ++x)
    {
      uint8_t* o = rgb + ((size_t)x + (size_t)y * m.W) * 3;
      o[0] = 76; o[1] = 153; o[2] = 229; // Color(0.3, 0.6, 0.9)
    }
}

} // namespace

extern "C" {

int fsb_render_rgb(fsb_ctx* c, int width, int height, float x_min, float x_max, float y_min,
                   float y_max, uint8_t* rgb)
{
  if (!c) return fsb_fail(nullptr, FSB_ERR_INVALID, "null context");
  FSB_CUDA(c, cudaSetDevice(c->device));
  if (width < 1 || height < 1 || !rgb || !(x_max > x_min) || !(y_max > y_min))
    return fsb_fail(c, FSB_ERR_INVALID, "bad canvas (%dx%d) or area", width, height);
  RenderMap m;
  m.W = width; m.H = height; m.nx = c->nx; m.ny = c->ny; m.ld = c->ld;
  // src/Renderer.cpp:24-32 in the reference's types: int / float, 0.5 (double) * float -> float
  m.scale_x = width / (x_max - x_min);
  m.scale_y = height / (y_max - y_min);
  m.translate_x = (float)(0.5 * (x_min * width));
  m.translate_y = (float)(0.5 * (y_min * height));
  m.cell_x = c->dx * m.scale_x;
  m.cell_y = c->dy * m.scale_y;
  const size_t bytes = (size_t)width * height * 3;
  uint8_t* dev = nullptr;
  if (cudaMalloc((void**)&dev, bytes) != cudaSuccess)
    return fsb_fail(c, FSB_ERR_NOMEM, "cudaMalloc of %zu bytes failed", bytes);
  k_render_cells<<<dim3(fsb_div_up(width, 128), height), 128, 0, c->stream>>>(dev, c->cell, m);
  c->launches++;
  if (c->n > 0)
  {
    k_render_particles<<<fsb_div_up(c->n, 256), 256, 0, c->stream>>>(dev, c->part[c->pcur], c->n, m);
    c->launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(rgb, dev, bytes, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(dev);
  if (e != cudaSuccess) return fsb_fail(c, FSB_ERR_CUDA, "frame rendering failed: %s", cudaGetErrorString(e));
  return FSB_OK;
}

int fsb_write_ppm(fsb_ctx* c, const char* path, int width, int height, float x_min, float x_max,
                  float y_min, float y_max)
{
  if (!c) return fsb_fail(nullptr, FSB_ERR_INVALID, "null context");
  if (!path) return fsb_fail(c, FSB_ERR_INVALID, "null path");
  if (width < 1 || height < 1) return fsb_fail(c, FSB_ERR_INVALID, "bad canvas (%dx%d)", width, height);
  std::vector<uint8_t> rgb((size_t)width * height * 3);
  FSB_TRY(fsb_render_rgb(c, width, height, x_min, x_max, y_min, y_max, rgb.data()));
  FILE* f = fopen(path, "wb");
  if (!f) return fsb_fail(c, FSB_ERR_INVALID, "cannot open %s for writing", path);
  fprintf(f, "P6\n%d %d\n255\n", width, height); // src/Renderer.cpp:244
  const bool ok = fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
  if (fclose(f) != 0 || !ok) return fsb_fail(c, FSB_ERR_INVALID, "short write to %s", path);
  return FSB_OK;
}

} // extern "C"
