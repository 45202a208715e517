/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * Plain-C restatement of the per-step hot path of kbladin/Fluid_Simulation
 * (reference commit 78af5c4).  Every function cites the reference file:line it
 * follows (paths relative to the reference tree).  The arithmetic is restated
 * literally -- including where the reference promotes to double because of an
 * unsuffixed literal and rounds back to float -- so that this file is
 * BIT-IDENTICAL to oracle/_ref/libfsref.so (the reference's own sources
 * compiled unchanged); tests/test_oracle.py asserts that, stage by stage and
 * over whole runs, and tests/golden/ holds vectors generated from _ref.
 *
 * Parity status: every stage except the linear solve is pinned against the
 * reference's own compiled code.  The conjugate-gradient solve restates
 * Eigen's ConjugateGradient (a dependency the reference neither vendors nor
 * pins and that is absent here) -- that part is
 * "parity unpinned" against real Eigen and says so in DESIGN.md.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (never -march=native / -ffast-math).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FSX(name) fso_##name
#include "fluid_oracle_api.h"

enum { LIQUID = 0, AIR = 1, SOLID = 2 }; /* include/MacGrid.h:14-17 */

typedef struct
{
  int nx, ny;
  float dx, dy;   /* MacGrid deltas: length/size in float (src/MacGrid.cpp:8) */
  float density, pic_ratio;
  /* MacGrid state (include/MacGrid.h:164-176) */
  float* u[2];
  float* v[2];
  int front; /* u[front] is the front buffer; swapVelocityBuffers flips it */
  float *u_prev, *v_prev, *u_diff, *v_diff;
  uint8_t* cell;
  /* FluidSolverMemoryPool (include/FluidSolver.h:28-42) */
  int* fluid_idx;
  uint8_t* n_part;
  uint8_t* mask_x[2];
  uint8_t* mask_y[2];
  int mask_front;
  float *sum_x, *sum_y, *w_x, *w_y;
  float pool_dx, pool_dy; /* copy-ctor passes deltaX twice (src/FluidSolver.cpp:56-65) */
  /* particles, AoS {px,py,vx,vy} (include/MarkerParticleSet.h:43-50) */
  float* part;
  int64_t n, cap;
  /* solver */
  int max_iters; /* src/FluidSolver.cpp:81 */
  float tol;     /* Eigen default: NumTraits<float>::epsilon() */
  int iters;
  float err;
  float* pressure; /* last CG solution scattered on the full grid */
  int integrator;  /* 0 = RK3 (include/FluidSolver.h:144), 1 = EulerExplicit */
} Ctx;

/* include/MathDefinitions.h:16-19 -- clamp THROUGH float, also for ints */
static inline float CLAMPf(float d, float mn, float mx)
{
  const float t = d < mn ? mn : d;
  return t > mx ? mx : t;
}
static inline int CLAMPi(int d, int mn, int mx)
{
  return (int)CLAMPf((float)d, (float)mn, (float)mx);
}

/* include/Grid.h:117-144  Grid<T>::valueInterpolated */
static float grid_interp(const float* g, int nx, int ny, float dx, float dy,
                         float x, float y)
{
  int i = (int)(x / dx);
  int j = (int)(y / dy);
  float i_frac = x / dx - i;
  float j_frac = y / dy - j;
  i = CLAMPi(i, 0, nx - 1);
  j = CLAMPi(j, 0, ny - 1);
  int i1 = CLAMPi(i + 1, 0, nx - 1);
  int j1 = CLAMPi(j + 1, 0, ny - 1);
  float v00 = g[i + (size_t)j * nx];
  float v10 = g[i1 + (size_t)j * nx];
  float v01 = g[i + (size_t)j1 * nx];
  float v11 = g[i1 + (size_t)j1 * nx];
  float v0 = (1 - i_frac) * v00 + i_frac * v10;
  float v1 = (1 - i_frac) * v01 + i_frac * v11;
  return (1 - j_frac) * v0 + j_frac * v1;
}

/* include/Grid.h:152-184  Grid<T>::addToValueInterpolated */
static void grid_splat(float* g, int nx, int ny, float dx, float dy, float x,
                       float y, float value)
{
  int i = (int)(x / dx);
  int j = (int)(y / dy);
  int i1 = i + 1;
  int j1 = j + 1;
  float i_frac = x / dx - i;
  float j_frac = y / dy - j;
  i = CLAMPi(i, 0, nx - 1);
  j = CLAMPi(j, 0, ny - 1);
  i1 = CLAMPi(i1, 0, nx - 1);
  j1 = CLAMPi(j1, 0, ny - 1);
  float v0 = (1 - j_frac) * value;
  float v1 = j_frac * value;
  float v00 = (1 - i_frac) * v0;
  float v10 = i_frac * v0;
  float v01 = (1 - i_frac) * v1;
  float v11 = i_frac * v1;
  g[i + (size_t)j * nx] += v00;
  g[i1 + (size_t)j * nx] += v10;
  g[i + (size_t)j1 * nx] += v01;
  g[i1 + (size_t)j1 * nx] += v11;
}

/* include/MacGrid.h:66-91 -- the MAC half-cell shift is computed in double
 * (`_DELTA_Y * 0.5`) and rounded to float when passed on. */
static float vel_x_interp(const Ctx* c, const float* u, float x, float y)
{
  return grid_interp(u, c->nx, c->ny, c->dx, c->dy, x, (float)(y - c->dy * 0.5));
}
static float vel_y_interp(const Ctx* c, const float* v, float x, float y)
{
  return grid_interp(v, c->nx, c->ny, c->dx, c->dy, (float)(x - c->dx * 0.5), y);
}
/* include/MacGrid.h:92-97 */
static int cell_type(const Ctx* c, int i, int j)
{
  i = CLAMPi(i, 0, c->nx - 1);
  j = CLAMPi(j, 0, c->ny - 1);
  return c->cell[i + (size_t)j * c->nx];
}
static void swap_velocity(Ctx* c) { c->front ^= 1; } /* src/MacGrid.cpp:89-93 */

#define UF(c) ((c)->u[(c)->front])
#define VF(c) ((c)->v[(c)->front])
#define UB(c) ((c)->u[(c)->front ^ 1])
#define VB(c) ((c)->v[(c)->front ^ 1])
#define AT(c, i, j) ((size_t)(i) + (size_t)(j) * (size_t)(c)->nx)

/* ---------------------------------------------------------------------- */

static void* zalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

void* fso_create(int size_x, int size_y, float length_x, float length_y,
                 float density, float pic_ratio)
{
  Ctx* c = (Ctx*)calloc(1, sizeof(Ctx));
  const size_t C = (size_t)size_x * (size_t)size_y;
  c->nx = size_x;
  c->ny = size_y;
  c->dx = length_x / size_x; /* src/MacGrid.cpp:8, src/FluidDomain.cpp:61 */
  c->dy = length_y / size_y;
  c->density = density;
  c->pic_ratio = pic_ratio;
  for (int k = 0; k < 2; ++k)
  {
    c->u[k] = (float*)zalloc(C, 4);
    c->v[k] = (float*)zalloc(C, 4);
    c->mask_x[k] = (uint8_t*)zalloc(C, 1);
    c->mask_y[k] = (uint8_t*)zalloc(C, 1);
  }
  c->u_prev = (float*)zalloc(C, 4);
  c->v_prev = (float*)zalloc(C, 4);
  c->u_diff = (float*)zalloc(C, 4);
  c->v_diff = (float*)zalloc(C, 4);
  c->cell = (uint8_t*)zalloc(C, 1);
  c->fluid_idx = (int*)zalloc(C, 4);
  c->n_part = (uint8_t*)zalloc(C, 1);
  c->sum_x = (float*)zalloc(C, 4);
  c->sum_y = (float*)zalloc(C, 4);
  c->w_x = (float*)zalloc(C, 4);
  c->w_y = (float*)zalloc(C, 4);
  c->pressure = (float*)zalloc(C, 4);
  c->pool_dx = c->dx; /* src/FluidSolver.cpp:28-30 then :56-65 */
  c->pool_dy = c->dx;
  c->max_iters = 100; /* src/FluidSolver.cpp:81 */
  c->tol = FLT_EPSILON;
  /* src/MacGrid.cpp:24,32-50  clearCellTypeBuffer in the ctor */
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i) c->cell[AT(c, i, j)] = SOLID;
  for (int j = 1; j < c->ny - 1; ++j)
    for (int i = 1; i < c->nx - 1; ++i) c->cell[AT(c, i, j)] = AIR;
  return c;
}

void fso_destroy(void* h)
{
  Ctx* c = (Ctx*)h;
  if (!c) return;
  for (int k = 0; k < 2; ++k)
  {
    free(c->u[k]);
    free(c->v[k]);
    free(c->mask_x[k]);
    free(c->mask_y[k]);
  }
  free(c->u_prev); free(c->v_prev); free(c->u_diff); free(c->v_diff);
  free(c->cell); free(c->fluid_idx); free(c->n_part);
  free(c->sum_x); free(c->sum_y); free(c->w_x); free(c->w_y);
  free(c->pressure); free(c->part);
  free(c);
}

float fso_delta_x(void* h) { return ((Ctx*)h)->dx; }
float fso_delta_y(void* h) { return ((Ctx*)h)->dy; }
void fso_set_cg(void* h, int max_iters, float tol)
{
  ((Ctx*)h)->max_iters = max_iters;
  ((Ctx*)h)->tol = tol;
}
/* not in the shared API: 0 = RK3 (reference default), 1 = EulerExplicit */
void fso_set_integrator(void* h, int integrator) { ((Ctx*)h)->integrator = integrator; }

void fso_append_particles(void* h, const float* a, int64_t n)
{
  Ctx* c = (Ctx*)h;
  if (c->n + n > c->cap)
  {
    c->cap = (c->n + n) * 2;
    c->part = (float*)realloc(c->part, (size_t)c->cap * 16);
  }
  memcpy(c->part + 4 * c->n, a, (size_t)n * 16);
  c->n += n;
}
void fso_set_particles(void* h, const float* a, int64_t n)
{
  ((Ctx*)h)->n = 0;
  fso_append_particles(h, a, n);
}
int64_t fso_num_particles(void* h) { return ((Ctx*)h)->n; }
void fso_get_particles(void* h, float* a)
{
  Ctx* c = (Ctx*)h;
  memcpy(a, c->part, (size_t)c->n * 16);
}

/* src/FluidDomain.cpp:29-52  FluidSource::update, the spawning branch */
int64_t fso_emit_source(void* h, float x_min, float x_max, float y_min, float y_max,
                        float delta_x, float delta_y, float vel_x, float vel_y)
{
  Ctx* c = (Ctx*)h;
  const int64_t before = c->n;
  float x_incr = (float)(delta_x / 2.5);
  float y_incr = (float)(delta_y / 2.5);
  for (float y = y_min; y < y_max; y += y_incr)
    for (float x = x_min; x < x_max; x += x_incr)
    {
      float p[4] = {x, y, vel_x, vel_y};
      fso_append_particles(h, p, 1);
    }
  return c->n - before;
}

static float* pick_grid(Ctx* c, int which)
{
  switch (which)
  {
  case FSX_U_FRONT: return UF(c);
  case FSX_V_FRONT: return VF(c);
  case FSX_U_BACK: return UB(c);
  case FSX_V_BACK: return VB(c);
  case FSX_U_PREV: return c->u_prev;
  case FSX_V_PREV: return c->v_prev;
  case FSX_U_DIFF: return c->u_diff;
  case FSX_V_DIFF: return c->v_diff;
  }
  return 0;
}
void fso_set_grid(void* h, int which, const float* src)
{
  Ctx* c = (Ctx*)h;
  memcpy(pick_grid(c, which), src, (size_t)c->nx * c->ny * 4);
}
void fso_get_grid(void* h, int which, float* dst)
{
  Ctx* c = (Ctx*)h;
  memcpy(dst, pick_grid(c, which), (size_t)c->nx * c->ny * 4);
}
void fso_set_cell_types(void* h, const uint8_t* src)
{
  Ctx* c = (Ctx*)h;
  memcpy(c->cell, src, (size_t)c->nx * c->ny);
}
void fso_get_cell_types(void* h, uint8_t* dst)
{
  Ctx* c = (Ctx*)h;
  memcpy(dst, c->cell, (size_t)c->nx * c->ny);
}

/* ---------------------------------------------------------------------- */

/* src/FluidDomain.cpp:150-180 (+ src/MacGrid.cpp:32-50) */
void fso_classify_cells(void* h)
{
  Ctx* c = (Ctx*)h;
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i) c->cell[AT(c, i, j)] = SOLID;
  for (int j = 1; j < c->ny - 1; ++j)
    for (int i = 1; i < c->nx - 1; ++i) c->cell[AT(c, i, j)] = AIR;
  /* lengthX() = _SIZE_X * _DELTA_X in float (include/Grid.h:54-55) */
  const float len_x = c->nx * c->dx;
  const float len_y = c->ny * c->dy;
  for (int64_t k = 0; k < c->n; ++k)
  {
    int x = (int)((c->part[4 * k] / len_x) * c->nx);
    int y = (int)((c->part[4 * k + 1] / len_y) * c->ny);
    x = CLAMPi(x, 0, c->nx - 1);
    y = CLAMPi(y, 0, c->ny - 1);
    c->cell[AT(c, x, y)] = LIQUID;
  }
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
      if (i == 0 || j == 0 || i == c->nx - 1 || j == c->ny - 1)
        c->cell[AT(c, i, j)] = SOLID;
}

/* src/FluidSolver.cpp:873-919  transferVelocityToGridSpread
 * The accumulators carry the POOL's deltas (src/FluidSolver.cpp:13-16). */
void fso_p2g_spread(void* h)
{
  Ctx* c = (Ctx*)h;
  const size_t C = (size_t)c->nx * c->ny;
  memset(c->sum_x, 0, C * 4);
  memset(c->sum_y, 0, C * 4);
  memset(c->w_x, 0, C * 4);
  memset(c->w_y, 0, C * 4);
  for (int64_t k = 0; k < c->n; ++k)
  {
    const float px = c->part[4 * k], py = c->part[4 * k + 1];
    const float vx = c->part[4 * k + 2], vy = c->part[4 * k + 3];
    const float ys = (float)(py - 0.5 * c->dy); /* mac_grid.deltaY(), :890 */
    const float xs = (float)(px - 0.5 * c->dx);
    grid_splat(c->sum_x, c->nx, c->ny, c->pool_dx, c->pool_dy, px, ys, vx);
    grid_splat(c->sum_y, c->nx, c->ny, c->pool_dx, c->pool_dy, xs, py, vy);
    grid_splat(c->w_x, c->nx, c->ny, c->pool_dx, c->pool_dy, px, ys, 1.0f);
    grid_splat(c->w_y, c->nx, c->ny, c->pool_dx, c->pool_dy, xs, py, 1.0f);
  }
  float* ub = UB(c);
  float* vb = VB(c);
  for (size_t k = 0; k < C; ++k)
  {
    if (c->w_x[k] > 0.000001) ub[k] = c->sum_x[k] / c->w_x[k];
    if (c->w_y[k] > 0.000001) vb[k] = c->sum_y[k] / c->w_y[k];
  }
  swap_velocity(c);
}

/* src/MacGrid.cpp:52-56 */
void fso_save_previous(void* h)
{
  Ctx* c = (Ctx*)h;
  const size_t C = (size_t)c->nx * c->ny;
  memcpy(c->u_prev, UF(c), C * 4);
  memcpy(c->v_prev, VF(c), C * 4);
}

/* src/MacGrid.cpp:58-70 */
void fso_update_diff(void* h)
{
  Ctx* c = (Ctx*)h;
  const size_t C = (size_t)c->nx * c->ny;
  const float *uf = UF(c), *vf = VF(c);
  for (size_t k = 0; k < C; ++k)
  {
    c->u_diff[k] = uf[k] - c->u_prev[k];
    c->v_diff[k] = vf[k] - c->v_prev[k];
  }
}

/* src/FluidSolver.cpp:276-295 */
void fso_add_acceleration(void* h, float ax, float ay, float dt)
{
  Ctx* c = (Ctx*)h;
  float *uf = UF(c), *vf = VF(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
      if (cell_type(c, i, j) == LIQUID)
      {
        uf[AT(c, i, j)] = uf[AT(c, i, j)] + ax * dt;
        vf[AT(c, i, j)] = vf[AT(c, i, j)] + ay * dt;
      }
}

/* src/FluidSolver.cpp:297-321 */
void fso_enforce_dirichlet(void* h)
{
  Ctx* c = (Ctx*)h;
  float *uf = UF(c), *vf = VF(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      int im1 = CLAMPi(i - 1, 0, c->nx - 1);
      int jm1 = CLAMPi(j - 1, 0, c->ny - 1);
      const size_t k = AT(c, i, j);
      if ((cell_type(c, im1, j) == SOLID && uf[k] < 0) ||
          (cell_type(c, i, j) == SOLID && uf[k] > 0))
        uf[k] = 0;
      if ((cell_type(c, i, jm1) == SOLID && vf[k] < 0) ||
          (cell_type(c, i, j) == SOLID && vf[k] > 0))
        vf[k] = 0;
    }
}

/* src/FluidSolver.cpp:485-622  extendVelocityIndividual (incl. the :527 typo) */
void fso_extend_velocity(void* h, int n_iter)
{
  Ctx* c = (Ctx*)h;
  float *uf = UF(c), *vf = VF(c), *ub = UB(c), *vb = VB(c);
  (void)vf;
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      const size_t k = AT(c, i, j);
      uint8_t* mxf = c->mask_x[c->mask_front];
      uint8_t* mxb = c->mask_x[c->mask_front ^ 1];
      uint8_t* myf = c->mask_y[c->mask_front];
      uint8_t* myb = c->mask_y[c->mask_front ^ 1];
      if (cell_type(c, i, j) == LIQUID || cell_type(c, i - 1, j) == LIQUID)
      {
        mxf[k] = 1;
        mxb[k] = 1;
        ub[k] = uf[k];
      }
      else
      {
        mxf[k] = 0;
        mxb[k] = 0;
        ub[k] = 0;
        uf[k] = 0;
      }
      if (cell_type(c, i, j) == LIQUID || cell_type(c, i, j - 1) == LIQUID)
      {
        myf[k] = 1;
        myb[k] = 1;
        vb[k] = VF(c)[k];
      }
      else
      {
        myf[k] = 0;
        myb[k] = 0;
        vb[k] = 0;
        uf[k] = 0; /* sic: setVelXHalfIndexed at src/FluidSolver.cpp:527 */
      }
    }
  for (int iter = 0; iter < n_iter; ++iter)
  {
    const uint8_t* mxf = c->mask_x[c->mask_front];
    uint8_t* mxb = c->mask_x[c->mask_front ^ 1];
    const uint8_t* myf = c->mask_y[c->mask_front];
    uint8_t* myb = c->mask_y[c->mask_front ^ 1];
    for (int j = 0; j < c->ny; ++j)
      for (int i = 0; i < c->nx; ++i)
      {
        /* The reference indexes (i-1,j) etc. unclamped; the conditions below
         * guarantee 1 <= i,j <= size-2 whenever a neighbour is read (the
         * border is always SOLID), so plain indexing is the same access. */
        const size_t k = AT(c, i, j);
        if (mxf[k] == 0 && cell_type(c, i, j) != SOLID &&
            cell_type(c, i - 1, j) != SOLID)
        {
          float nv = 0;
          int n = 0;
          if (mxf[AT(c, i - 1, j)] == 1) { nv += ub[AT(c, i - 1, j)]; n++; }
          if (mxf[AT(c, i, j - 1)] == 1) { nv += ub[AT(c, i, j - 1)]; n++; }
          if (mxf[AT(c, i, j + 1)] == 1) { nv += ub[AT(c, i, j + 1)]; n++; }
          if (mxf[AT(c, i + 1, j)] == 1) { nv += ub[AT(c, i + 1, j)]; n++; }
          if (n > 0)
          {
            nv /= n;
            ub[k] = nv;
            mxb[k] = 1;
          }
        }
        if (myf[k] == 0 && cell_type(c, i, j) != SOLID &&
            cell_type(c, i, j - 1) != SOLID)
        {
          float nv = 0;
          int n = 0;
          if (myf[AT(c, i - 1, j)] == 1) { nv += vb[AT(c, i - 1, j)]; n++; }
          if (myf[AT(c, i, j - 1)] == 1) { nv += vb[AT(c, i, j - 1)]; n++; }
          if (myf[AT(c, i, j + 1)] == 1) { nv += vb[AT(c, i, j + 1)]; n++; }
          if (myf[AT(c, i + 1, j)] == 1) { nv += vb[AT(c, i + 1, j)]; n++; }
          if (n > 0)
          {
            nv /= n;
            vb[k] = nv;
            myb[k] = 1;
          }
        }
      }
    c->mask_front ^= 1; /* swapValidMaskBuffer, :619 */
  }
  swap_velocity(c);
}

/* src/FluidSolver.cpp:253-274  addExternalForce: F / density * dt on the left and bottom faces
 * of LIQUID cells (not called by any step) */
void fso_add_external_force(void* h, float fx, float fy, float dt)
{
  Ctx* c = (Ctx*)h;
  float *uf = UF(c), *vf = VF(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
      if (cell_type(c, i, j) == LIQUID)
      {
        const size_t k = AT(c, i, j);
        uf[k] = uf[k] + fx / c->density * dt;
        vf[k] = vf[k] + fy / c->density * dt;
      }
}

/* src/FluidSolver.cpp:625-707  extendVelocityAvarageing (not called by any step).  One validity mask
 * per CELL (the x masks of the pool); a non-SOLID cell without the mask takes the mean of the
 * cell-centred back-buffer velocities ((face + next face) / 2, include/MacGrid.h:56-65) of its masked
 * neighbours, visited in the order (i-1,j), (i,j-1), (i,j+1), (i+1,j), and writes it to BOTH of its
 * faces per component (:124-133) -- so a sweep depends on the scan order (row-major), which this
 * restatement keeps.  Masks are swapped after every sweep, velocities once at the end. */
void fso_extend_velocity_avg(void* h, int n_iterations)
{
  Ctx* c = (Ctx*)h;
  float *uf = UF(c), *vf = VF(c), *ub = UB(c), *vb = VB(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      const uint8_t m = cell_type(c, i, j) == LIQUID ? 1 : 0;
      c->mask_x[c->mask_front][AT(c, i, j)] = m;
      c->mask_x[c->mask_front ^ 1][AT(c, i, j)] = m;
      ub[AT(c, i, j)] = uf[AT(c, i, j)];
      vb[AT(c, i, j)] = vf[AT(c, i, j)];
    }
#define UBC(i, j) ((ub[AT(c, i, j)] + ub[AT(c, (i) + 1, j)]) / 2)
#define VBC(i, j) ((vb[AT(c, i, j)] + vb[AT(c, i, (j) + 1)]) / 2)
  for (int iter = 0; iter < n_iterations; ++iter)
  {
    const uint8_t* mf = c->mask_x[c->mask_front];
    uint8_t* mb = c->mask_x[c->mask_front ^ 1];
    for (int j = 0; j < c->ny; ++j)
      for (int i = 0; i < c->nx; ++i)
        if (mf[AT(c, i, j)] == 0 && cell_type(c, i, j) != SOLID)
        {
          float nvx = 0, nvy = 0;
          int n = 0;
          if (mf[AT(c, i - 1, j)] == 1) { nvx += UBC(i - 1, j); nvy += VBC(i - 1, j); n++; }
          if (mf[AT(c, i, j - 1)] == 1) { nvx += UBC(i, j - 1); nvy += VBC(i, j - 1); n++; }
          if (mf[AT(c, i, j + 1)] == 1) { nvx += UBC(i, j + 1); nvy += VBC(i, j + 1); n++; }
          if (mf[AT(c, i + 1, j)] == 1) { nvx += UBC(i + 1, j); nvy += VBC(i + 1, j); n++; }
          if (n > 0)
          {
            nvx /= n;
            nvy /= n;
            ub[AT(c, i, j)] = nvx; ub[AT(c, i + 1, j)] = nvx;
            vb[AT(c, i, j)] = nvy; vb[AT(c, i, j + 1)] = nvy;
            mb[AT(c, i, j)] = 1;
          }
        }
    c->mask_front ^= 1;
  }
#undef UBC
#undef VBC
  c->front ^= 1;
}

/* src/FluidSolver.cpp:816-871  transferVelocityToGridGather: for every face the particles whose
 * hat weight 1 - (|dx|/deltaX + |dy|/deltaY) is >= 1, i.e. (numerically) ON the face position;
 * mean of their velocities, written to the back buffer, swap (not called by any step) */
void fso_p2g_gather(void* h)
{
  Ctx* c = (Ctx*)h;
  float *ub = UB(c), *vb = VB(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      const float x_u = i * c->dx;
      const float y_u = (float)((j + 0.5) * c->dy);
      const float x_v = (float)((i + 0.5) * c->dx);
      const float y_v = j * c->dy;
      float wx = 0, sx = 0, wy = 0, sy = 0;
      for (int64_t q = 0; q < c->n; ++q)
      {
        const float* p = c->part + 4 * q;
        float ax = fabsf(p[0] - x_u);
        float ay = fabsf(p[1] - y_u);
        const float w_u = 1 - (ax / c->dx + ay / c->dy);
        if (w_u >= 1)
        {
          wx += w_u;
          sx += p[2];
        }
        ax = fabsf(p[0] - x_v);
        ay = fabsf(p[1] - y_v);
        const float w_v = 1 - (ax / c->dx + ay / c->dy);
        if (w_v >= 1)
        {
          wy += w_v;
          sy += p[3];
        }
      }
      if (wx)
      {
        sx /= wx;
        ub[AT(c, i, j)] = sx;
      }
      if (wy)
      {
        sy /= wy;
        vb[AT(c, i, j)] = sy;
      }
    }
  swap_velocity(c);
}

/* examples/simple.cpp:73-82 frame: src/Renderer.cpp clearCanvas (:14-17), renderGridCellsToCanvas
 * (:19-56), renderParticlesToCanvas (:141-162) with src/Canvas.cpp fillRectangle / drawPoint
 * (:62-92), and the float -> byte conversion of writeCanvasToPpm (:217-248). */
typedef struct { float r, g, b; } Rgb;
static void canvas_fill_rect(Rgb* px, int W, int H, int x0, int x1, int y0, int y1, Rgb line, Rgb fill)
{
  x0 = (int)CLAMPf((float)x0, 0, (float)(W - 1));
  x1 = (int)CLAMPf((float)x1, 0, (float)(W - 1));
  y0 = (int)CLAMPf((float)y0, 0, (float)(H - 1));
  y1 = (int)CLAMPf((float)y1, 0, (float)(H - 1));
  for (int j = y0; j <= y1; ++j)
    for (int i = x0; i <= x1; ++i)
      px[i + (size_t)j * W] = (i == x0 || i == x1 || j == y0 || j == y1) ? line : fill;
}
void fso_render_rgb(void* h, int W, int H, float x_min, float x_max, float y_min, float y_max,
                    uint8_t* rgb)
{
  Ctx* c = (Ctx*)h;
  Rgb* px = (Rgb*)malloc(sizeof(Rgb) * (size_t)W * H);
  const Rgb white = {1, 1, 1};
  for (size_t k = 0; k < (size_t)W * H; ++k) px[k] = white;
  const float scale_x = W / (x_max - x_min);
  const float scale_y = H / (y_max - y_min);
  const float translate_x = (float)(0.5 * (x_min * W));
  const float translate_y = (float)(0.5 * (y_min * H));
  const float cell_x = c->dx * scale_x;
  const float cell_y = c->dy * scale_y;
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      const int t = cell_type(c, i, j);
      Rgb fill;
      if (t == LIQUID) { fill.r = (float)0.7; fill.g = (float)0.7; fill.b = 1; }
      else if (t == AIR) { fill = white; }
      else { fill.r = fill.g = fill.b = (float)0.5; }
      canvas_fill_rect(px, W, H, (int)(-translate_x + i * cell_x), (int)(-translate_x + (i + 1) * cell_x),
                       (int)(-translate_y + j * cell_y), (int)(-translate_y + (j + 1) * cell_y), white,
                       fill);
    }
  Rgb blue;
  blue.r = (float)0.3; blue.g = (float)0.6; blue.b = (float)0.9;
  for (int64_t q = 0; q < c->n; ++q)
  {
    const int pos_x = (int)(-translate_x + scale_x * c->part[4 * q]);
    const int pos_y = (int)(-translate_y + scale_y * c->part[4 * q + 1]);
    canvas_fill_rect(px, W, H, pos_x - 3 / 2, pos_x + 3 / 2, pos_y - 3 / 2, pos_y + 3 / 2, blue, blue);
  }
  for (size_t k = 0; k < (size_t)W * H; ++k)
  {
    rgb[3 * k + 0] = (unsigned char)(CLAMPf(px[k].r, 0, 1) * 255);
    rgb[3 * k + 1] = (unsigned char)(CLAMPf(px[k].g, 0, 1) * 255);
    rgb[3 * k + 2] = (unsigned char)(CLAMPf(px[k].b, 0, 1) * 255);
  }
  free(px);
}

/* Eigen ConjugateGradient<SparseMatrix<float>, Lower, DiagonalPreconditioner>
 * restated matrix-free on the compact liquid numbering, in the same operation
 * order as oracle/eigen_shim/Eigen/IterativeLinearSolvers (column sweep over
 * the lower triangle: diagonal, +x neighbour, +y neighbour). */
static float dotd(const float* a, const float* b, int n)
{
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += (double)a[i] * (double)b[i];
  return (float)s;
}

typedef struct
{
  int n;
  int* east;  /* compact index of the +x liquid neighbour or -1 */
  int* north; /* compact index of the +y liquid neighbour or -1 */
  float* diag;
  float off;
} LowerLaplacian;

static void lower_product(const LowerLaplacian* A, const float* v, float* res)
{
  memset(res, 0, (size_t)A->n * 4);
  for (int j = 0; j < A->n; ++j)
  {
    res[j] += A->diag[j] * v[j];
    const float v_j = v[j];
    float res_j = 0;
    if (A->east[j] >= 0)
    {
      res_j += A->off * v[A->east[j]];
      res[A->east[j]] += A->off * v_j;
    }
    if (A->north[j] >= 0)
    {
      res_j += A->off * v[A->north[j]];
      res[A->north[j]] += A->off * v_j;
    }
    res[j] += res_j;
  }
}

static void cg_solve(Ctx* c, const LowerLaplacian* A, const float* b, float* x)
{
  const int n = A->n;
  float* r = (float*)malloc((size_t)n * 4);
  float* p = (float*)malloc((size_t)n * 4);
  float* z = (float*)malloc((size_t)n * 4);
  float* tmp = (float*)malloc((size_t)n * 4);
  float* invdiag = (float*)malloc((size_t)n * 4);
  for (int i = 0; i < n; ++i)
    invdiag[i] = (A->diag[i] != 0.0f) ? 1.0f / A->diag[i] : 1.0f;
  const int max_iters = c->max_iters < 0 ? 2 * n : c->max_iters;
  memset(x, 0, (size_t)n * 4);
  memcpy(r, b, (size_t)n * 4);
  const float rhs2 = dotd(b, b, n);
  float r2 = 0;
  int it = 0;
  if (rhs2 == 0.0f)
  {
    c->iters = 0;
    c->err = 0;
    goto done;
  }
  {
    float thr = c->tol * c->tol * rhs2;
    if (thr < FLT_MIN) thr = FLT_MIN;
    r2 = dotd(r, r, n);
    if (r2 < thr)
    {
      c->iters = 0;
      c->err = sqrtf(r2 / rhs2);
      goto done;
    }
    for (int i = 0; i < n; ++i) p[i] = invdiag[i] * r[i];
    float abs_new = dotd(r, p, n);
    while (it < max_iters)
    {
      lower_product(A, p, tmp);
      const float alpha = abs_new / dotd(p, tmp, n);
      for (int i = 0; i < n; ++i) x[i] = x[i] + alpha * p[i];
      for (int i = 0; i < n; ++i) r[i] = r[i] - alpha * tmp[i];
      r2 = dotd(r, r, n);
      if (r2 < thr) break;
      for (int i = 0; i < n; ++i) z[i] = invdiag[i] * r[i];
      const float abs_old = abs_new;
      abs_new = dotd(r, z, n);
      const float beta = abs_new / abs_old;
      for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
      ++it;
    }
    c->err = sqrtf(r2 / rhs2);
    c->iters = it;
  }
done:
  free(r); free(p); free(z); free(tmp); free(invdiag);
}

/* src/FluidSolver.cpp:323-483  pressureSolve */
void fso_pressure_solve(void* h, float density, float dt)
{
  Ctx* c = (Ctx*)h;
  const size_t C = (size_t)c->nx * c->ny;
  int n_fluid = 0;
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      if (cell_type(c, i, j) == LIQUID) c->fluid_idx[AT(c, i, j)] = n_fluid++;
      else c->fluid_idx[AT(c, i, j)] = -1;
      c->n_part[AT(c, i, j)] = 0;
    }
  if (n_fluid == 0) return; /* :347-350 -- no swap either */

  /* :353-360 particle count per cell.  Dead in the result (multiplied by
   * k = 0.0 at :443) but kept: the unclamped index would trip the
   * reference's assert for a particle outside the grid; here it is skipped. */
  {
    const float len_x = c->nx * c->dx, len_y = c->ny * c->dy;
    for (int64_t k = 0; k < c->n; ++k)
    {
      int x = (int)((c->part[4 * k] / len_x) * c->nx);
      int y = (int)((c->part[4 * k + 1] / len_y) * c->ny);
      if (x >= 0 && x < c->nx && y >= 0 && y < c->ny) c->n_part[AT(c, x, y)]++;
    }
  }

  LowerLaplacian A;
  A.n = n_fluid;
  A.east = (int*)malloc((size_t)n_fluid * 4);
  A.north = (int*)malloc((size_t)n_fluid * 4);
  A.diag = (float*)malloc((size_t)n_fluid * 4);
  /* :382  1 / pow(deltaX, 2) is evaluated in double and stored as float */
  A.off = (float)(1 / pow(c->dx, 2));
  float* b = (float*)malloc((size_t)n_fluid * 4);
  float* x = (float*)malloc((size_t)n_fluid * 4);
  const float *uf = UF(c), *vf = VF(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      const int idx = c->fluid_idx[AT(c, i, j)];
      if (idx == -1) continue;
      int n_non_solid = 0;
      if (cell_type(c, i - 1, j) != SOLID) n_non_solid++;
      if (cell_type(c, i + 1, j) != SOLID) n_non_solid++;
      if (cell_type(c, i, j - 1) != SOLID) n_non_solid++;
      if (cell_type(c, i, j + 1) != SOLID) n_non_solid++;
      A.east[idx] = cell_type(c, i + 1, j) == LIQUID ? c->fluid_idx[AT(c, i + 1, j)] : -1;
      A.north[idx] = cell_type(c, i, j + 1) == LIQUID ? c->fluid_idx[AT(c, i, j + 1)] : -1;
      A.diag[idx] = (float)(-n_non_solid / pow(c->dx, 2)); /* :409-410 */
      /* include/MacGrid.h:98-111  divVelX + divVelY */
      b[idx] = (uf[AT(c, i + 1, j)] - uf[AT(c, i, j)]) / c->dx +
               (vf[AT(c, i, j + 1)] - vf[AT(c, i, j)]) / c->dy;
    }

  cg_solve(c, &A, b, x);

  memset(c->pressure, 0, C * 4);
  for (size_t k = 0; k < C; ++k)
    if (c->fluid_idx[k] >= 0) c->pressure[k] = x[c->fluid_idx[k]];

  /* :428-482 velocity patch into the back buffer, then swap */
  float *ub = UB(c), *vb = VB(c);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      int im1 = CLAMPi(i - 1, 0, c->nx - 1);
      int jm1 = CLAMPi(j - 1, 0, c->ny - 1);
      int idx = c->fluid_idx[AT(c, i, j)];
      int idx_im1 = c->fluid_idx[AT(c, im1, j)];
      int idx_jm1 = c->fluid_idx[AT(c, i, jm1)];
      if (idx >= 0 || idx_im1 >= 0 || idx_jm1 >= 0)
      {
        const float k = 0.0; /* :443 */
        /* the three particle-pressure terms are k * <int> = +0.0f (:445-453) */
        float pp = k * 0.0f;
        float p = idx >= 0 ? x[idx] + pp : 0;
        float p_im1 = idx_im1 >= 0 ? x[idx_im1] + pp : 0;
        float p_jm1 = idx_jm1 >= 0 ? x[idx_jm1] + pp : 0;
        float pdx = p - p_im1;
        float pdy = p - p_jm1;
        float vx = uf[AT(c, i, j)];
        float vy = vf[AT(c, i, j)];
        ub[AT(c, i, j)] = vx - dt / density * pdx / c->dx;
        vb[AT(c, i, j)] = vy - dt / density * pdy / c->dy;
      }
    }
  swap_velocity(c);
  free(A.east); free(A.north); free(A.diag); free(b); free(x);
}

void fso_get_pressure(void* h, float* dst)
{
  Ctx* c = (Ctx*)h;
  memcpy(dst, c->pressure, (size_t)c->nx * c->ny * 4);
}
int fso_cg_iterations(void* h) { return ((Ctx*)h)->iters; }
float fso_cg_error(void* h) { return ((Ctx*)h)->err; }

/* src/FluidSolver.cpp:921-963 */
void fso_g2p(void* h, int mode, float pic_ratio)
{
  Ctx* c = (Ctx*)h;
  const float *uf = UF(c), *vf = VF(c);
  for (int64_t k = 0; k < c->n; ++k)
  {
    float* q = c->part + 4 * k;
    if (mode == FSX_G2P_PIC)
    {
      float nx_ = vel_x_interp(c, uf, q[0], q[1]);
      float ny_ = vel_y_interp(c, vf, q[0], q[1]);
      q[2] = nx_;
      q[3] = ny_;
    }
    else if (mode == FSX_G2P_FLIP)
    {
      float nx_ = q[2] + vel_x_interp(c, c->u_diff, q[0], q[1]);
      float ny_ = q[3] + vel_y_interp(c, c->v_diff, q[0], q[1]);
      q[2] = nx_;
      q[3] = ny_;
    }
    else
    {
      float pic_x = vel_x_interp(c, uf, q[0], q[1]);
      float pic_y = vel_y_interp(c, vf, q[0], q[1]);
      float flip_x = q[2] + vel_x_interp(c, c->u_diff, q[0], q[1]);
      float flip_y = q[3] + vel_y_interp(c, c->v_diff, q[0], q[1]);
      q[2] = pic_x * pic_ratio + flip_x * (1 - pic_ratio);
      q[3] = pic_y * pic_ratio + flip_y * (1 - pic_ratio);
    }
  }
}

/* src/MarkerParticleSet.cpp:40-62, include/MarkerParticleSet.h:37-41 */
void fso_advect_particles(void* h, float dt, int ensure_outside)
{
  Ctx* c = (Ctx*)h;
  for (int64_t k = 0; k < c->n; ++k)
  {
    float* q = c->part + 4 * k;
    q[0] += q[2] * dt;
    q[1] += q[3] * dt;
    if (ensure_outside)
    {
      int x = (int)(q[0] / c->dx);
      int y = (int)(q[1] / c->dy);
      if (cell_type(c, x, y) == SOLID)
      {
        q[0] += q[2] * -dt;
        q[1] += q[3] * -dt;
      }
    }
  }
}

/* src/FluidSolver.cpp:793-814 with include/OdeSolver.h:102-113 (RK3) or
 * :78-86 (EulerExplicit) and the Vec2 operators of include/FluidSolver.h:119-141 */
static void advected_position(const Ctx* c, float x_pos, float y_pos, float dt,
                              float* xo, float* yo)
{
  const float *uf = UF(c), *vf = VF(c);
  float dxv, dyv;
  if (c->integrator == 1)
  {
    /* f(x + h, y) * h ; Vec2 + MyFloat adds h to BOTH coordinates */
    float ax = x_pos + dt, ay = y_pos + dt;
    dxv = vel_x_interp(c, uf, ax, ay) * dt;
    dyv = vel_y_interp(c, vf, ax, ay) * dt;
  }
  else
  {
    float k1x = vel_x_interp(c, uf, x_pos, y_pos);
    float k1y = vel_y_interp(c, vf, x_pos, y_pos);
    float ax = x_pos + ((k1x * dt) * 1.0f) / 2.0f;
    float ay = y_pos + ((k1y * dt) * 1.0f) / 2.0f;
    float k2x = vel_x_interp(c, uf, ax, ay);
    float k2y = vel_y_interp(c, vf, ax, ay);
    float bx = x_pos + ((k2x * dt) * 3.0f) / 4.0f;
    float by = y_pos + ((k2y * dt) * 3.0f) / 4.0f;
    float k3x = vel_x_interp(c, uf, bx, by);
    float k3y = vel_y_interp(c, vf, bx, by);
    dxv = ((((k1x * 2.0f + k2x * 3.0f) + k3x * 4.0f) * dt) * 1.0f) / 9.0f;
    dyv = ((((k1y * 2.0f + k2y * 3.0f) + k3y * 4.0f) * dt) * 1.0f) / 9.0f;
  }
  *xo = x_pos + dxv;
  *yo = y_pos + dyv;
}

/* src/FluidSolver.cpp:709-772  advectVelocitySemiLagrangian (no swap!) */
void fso_advect_velocity_sl(void* h, float dt)
{
  Ctx* c = (Ctx*)h;
  const size_t C = (size_t)c->nx * c->ny;
  float *ub = UB(c), *vb = VB(c);
  const float *uf = UF(c), *vf = VF(c);
  memset(ub, 0, C * 4);
  memset(vb, 0, C * 4);
  for (int j = 0; j < c->ny; ++j)
    for (int i = 0; i < c->nx; ++i)
    {
      if (cell_type(c, i, j) == LIQUID || cell_type(c, i - 1, j) == LIQUID)
      {
        float x_pos = i * c->dx;
        float y_pos = (float)((j + 0.5) * c->dy);
        float xq, yq;
        advected_position(c, x_pos, y_pos, -dt, &xq, &yq);
        float val = vel_x_interp(c, uf, x_pos, y_pos);
        /* addToVelXInterpolated, include/MacGrid.h:146-151 */
        grid_splat(ub, c->nx, c->ny, c->dx, c->dy, xq, (float)(yq - 0.5 * c->dy), val);
      }
      if (cell_type(c, i, j) == LIQUID || cell_type(c, i, j - 1) == LIQUID)
      {
        float x_pos = (float)((i + 0.5) * c->dx);
        float y_pos = j * c->dy;
        float xq, yq;
        advected_position(c, x_pos, y_pos, -dt, &xq, &yq);
        float val = vel_y_interp(c, vf, x_pos, y_pos);
        grid_splat(vb, c->nx, c->ny, c->dx, c->dy, (float)(xq - 0.5 * c->dx), yq, val);
      }
    }
}

/* src/FluidSolver.cpp:774-791 */
void fso_advect_particles_grid(void* h, float dt)
{
  Ctx* c = (Ctx*)h;
  for (int64_t k = 0; k < c->n; ++k)
  {
    float* q = c->part + 4 * k;
    float xn, yn;
    advected_position(c, q[0], q[1], dt, &xn, &yn);
    q[0] = xn;
    q[1] = yn;
  }
}

/* src/FluidSolver.cpp:89-97 */
static int validate(const Ctx* c)
{
  return fabsf(c->dx - c->pool_dx) < 0.0000001 && fabsf(c->dy - c->pool_dy) < 0.0000001;
}

/* src/FluidSolver.cpp:99-251  the four step drivers */
int fso_step(void* h, int kind, float dt)
{
  Ctx* c = (Ctx*)h;
  if (!validate(c)) return 1;
  const float g = (float)-9.82; /* :115,154,192,232 */
  fso_classify_cells(h);
  if (kind == FSX_STEP_SEMILAGRANGIAN)
  {
    fso_advect_velocity_sl(h, dt);
    fso_add_acceleration(h, 0, g, dt);
    fso_enforce_dirichlet(h);
    fso_pressure_solve(h, c->density, dt);
    fso_enforce_dirichlet(h);
    fso_advect_particles_grid(h, dt);
    return 0;
  }
  fso_p2g_spread(h);
  if (kind != FSX_STEP_PIC) fso_save_previous(h);
  fso_add_acceleration(h, 0, g, dt);
  fso_enforce_dirichlet(h);
  fso_extend_velocity(h, 2);
  fso_pressure_solve(h, c->density, dt);
  fso_enforce_dirichlet(h);
  if (kind == FSX_STEP_PIC)
  {
    fso_g2p(h, FSX_G2P_PIC, 0);
    fso_advect_particles(h, dt, 0);
  }
  else if (kind == FSX_STEP_FLIP)
  {
    fso_update_diff(h);
    fso_g2p(h, FSX_G2P_FLIP, 0);
    fso_advect_particles(h, dt, 0);
  }
  else
  {
    fso_update_diff(h);
    fso_g2p(h, FSX_G2P_PICFLIP, c->pic_ratio);
    fso_advect_particles(h, dt, 1);
  }
  return 0;
}
