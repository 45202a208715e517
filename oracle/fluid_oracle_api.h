/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * One C API shape shared by the two CPU checkers:
 *   prefix fsr_  oracle/_ref/libfsref.so   the reference's own sources, compiled
 *                                          unchanged from /root/reference against
 *                                          oracle/eigen_shim (oracle/ref_driver.cpp)
 *   prefix fso_  oracle/libfsoracle.so     plain-C restatement (oracle/fluid_oracle.c)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load either library.
 */
#ifndef FSB_ORACLE_API_H
#define FSB_ORACLE_API_H

#include <stdint.h>

#ifndef FSX
#error "define FSX(name) to the symbol prefix before including"
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* grid selectors (same numbering as include/fsb.h) */
enum { FSX_U_FRONT = 0, FSX_V_FRONT, FSX_U_BACK, FSX_V_BACK,
       FSX_U_PREV, FSX_V_PREV, FSX_U_DIFF, FSX_V_DIFF };
/* g2p modes / step kinds */
enum { FSX_G2P_PIC = 0, FSX_G2P_FLIP = 1, FSX_G2P_PICFLIP = 2 };
enum { FSX_STEP_SEMILAGRANGIAN = 0, FSX_STEP_PIC = 1, FSX_STEP_FLIP = 2, FSX_STEP_PICFLIP = 3 };

void* FSX(create)(int size_x, int size_y, float length_x, float length_y,
                  float density, float pic_ratio);
void FSX(destroy)(void* h);
float FSX(delta_x)(void* h);
float FSX(delta_y)(void* h);
void FSX(set_cg)(void* h, int max_iters, float tol);

void FSX(set_particles)(void* h, const float* aos4, int64_t n);
void FSX(append_particles)(void* h, const float* aos4, int64_t n);
int64_t FSX(num_particles)(void* h);
void FSX(get_particles)(void* h, float* aos4);
/* One FluidSource spawn (src/FluidDomain.cpp:29-52); returns particles added. */
int64_t FSX(emit_source)(void* h, float x_min, float x_max, float y_min, float y_max,
                         float delta_x, float delta_y, float vel_x, float vel_y);

void FSX(set_grid)(void* h, int which, const float* src);
void FSX(get_grid)(void* h, int which, float* dst);
void FSX(set_cell_types)(void* h, const uint8_t* src);
void FSX(get_cell_types)(void* h, uint8_t* dst);

void FSX(classify_cells)(void* h);
void FSX(p2g_spread)(void* h);
void FSX(save_previous)(void* h);
void FSX(add_acceleration)(void* h, float ax, float ay, float dt);
void FSX(enforce_dirichlet)(void* h);
void FSX(extend_velocity)(void* h, int n_iter);
void FSX(pressure_solve)(void* h, float density, float dt);
void FSX(get_pressure)(void* h, float* dst_full_grid);
int FSX(cg_iterations)(void* h);
float FSX(cg_error)(void* h);
void FSX(update_diff)(void* h);
void FSX(g2p)(void* h, int mode, float pic_ratio);
void FSX(advect_particles)(void* h, float dt, int ensure_outside);
void FSX(advect_velocity_sl)(void* h, float dt);
void FSX(advect_particles_grid)(void* h, float dt);
/* routines of FluidSolver that no step* calls */
/* addExternalForce src/FluidSolver.cpp:253-274 (reads the domain's density) */
void FSX(add_external_force)(void* h, float fx, float fy, float dt);
/* transferVelocityToGridGather src/FluidSolver.cpp:816-871 */
void FSX(p2g_gather)(void* h);
/* extendVelocityAvarageing src/FluidSolver.cpp:625-707 (border cells must be SOLID: the reference
 * asserts on the index otherwise) */
void FSX(extend_velocity_avg)(void* h, int n_iterations);
/* One frame as examples/simple.cpp:73-82 draws it: Renderer::clearCanvas, renderGridCellsToCanvas,
 * renderParticlesToCanvas (src/Renderer.cpp:14-56,141-162) on a width x height canvas over the
 * world-space area, converted to bytes like writeCanvasToPpm (:217-248).  rgb: width*height*3. */
void FSX(render_rgb)(void* h, int width, int height, float x_min, float x_max, float y_min,
                     float y_max, uint8_t* rgb);
/* returns 0, or 1 when the reference's validate() would throw */
int FSX(step)(void* h, int kind, float dt);

#ifdef __cplusplus
}
#endif
#endif
