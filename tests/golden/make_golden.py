"""Generate the golden vectors under tests/golden/ from the reference's OWN code.

Runs oracle/_ref/libfsref.so -- the reference's five simulation sources compiled
unchanged from /root/reference against oracle/eigen_shim (see oracle/Makefile) --
so it only works in the build container.  The outputs are committed; the tests
(CPU and GPU) read them and never need /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402


def stage_vectors(ref):
    """Known-answer vectors for every stage on a small seeded scene (24 x 20 cells)."""
    nx, ny = 24, 20
    rng = np.random.default_rng(20240611)
    lab = scenes.random_labels(nx, ny, rng, p_liquid=0.45, p_solid=0.04)
    s = ref.sim(nx, ny, 1.0, float(np.float32(ny) / np.float32(nx)), 0.01, 0.05)
    fields = {w: scenes.random_field(nx, ny, rng) for w in range(8)}
    parts = scenes.particles_in_liquid(lab, s.dx, rng, 3)
    out = {"nx": nx, "ny": ny, "labels": lab, "particles": parts}
    for w, f in fields.items():
        out[f"in_grid{w}"] = f

    def load():
        s.set_cell_types(lab)
        for w, f in fields.items():
            s.set_grid(w, f)
        s.set_particles(parts)

    def grids(tag, which=range(8)):
        for w in which:
            out[f"{tag}_grid{w}"] = s.get_grid(w)

    load(); s.classify_cells(); out["classify_labels"] = s.get_cell_types()
    load(); s.p2g_spread(); grids("p2g", range(4))
    load(); s.save_previous(); s.add_acceleration(0.0, float(np.float32(-9.82)), 0.01)
    s.enforce_dirichlet(); s.update_diff(); grids("gridpre")
    for it in (1, 2, 3):
        load(); s.extend_velocity(it); grids(f"extend{it}", range(4))
    load(); s.set_cg(100, float(np.finfo(np.float32).eps)); s.pressure_solve(0.01, 0.01)
    grids("pressure", range(4)); out["pressure_x"] = s.get_pressure()
    out["pressure_cg"] = np.array(s.cg_info(), dtype=np.float64)
    for mode in (0, 1, 2):
        load(); s.g2p(mode, 0.05); out[f"g2p{mode}_particles"] = s.get_particles()
    load(); s.advect_particles(0.01, True); out["advect_particles"] = s.get_particles()
    load(); s.advect_velocity_sl(0.25 * s.dx); grids("advsl", range(4))
    load(); s.advect_particles_grid(0.01); out["advgrid_particles"] = s.get_particles()
    np.savez_compressed(os.path.join(HERE, "stages_24x20.npz"), **out)


def uncalled_vectors(ref):
    """addExternalForce and transferVelocityToGridGather (no step calls them): a separate file so
    that stages_24x20.npz stays byte-identical to its first commit."""
    nx, ny = 24, 20
    rng = np.random.default_rng(20240612)
    lab = scenes.random_labels(nx, ny, rng, p_liquid=0.45, p_solid=0.04)
    s = ref.sim(nx, ny, 1.0, float(np.float32(ny) / np.float32(nx)), 0.013, 0.05)
    fields = {w: scenes.random_field(nx, ny, rng) for w in range(4)}
    parts = scenes.particles_in_liquid(lab, s.dx, rng, 3)
    dx, dy = np.float32(s.dx), np.float32(s.dy)
    k = 0  # particles exactly ON face positions: the only ones the gather transfer selects
    for (i, j) in [(3, 4), (5, 5), (5, 5), (7, 2), (10, 10), (0, 3), (nx - 1, 5), (12, ny - 1)]:
        parts[k, 0] = np.float32(i) * dx; parts[k, 1] = np.float32((j + 0.5) * np.float64(dy)); k += 1
        parts[k, 0] = np.float32((i + 0.5) * np.float64(dx)); parts[k, 1] = np.float32(j) * dy; k += 1
    out = {"nx": nx, "ny": ny, "labels": lab, "particles": parts, "density": np.float32(0.013)}
    for w, f in fields.items():
        out[f"in_grid{w}"] = f
    s.set_cell_types(lab)
    for w, f in fields.items():
        s.set_grid(w, f)
    s.set_particles(parts)
    s.add_external_force(0.3, -1.7, 0.01)
    for w in range(4):
        out[f"force_grid{w}"] = s.get_grid(w)
    s.p2g_gather()
    for w in range(4):
        out[f"gather_grid{w}"] = s.get_grid(w)
    # extendVelocityAvarageing (src/FluidSolver.cpp:625-707) from the same inputs, 1 / 2 / 3 sweeps
    for it in (1, 2, 3):
        s.set_cell_types(lab)
        for w, f in fields.items():
            s.set_grid(w, f)
        s.extend_velocity_avg(it)
        for w in range(4):
            out[f"extavg{it}_grid{w}"] = s.get_grid(w)
    np.savez_compressed(os.path.join(HERE, "uncalled_24x20.npz"), **out)


def frame_vectors(ref):
    """The reference renderer's frame of the initial examples/simple.cpp scene (particles emitted,
    cells classified; no solver step, so the state is reproducible without the CG)."""
    n = 64
    s = ref.sim(n, n, 1.0, 1.0, 0.01, 0.05)
    s.emit_source(*scenes.dam_break_args(n))
    s.classify_cells()
    out = {}
    for tag, (w, h, area) in {"full": (160, 160, (0, 1, 0, 1)), "zoom": (97, 61, (0.1, 0.6, 0.3, 0.9)),
                              "wide": (64, 48, (-0.5, 1.5, -0.2, 1.3))}.items():
        out[f"{tag}_area"] = np.array(area, dtype=np.float32)
        out[f"{tag}_rgb"] = s.render_rgb(w, h, area)
    np.savez_compressed(os.path.join(HERE, "frames_config0.npz"), **out)


def config0_trace(ref):
    """examples/simple.cpp scene: 64 x 64, one source, dt = 0.01, stepPICFLIP."""
    n = 64
    out = {}
    for kind, name in ((ol.STEP_PICFLIP, "picflip"), (ol.STEP_SL, "sl"), (ol.STEP_FLIP, "flip"),
                       (ol.STEP_PIC, "pic")):
        s = ref.sim(n, n, 1.0, 1.0, 0.01, 0.05)
        cnt = s.emit_source(*scenes.dam_break_args(n))
        assert cnt == 7800
        liquid, cg, mean = [], [], []
        for step in range(10):
            s.step(kind, 0.01)
            lab = s.get_cell_types()
            p = s.get_particles()
            liquid.append(int((lab == 0).sum()))
            cg.append(s.cg_info())
            mean.append(p.astype(np.float64).mean(axis=0))
            if step in (0, 2):
                out[f"{name}_labels_step{step}"] = np.packbits(lab == 0)
                out[f"{name}_particles_step{step}"] = p[::13].copy()
        out[f"{name}_liquid"] = np.array(liquid)
        out[f"{name}_cg"] = np.array(cg, dtype=np.float64)
        out[f"{name}_mean"] = np.array(mean)
    np.savez_compressed(os.path.join(HERE, "config0_trace.npz"), **out)


if __name__ == "__main__":
    assert ol.available("fsr"), "build oracle/_ref first: make -C oracle ref"
    ref = ol.OracleLib("fsr")
    if "--only-uncalled" not in sys.argv:
        stage_vectors(ref)
        config0_trace(ref)
    uncalled_vectors(ref)
    frame_vectors(ref)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
