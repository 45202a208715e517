"""CPU tests of the checkers themselves (no GPU).

The plain-C restatement (oracle/libfsoracle.so) must be BIT-IDENTICAL to
  (a) the committed golden vectors in tests/golden/ (generated from the reference's own
      sources by tests/golden/make_golden.py), and
  (b) oracle/_ref/libfsref.so when it is present (the build container),
and must reproduce the Eigen-independent facts recorded in SURVEY.md 8(c).
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_stage_case(sim, z):
    sim.set_cell_types(z["labels"])
    for w in range(8):
        sim.set_grid(w, z[f"in_grid{w}"])
    sim.set_particles(z["particles"])


def test_port_matches_golden_stage_vectors(port):
    z = np.load(os.path.join(GOLD, "stages_24x20.npz"))
    nx, ny = int(z["nx"]), int(z["ny"])
    s = port.sim(nx, ny, 1.0, float(np.float32(ny) / np.float32(nx)), 0.01, 0.05)

    def same_grids(tag, which):
        for w in which:
            assert np.array_equal(s.get_grid(w), z[f"{tag}_grid{w}"]), (tag, w)

    _load_stage_case(s, z); s.classify_cells()
    assert np.array_equal(s.get_cell_types(), z["classify_labels"])
    _load_stage_case(s, z); s.p2g_spread(); same_grids("p2g", range(4))
    _load_stage_case(s, z); s.save_previous()
    s.add_acceleration(0.0, float(np.float32(-9.82)), 0.01); s.enforce_dirichlet(); s.update_diff()
    same_grids("gridpre", range(8))
    for it in (1, 2, 3):
        _load_stage_case(s, z); s.extend_velocity(it); same_grids(f"extend{it}", range(4))
    _load_stage_case(s, z); s.set_cg(100, float(np.finfo(np.float32).eps)); s.pressure_solve(0.01, 0.01)
    same_grids("pressure", range(4))
    assert np.array_equal(s.get_pressure(), z["pressure_x"])
    assert np.array_equal(np.array(s.cg_info(), dtype=np.float64), z["pressure_cg"])
    for mode in (0, 1, 2):
        _load_stage_case(s, z); s.g2p(mode, 0.05)
        assert np.array_equal(s.get_particles(), z[f"g2p{mode}_particles"])
    _load_stage_case(s, z); s.advect_particles(0.01, True)
    assert np.array_equal(s.get_particles(), z["advect_particles"])
    _load_stage_case(s, z); s.advect_velocity_sl(0.25 * s.dx); same_grids("advsl", range(4))
    _load_stage_case(s, z); s.advect_particles_grid(0.01)
    assert np.array_equal(s.get_particles(), z["advgrid_particles"])


@pytest.mark.parametrize("kind,name", [(ol.STEP_PICFLIP, "picflip"), (ol.STEP_SL, "sl"),
                                        (ol.STEP_FLIP, "flip"), (ol.STEP_PIC, "pic")])
def test_port_matches_golden_config0_trace(port, kind, name):
    z = np.load(os.path.join(GOLD, "config0_trace.npz"))
    s = port.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
    assert s.emit_source(*scenes.dam_break_args(64)) == 7800
    for step in range(10):
        s.step(kind, 0.01)
        lab, p = s.get_cell_types(), s.get_particles()
        assert int((lab == 0).sum()) == z[f"{name}_liquid"][step]
        assert np.array_equal(np.array(s.cg_info(), dtype=np.float64), z[f"{name}_cg"][step])
        assert np.array_equal(p.astype(np.float64).mean(axis=0), z[f"{name}_mean"][step])
        if step in (0, 2):
            assert np.array_equal(np.packbits(lab == 0), z[f"{name}_labels_step{step}"])
            assert np.array_equal(p[::13], z[f"{name}_particles_step{step}"])


def test_survey_known_facts(port):
    """7 800 particles, 1 260 / 1 281 / 1 317 LIQUID cells at steps 0-2 with 252
    SOLID border cells, mean particle position after step 0 = (0.190625759, 0.495918758),
    99 CG iterations at step 0."""
    s = port.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
    assert s.emit_source(*scenes.dam_break_args(64)) == 7800
    liquid = []
    for step in range(3):
        s.step(ol.STEP_PICFLIP, 0.01)
        lab = s.get_cell_types()
        liquid.append(int((lab == 0).sum()))
        assert int((lab == 2).sum()) == 252
        if step == 0:
            m = s.get_particles().astype(np.float64).mean(axis=0)
            assert abs(m[0] - 0.190625759) < 1e-8 and abs(m[1] - 0.495918758) < 1e-8
            assert s.cg_info()[0] == 99
    assert liquid == [1260, 1281, 1317]


def test_port_equals_compiled_reference_long_run(port, ref):
    """100 PIC/FLIP steps and 60 semi-Lagrangian steps: every particle and every grid bit."""
    for kind, n_steps in ((ol.STEP_PICFLIP, 100), (ol.STEP_SL, 60)):
        a = port.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
        b = ref.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
        a.emit_source(*scenes.dam_break_args(64))
        b.emit_source(*scenes.dam_break_args(64))
        for _ in range(n_steps):
            a.step(kind, 0.01)
            b.step(kind, 0.01)
        assert np.array_equal(a.get_particles(), b.get_particles())
        assert np.array_equal(a.get_cell_types(), b.get_cell_types())
        for w in range(8):
            assert np.array_equal(a.get_grid(w), b.get_grid(w))
        assert a.cg_info() == b.cg_info()
        assert np.array_equal(a.get_pressure(), b.get_pressure())


@pytest.mark.parametrize("nx,ny", [(37, 53), (96, 40)])
def test_port_equals_compiled_reference_random_stages(port, ref, nx, ny):
    rng = np.random.default_rng(99)
    ly = float(np.float32(ny) / np.float32(nx))
    a, b = port.sim(nx, ny, 1.0, ly), ref.sim(nx, ny, 1.0, ly)
    lab = scenes.random_labels(nx, ny, rng)
    fields = {w: scenes.random_field(nx, ny, rng) for w in range(8)}
    parts = scenes.particles_in_liquid(lab, a.dx, rng, 4)

    def load():
        for s in (a, b):
            s.set_cell_types(lab)
            for w, f in fields.items():
                s.set_grid(w, f)
            s.set_particles(parts)

    ops = [lambda s: s.classify_cells(), lambda s: s.p2g_spread(),
           lambda s: (s.save_previous(), s.add_acceleration(0.3, -9.0, 0.01), s.enforce_dirichlet(),
                      s.update_diff()),
           lambda s: s.extend_velocity(2), lambda s: s.extend_velocity(3),
           lambda s: s.pressure_solve(0.02, 0.01), lambda s: s.g2p(2, 0.3),
           lambda s: s.advect_particles(0.02, True), lambda s: s.advect_velocity_sl(0.004),
           lambda s: s.advect_particles_grid(-0.01)]
    for op in ops:
        load()
        op(a)
        op(b)
        assert np.array_equal(a.get_particles(), b.get_particles())
        assert np.array_equal(a.get_cell_types(), b.get_cell_types())
        for w in range(8):
            assert np.array_equal(a.get_grid(w), b.get_grid(w))
        assert np.array_equal(a.get_pressure(), b.get_pressure())


def test_port_matches_golden_uncalled_routines(checkers):
    """addExternalForce / transferVelocityToGridGather against vectors from the reference's code."""
    z = np.load(os.path.join(GOLD, "uncalled_24x20.npz"))
    nx, ny = int(z["nx"]), int(z["ny"])
    for chk in checkers:
        s = chk.sim(nx, ny, 1.0, float(np.float32(ny) / np.float32(nx)), float(z["density"]), 0.05)
        s.set_cell_types(z["labels"])
        for w in range(4):
            s.set_grid(w, z[f"in_grid{w}"])
        s.set_particles(z["particles"])
        s.add_external_force(0.3, -1.7, 0.01)
        for w in range(4):
            assert np.array_equal(s.get_grid(w), z[f"force_grid{w}"]), ("force", w)
        s.p2g_gather()
        for w in range(4):
            assert np.array_equal(s.get_grid(w), z[f"gather_grid{w}"]), ("gather", w)
        assert (z["gather_grid0"] != z["force_grid2"]).sum() >= 5  # planted particles were selected
        # extendVelocityAvarageing (src/FluidSolver.cpp:625-707), 1 / 2 / 3 sweeps
        for it in (1, 2, 3):
            s.set_cell_types(z["labels"])
            for w in range(4):
                s.set_grid(w, z[f"in_grid{w}"])
            s.extend_velocity_avg(it)
            for w in range(4):
                assert np.array_equal(s.get_grid(w), z[f"extavg{it}_grid{w}"]), ("extend_avg", it, w)
        assert (z["extavg2_grid0"] != z["in_grid2"]).sum() > 20  # it did extend something


def test_port_matches_golden_frames(checkers):
    """The frame renderer (Renderer + Canvas restated) against frames from the reference's code."""
    z = np.load(os.path.join(GOLD, "frames_config0.npz"))
    for chk in checkers:
        s = chk.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
        assert s.emit_source(*scenes.dam_break_args(64)) == 7800
        s.classify_cells()
        for tag in ("full", "zoom", "wide"):
            ref_rgb = z[f"{tag}_rgb"]
            h, w, _ = ref_rgb.shape
            got = s.render_rgb(w, h, tuple(float(v) for v in z[f"{tag}_area"]))
            assert np.array_equal(got, ref_rgb), tag
    assert len(np.unique(z["full_rgb"].reshape(-1, 3), axis=0)) >= 3  # white, solid, particles (+ liquid)


def test_validate_rejects_non_square_cells(port):
    s = port.sim(32, 16, 1.0, 1.0)  # dx = 1/32, dy = 1/16 (src/FluidSolver.cpp:56-65,89-97)
    with pytest.raises(RuntimeError):
        s.step(ol.STEP_PICFLIP, 0.01)


def test_cg_converged_solution_solves_the_system(port):
    """With the cap lifted the restated Eigen CG drives the residual below tol."""
    s = port.sim(48, 48, 1.0, 1.0, 0.01, 0.05)
    s.emit_source(*scenes.dam_break_args(48))
    s.set_cg(5000, 1e-6)
    s.step(ol.STEP_PICFLIP, 0.01)
    it, err = s.cg_info()
    assert 0 < it < 5000 and err < 1e-6
