"""The C++ host classes of include/fsb/ (the reference's FluidSolver / FluidDomain / MacGrid /
MarkerParticleSet API over libfsb.so), driven by tests/host_api_demo.cpp the way the reference's
examples/simple.cpp drives fluidsim_lib."""
import os
import re
import subprocess

import numpy as np
import pytest

import scenes
from oracle_lib import G2P_PICFLIP, STEP_PICFLIP, STEP_SL, U_FRONT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "host_api_demo")


@pytest.fixture(scope="session")
def demo(built):
    import __graft_entry__ as g
    g.build_host_demo()
    assert os.path.exists(DEMO)
    return DEMO


def test_host_classes_compile_and_fail_loudly_without_gpu(demo):
    """CPU box: the program builds against include/fsb/ and libfsb.so; with no device the first
    FluidDomain constructor throws (no CPU fallback) and the program reports it."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([demo, "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout or "no CPU fallback" in r.stdout


def test_headers_keep_the_reference_api():
    """Every public name the reference's callers use exists in include/fsb/."""
    text = "".join(open(os.path.join(ROOT, "include", "fsb", f)).read()
                   for f in os.listdir(os.path.join(ROOT, "include", "fsb")))
    for name in ["class FluidSolver", "class FluidSolverMemoryPool", "class FluidDomain",
                 "class FluidSource", "class MacGrid", "class MarkerParticleSet",
                 "class MarkerParticle", "class GridInterface", "struct BBox",
                 "stepSemiLagrangian", "stepPIC", "stepFLIP", "stepPICFLIP", "addFluidSource",
                 "clearFluidSources", "resetParticleSet", "setPicRatio", "macGrid",
                 "markerParticleSet", "density", "picRatio", "classifyCells", "addParticle",
                 "advectAndEnsureOutsideObstacles", "velXHalfIndexed", "velYHalfIndexed",
                 "velXBackBufferHalfIndexed", "velXInterpolated", "velYInterpolated",
                 "velXDiffInterpolated", "cellType", "divVelX", "divVelY", "setVelXHalfIndexed",
                 "setVelYBackBuffer", "setCellType", "addToVelXInterpolated",
                 "swapVelocityBuffers", "clearCellTypeBuffer", "updatePreviousVelocityBuffer",
                 "updateVelocityDiffBuffer", "isFinished", "resetSpawns", "lengthX", "worldToCell"]:
        assert name in text, name


def _parse(out):
    steps, other = [], {}
    for line in out.splitlines():
        t = line.split()
        if t and t[0] == "STEP":
            steps.append(dict(step=int(t[1]), particles=int(t[3]), liquid=int(t[5]), solid=int(t[7]),
                              mean=[float(x) for x in t[9:13]], cg=(int(t[14]), float(t[15])),
                              probe=(float(t[17]), float(t[18]))))
        elif t:
            other[t[0]] = line
    return steps, other


@pytest.mark.gpu
@pytest.mark.parametrize("kind,step_kind", [("picflip", STEP_PICFLIP), ("sl", STEP_SL)])
def test_simple_cpp_scene_through_host_classes(demo, port, kind, step_kind):
    n_steps = 10
    r = subprocess.run([demo, str(n_steps), "64", kind], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    steps, other = _parse(r.stdout)
    assert len(steps) == n_steps
    c = port.sim(64, 64, 1.0, 1.0, 0.01, 0.05)
    assert c.emit_source(*scenes.dam_break_args(64)) == 7800
    for k, s in enumerate(steps):
        c.step(step_kind, 0.01)
        lab, p = c.get_cell_types(), c.get_particles()
        assert s["particles"] == 7800 and s["solid"] == 252
        n_liq = int((lab == 0).sum())
        if k < 3:
            assert s["liquid"] == n_liq
        assert abs(s["liquid"] - n_liq) <= 0.02 * n_liq + 2
        mean = p.astype(np.float64).mean(axis=0)
        assert np.abs(np.array(s["mean"][:2]) - mean[:2]).max() < 1e-4 * (k + 1)
        ic = c.cg_info()[0]
        assert abs(s["cg"][0] - ic) <= max(2, 0.1 * ic)
        # MacGrid::velX / velY (cell-centred averages) of one cell, read through the host mirror
        u, v = c.get_grid(0), c.get_grid(1)
        ci = cj = 16
        ref_probe = ((u[cj, ci] + u[cj, ci + 1]) / 2, (v[cj, ci] + v[cj + 1, ci]) / 2)
        assert np.allclose(s["probe"], ref_probe, atol=2e-3 * (k + 1))
    if kind != "picflip":
        return
    # host edits between steps: addParticle, setPicRatio (clamped), setVelXHalfIndexed, setCellType
    assert "7800 -> 7801" in other["EDIT"] and "picratio 1 " in other["EDIT"]
    c.append_particles(np.array([[0.5, 0.5, 0.25, -0.5]], dtype=np.float32))
    u = c.get_grid(U_FRONT); u[5, 5] = 1.5; c.set_grid(U_FRONT, u)
    lab = c.get_cell_types(); lab[32, 32] = 2; c.set_cell_types(lab)
    grav = float(np.float32(-9.82))
    c.classify_cells(); c.p2g_spread(); c.save_previous(); c.add_acceleration(0.0, grav, 0.01)
    c.enforce_dirichlet(); c.extend_velocity(2); c.pressure_solve(0.01, 0.01); c.enforce_dirichlet()
    c.update_diff(); c.g2p(G2P_PICFLIP, 1.0); c.advect_particles(0.01, True)
    last = c.get_particles()[-1]
    got = [float(x) for x in other["EDIT"].split()[-2:]]
    assert np.allclose(got, last[:2], atol=1e-4)
    assert "RESET particles 0 liquid0 0" in other["RESET"]
    assert re.search(r"ERROR-CASE Memory pool and fluid domain does not match", other["ERROR-CASE"])
