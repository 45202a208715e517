"""GPU checks at BASELINE.json's full single-GPU size (4096^2, 6.3e7 particles), where the CPU
checkers are too slow: size-independent properties, each verified with plain numpy on the host.

  * classification: bit-exact against the reference's index formula evaluated in numpy float32
  * particle order: the caller's order survives the device cell sort
  * P2G: a constant particle velocity reproduces the constant on every weighted face, faces without
    weight keep the stale buffer bit for bit; scaling the velocities by 2 scales the result exactly
  * G2P: PIC from a constant grid returns the constant; FLIP with front == previous leaves the
    particle velocities untouched, bit for bit
  * pressure solve: recursive residual below the tolerance AND the true residual ||b - A x|| / ||b||
    recomputed in float64 on the host is small; x is exactly zero outside LIQUID cells; velocity
    after the patch is divergence-free to the same level
  * full PIC/FLIP step: particle count conserved, border stays SOLID, nothing non-finite
"""
import numpy as np
import pytest

import scenes
from oracle_lib import (G2P_FLIP, G2P_PIC, STEP_PICFLIP, U_BACK, U_FRONT, U_PREV, V_BACK, V_FRONT,
                        V_PREV)

pytestmark = pytest.mark.gpu

N = 4096


@pytest.fixture(scope="module")
def scene(capi):
    rng = np.random.default_rng(1234)
    parts = scenes.tank_particles(N, rng, 2)
    dt = float(np.float32(0.01 * 64.0 / N))
    sim = capi.Sim(N, N, 1.0, 1.0, dt, 0.02)
    sim.set_particles(parts)
    return sim, parts, dt


def numpy_labels(parts, n, dx):
    """src/FluidDomain.cpp:157-179 in numpy float32 (IEEE division and multiplication)."""
    f = np.float32
    length = f(n) * f(dx)
    x = ((parts[:, 0] / length) * f(n)).astype(np.int32)
    y = ((parts[:, 1] / length) * f(n)).astype(np.int32)
    x = np.clip(x, 0, n - 1)
    y = np.clip(y, 0, n - 1)
    lab = np.full((n, n), scenes.AIR, dtype=np.uint8)
    lab[y, x] = scenes.LIQUID
    lab[0, :] = lab[-1, :] = scenes.SOLID
    lab[:, 0] = lab[:, -1] = scenes.SOLID
    return lab


def test_classify_bit_exact_at_full_size(scene):
    sim, parts, dt = scene
    sim.classify_cells()
    assert np.array_equal(sim.get_cell_types(), numpy_labels(parts, N, sim.dx))


def test_sort_keeps_caller_order_at_full_size(scene):
    sim, parts, dt = scene
    sim.p2g_spread()  # sorts 6.3e7 particles on the device
    assert np.array_equal(sim.get_particles(), parts)


def test_p2g_constant_field_and_exact_scaling(capi, scene):
    _, parts, dt = scene
    p = parts.copy()
    p[:, 2], p[:, 3] = np.float32(0.75), np.float32(-1.5)
    stale = np.float32(123.0)
    out = []
    for scale in (1.0, 2.0):
        g = capi.Sim(N, N, 1.0, 1.0, dt, 0.02)
        for w in (U_BACK, V_BACK):
            g.set_grid(w, np.full((N, N), stale, dtype=np.float32))
        q = p.copy()
        q[:, 2:] *= np.float32(scale)
        g.set_particles(q)
        g.p2g_spread()
        out.append((g.get_grid(U_FRONT), g.get_grid(V_FRONT)))
        del g
    (u1, v1), (u2, v2) = out
    for a, c in ((u1, 0.75), (v1, -1.5)):
        w = a != stale
        assert w.sum() > 0.9 * (15 / 16) * N * N  # the tank
        assert np.abs(a[w] - np.float32(c)).max() <= 2e-6  # sum(w*c)/sum(w) in fp32
    # power-of-two scaling commutes with every rounding: exact
    wu, wv = u1 != stale, v1 != stale
    assert np.array_equal(u2[wu], u1[wu] * np.float32(2)) and np.array_equal(v2[wv], v1[wv] * np.float32(2))
    assert np.array_equal(u2 == stale, ~wu) and np.array_equal(v2 == stale, ~wv)


def test_g2p_identities_at_full_size(capi, scene):
    _, parts, dt = scene
    g = capi.Sim(N, N, 1.0, 1.0, dt, 0.02)
    g.set_particles(parts)
    cu, cv = np.float32(0.375), np.float32(-2.25)
    for w, c in ((U_FRONT, cu), (V_FRONT, cv), (U_PREV, cu), (V_PREV, cv)):
        g.set_grid(w, np.full((N, N), c, dtype=np.float32))
    g.update_diff()  # diff == 0 exactly
    g.g2p(G2P_FLIP)
    assert np.array_equal(g.get_particles(), parts)  # v + 0 == v
    g.g2p(G2P_PIC)
    got = g.get_particles()
    assert np.array_equal(got[:, :2], parts[:, :2])
    # (1-f)*c + f*c differs from c by at most an ulp or two
    assert np.abs(got[:, 2] - cu).max() <= 1e-6 and np.abs(got[:, 3] - cv).max() <= 1e-6


def laplacian_residual(lab, x, b_u, b_v, dx):
    """||b - A x|| / ||b|| in float64 with the reference's operator."""
    liq = lab == scenes.LIQUID
    nonsolid = lab != scenes.SOLID
    inv = 1.0 / (float(dx) ** 2)
    x = x.astype(np.float64)
    nb = np.zeros_like(x)
    cnt = np.zeros_like(x)
    for sh, ax in ((1, 1), (-1, 1), (1, 0), (-1, 0)):
        nb += np.roll(np.where(liq, x, 0.0), sh, axis=ax)
        cnt += np.roll(nonsolid, sh, axis=ax)
    ax_ = (nb - cnt * x) * inv
    b = ((np.roll(b_u, -1, axis=1) - b_u) / float(dx) + (np.roll(b_v, -1, axis=0) - b_v) / float(dx)).astype(np.float64)
    r = np.where(liq, b - ax_, 0.0)
    return float(np.linalg.norm(r) / np.linalg.norm(np.where(liq, b, 0.0))), liq


def test_pressure_solve_true_residual_at_full_size(capi, scene):
    _, parts, dt = scene
    g = capi.Sim(N, N, 1.0, 1.0, dt, 0.02)
    g.set_cg(400000, 1e-6)
    g.set_particles(parts)
    g.classify_cells(); g.p2g_spread(); g.save_previous()
    g.add_acceleration(0.0, float(np.float32(-9.82)), dt); g.enforce_dirichlet(); g.extend_velocity(2)
    lab, u0, v0 = g.get_cell_types(), g.get_grid(U_FRONT), g.get_grid(V_FRONT)
    g.pressure_solve(dt, dt)
    iters, relres = g.cg_info()
    assert 1000 < iters < 400000 and relres < 1e-6
    x = g.get_pressure()
    true_res, liq = laplacian_residual(lab, x, u0, v0, g.dx)
    assert true_res < 2e-4, true_res  # fp32 recursive residual vs fp64 true residual at 1.6e7 unknowns
    assert np.all(x[~liq] == 0.0)
    # dt / density == 1: the patched field is divergence-free on liquid cells to the same level
    u1, v1 = g.get_grid(U_FRONT), g.get_grid(V_FRONT)
    div0 = ((np.roll(u0, -1, 1) - u0) + (np.roll(v0, -1, 0) - v0)).astype(np.float64)[liq]
    div1 = ((np.roll(u1, -1, 1) - u1) + (np.roll(v1, -1, 0) - v1)).astype(np.float64)[liq]
    assert np.linalg.norm(div1) < 1e-3 * np.linalg.norm(div0)


def test_full_step_invariants_at_full_size(scene):
    sim, parts, dt = scene
    sim.set_cg(400000, 1e-6)
    sim.set_particles(parts)
    for _ in range(2):
        sim.step(STEP_PICFLIP, dt)
    p = sim.get_particles()
    assert p.shape == parts.shape and np.isfinite(p).all()
    lab = sim.get_cell_types()
    assert (lab[0] == 2).all() and (lab[-1] == 2).all() and (lab[:, 0] == 2).all() and (lab[:, -1] == 2).all()
    # classification of the advected particles, again bit-exact against numpy
    sim.classify_cells()
    assert np.array_equal(sim.get_cell_types(), numpy_labels(p, N, sim.dx))
    # particles moved by at most a CFL-sized distance per step
    assert np.abs(p[:, :2] - parts[:, :2]).max() < 4 * sim.dx
