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


def numpy_operator(lab, dx):
    """The reference's pressure operator as float64 numpy closures."""
    liq = lab == scenes.LIQUID
    nonsolid = lab != scenes.SOLID
    inv = 1.0 / (float(dx) ** 2)
    cnt = np.zeros(lab.shape)
    for sh, ax in ((1, 1), (-1, 1), (1, 0), (-1, 0)):
        cnt += np.roll(nonsolid, sh, axis=ax)
    diag = np.where(liq, -cnt * inv, 1.0)

    def apply(x):
        xm = np.where(liq, x, 0.0)
        nb = np.zeros(lab.shape)
        for sh, ax in ((1, 1), (-1, 1), (1, 0), (-1, 0)):
            nb += np.roll(xm, sh, axis=ax)
        return np.where(liq, nb * inv + diag * xm, 0.0)

    return liq, diag, apply


def numpy_rhs(liq, u, v, dx):
    b = (np.roll(u, -1, axis=1) - u).astype(np.float64) / float(dx) + \
        (np.roll(v, -1, axis=0) - v).astype(np.float64) / float(dx)
    return np.where(liq, b, 0.0)


def numpy_pcg(liq, diag, apply, b, n_iter):
    """Eigen's Jacobi-PCG recurrences in float64, exactly n_iter iterations."""
    x = np.zeros_like(b)
    r = b.copy()
    p = np.where(liq, r / diag, 0.0)
    abs_new = float((r * p).sum())
    for _ in range(n_iter):
        q = apply(p)
        alpha = abs_new / float((p * q).sum())
        x += alpha * p
        r -= alpha * q
        z = np.where(liq, r / diag, 0.0)
        abs_old, abs_new = abs_new, float((r * z).sum())
        p = z + (abs_new / abs_old) * p
    return x


def prepared_for_solve(capi, parts, dt, n):
    g = capi.Sim(n, n, 1.0, 1.0, dt, 0.02)
    g.set_particles(parts)
    g.classify_cells(); g.p2g_spread(); g.save_previous()
    g.add_acceleration(0.0, float(np.float32(-9.82)), dt); g.enforce_dirichlet(); g.extend_velocity(2)
    return g


@pytest.mark.parametrize("n_iter", [1, 3])
def test_first_cg_iterates_match_float64_at_full_size(capi, scene, n_iter):
    """The iteration kernels at 4096^2 (TMA tiles, halo recompute, both ping-pong phases, the
    reductions) against the same recurrences in float64 numpy: with the iteration count capped at
    1 and 3 the iterate is a short, well-conditioned expression of b, so fp32 agrees to ~1e-5."""
    _, parts, dt = scene
    g = prepared_for_solve(capi, parts, dt, N)
    lab, u0, v0 = g.get_cell_types(), g.get_grid(U_FRONT), g.get_grid(V_FRONT)
    g.set_cg(n_iter, 1e-30)
    g.pressure_solve(dt, dt)
    assert g.cg_info()[0] == n_iter
    x = g.get_pressure().astype(np.float64)
    liq, diag, apply = numpy_operator(lab, g.dx)
    ref = numpy_pcg(liq, diag, apply, numpy_rhs(liq, u0, v0, g.dx), n_iter)
    assert np.all(x[~liq] == 0.0)
    err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
    assert err < 2e-5, err
    assert np.abs(x - ref).max() <= 1e-4 * np.abs(ref).max()


def test_converged_solve_at_full_size(capi, scene):
    """Converged solve at 4096^2: the stopping rule (recursive residual, as Eigen's) is met and x is
    exactly zero outside LIQUID cells.  The TRUE residual of an fp32 CG at kappa ~ 7e6 stagnates
    near kappa * eps (the reference has the same property), so it is checked at a
    well-conditioned size in the next test instead."""
    _, parts, dt = scene
    g = prepared_for_solve(capi, parts, dt, N)
    g.set_cg(400000, 1e-6)
    lab = g.get_cell_types()
    g.pressure_solve(dt, dt)
    iters, relres = g.cg_info()
    assert 1000 < iters < 400000 and relres < 1e-6
    x = g.get_pressure()
    assert np.isfinite(x).all() and np.all(x[lab != scenes.LIQUID] == 0.0)


def test_true_residual_and_divergence_at_128(capi):
    """(the CPU checker reaches a true residual of 1.3e-3 on this scene; kappa * eps grows with n^2)"""
    n = 128
    dt = float(np.float32(0.01 * 64.0 / n))
    parts = scenes.tank_particles(n, np.random.default_rng(7), 2)
    g = prepared_for_solve(capi, parts, dt, n)
    g.set_cg(400000, 1e-6)
    lab, u0, v0 = g.get_cell_types(), g.get_grid(U_FRONT), g.get_grid(V_FRONT)
    g.pressure_solve(dt, dt)
    assert g.cg_info()[1] < 1e-6
    x = g.get_pressure().astype(np.float64)
    liq, diag, apply = numpy_operator(lab, g.dx)
    b = numpy_rhs(liq, u0, v0, g.dx)
    true_res = np.linalg.norm(b - apply(x)) / np.linalg.norm(b)
    assert true_res < 5e-3, true_res
    # dt / density == 1: the patched field has divergence b - A x on every liquid cell without a
    # SOLID neighbour.  (Next to a wall the reference's patch uses p = 0 for the SOLID side,
    # src/FluidSolver.cpp:455-460, while the operator treats the wall as Neumann: those faces are
    # only settled by the enforceDirichlet that follows, so they are excluded here.)
    u1, v1 = g.get_grid(U_FRONT), g.get_grid(V_FRONT)
    solid = lab == scenes.SOLID
    near_wall = np.zeros_like(solid)
    for sh, ax in ((1, 1), (-1, 1), (1, 0), (-1, 0)):
        near_wall |= np.roll(solid, sh, axis=ax)
    inner = liq & ~near_wall
    d1 = numpy_rhs(inner, u1, v1, g.dx)
    res = np.where(inner, b - apply(x), 0.0)
    assert np.linalg.norm(d1) < 1e-2 * np.linalg.norm(b)
    assert np.linalg.norm(d1 - res) < 1e-3 * np.linalg.norm(b)


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
