"""Seeded synthetic scenes shared by the CPU and GPU tests (numpy only)."""
import numpy as np

LIQUID, AIR, SOLID = 0, 1, 2


def dam_break_args(n):
    """The FluidSource box of examples/simple.cpp:29 scaled to an n x n grid."""
    f = np.float32
    return f(2.0 / n), f(0.35), f(2.0 / n), f(1 - 2.0 / n)


def random_labels(nx, ny, rng, p_liquid=0.5, p_solid=0.03):
    """Border SOLID (as classifyCells always leaves it), blobby liquid, sparse interior solids,
    and an AIR cap in the upper rows so the Poisson system has a Dirichlet boundary."""
    lab = np.full((ny, nx), AIR, dtype=np.uint8)
    yy, xx = np.mgrid[0:ny, 0:nx]
    field = np.zeros((ny, nx))
    for _ in range(6):
        cx, cy = rng.uniform(0, nx), rng.uniform(0, ny * 0.8)
        r = rng.uniform(0.1, 0.35) * min(nx, ny)
        field += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))
    thr = np.quantile(field, 1 - p_liquid)
    lab[field > thr] = LIQUID
    lab[rng.uniform(size=lab.shape) < p_solid] = SOLID
    lab[int(ny * 0.9):, :] = np.where(lab[int(ny * 0.9):, :] == LIQUID, AIR, lab[int(ny * 0.9):, :])
    lab[0, :] = lab[-1, :] = SOLID
    lab[:, 0] = lab[:, -1] = SOLID
    return lab


def random_field(nx, ny, rng, scale=1.0):
    return (rng.standard_normal((ny, nx)) * scale).astype(np.float32)


def particles_in_liquid(lab, dx, rng, per_cell=4, vel_scale=1.0, jitter=True):
    """per_cell particles in every LIQUID cell of `lab`, random sub-cell positions."""
    ny, nx = lab.shape
    jj, ii = np.nonzero(lab == LIQUID)
    n = ii.size * per_cell
    ii = np.repeat(ii, per_cell).astype(np.float64)
    jj = np.repeat(jj, per_cell).astype(np.float64)
    ox = rng.uniform(0.02, 0.98, n) if jitter else np.full(n, 0.5)
    oy = rng.uniform(0.02, 0.98, n) if jitter else np.full(n, 0.5)
    p = np.empty((n, 4), dtype=np.float32)
    p[:, 0] = ((ii + ox) * dx).astype(np.float32)
    p[:, 1] = ((jj + oy) * dx).astype(np.float32)
    p[:, 2:] = (rng.standard_normal((n, 2)) * vel_scale).astype(np.float32)
    perm = rng.permutation(n)  # host order is NOT cell order
    return p[perm]


def tank_particles(n, rng, per_side=2, fill=15.0 / 16.0):
    """SURVEY.md 8(d) config 2 'tank': per_side^2 stratified-jittered particles in every
    non-SOLID cell with j < fill*n, swirl velocity (sin pi x cos pi y, -cos pi x sin pi y)."""
    dx = np.float32(1.0) / np.float32(n)
    j_top = int(fill * n)
    ii, jj = np.meshgrid(np.arange(1, n - 1), np.arange(1, j_top), indexing="xy")
    ii = ii.ravel().astype(np.float64)
    jj = jj.ravel().astype(np.float64)
    parts = []
    for sy in range(per_side):
        for sx in range(per_side):
            ox = (sx + rng.uniform(0.05, 0.95, ii.size)) / per_side
            oy = (sy + rng.uniform(0.05, 0.95, ii.size)) / per_side
            x = (ii + ox) * float(dx)
            y = (jj + oy) * float(dx)
            u = np.sin(np.pi * x) * np.cos(np.pi * y)
            v = -np.cos(np.pi * x) * np.sin(np.pi * y)
            parts.append(np.stack([x, y, u, v], axis=1))
    p = np.concatenate(parts).astype(np.float32)
    return p


def field_rel_err(a, b):
    """max |a-b| / max |b| : the field-relative measure of SURVEY.md 8(d)."""
    den = float(np.abs(b).max())
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / (den if den > 0 else 1.0)
