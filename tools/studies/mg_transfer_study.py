"""Why the opt-in multigrid PCG (fsb_mg.cu) needed 17 / 24 / 50 iterations at 1024^2 / 2048^2 / 4096^2, and the fix.

Findings (numpy, tools/studies/mgpcg_prototype.py = the V-cycle of the CUDA code):
  * the convergence history has a plateau: x10 per iteration down to 2e-4, then ~10 iterations around 1e-4 .. 1e-5.
    Same history in float64: not a precision effect;
  * largest eigenvalue of M A (power iteration, exact coarsest solve) on the tank scene:
        n = 64 / 128 / 256 / 512:   1.18 / 1.73 / 3.32 / 6.07      plain transfers       (grows like n)
                                    1.18 / 1.38 / 1.59 / -         renormalised transfers
    The plain full-weighting restriction drops the share of a wall cell's residual that belongs to the coarse
    cell behind the wall, so a zero-mean residual next to a wall gets a net mass and the coarse levels answer
    with a smooth correction O(n) times too large;
  * 40 Jacobi sweeps on a 32 x 32 coarsest level leave its lowest modes nearly untouched (smallest eigenvalues
    of M A ~ 0.02); continuing the hierarchy to 4 x 4 fixes that -- but only together with the conservative
    transfers (coarser levels amplify the wall error further otherwise: 59 instead of 50 iterations in round 1);
  * either fix alone changes little (18 - 22 iterations at 512^2 .. 1024^2); both: 7 / 7 / 8 / 8 at
    512^2 / 1024^2 / 2048^2 / 4096^2 (3 + 3 sweeps; 9 with 2 + 2).

    python tools/studies/mg_transfer_study.py pcg 512 1024        # iteration counts, four variants
    python tools/studies/mg_transfer_study.py maxeig 64 128 256   # largest eigenvalue of M A, exact coarsest solve
    python tools/studies/mg_transfer_study.py blobs 256 512       # random blobby scene with interior solids
"""
import sys
import time

import numpy as np

sys.path.insert(0, '/root/repo/tools/studies'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import mgpcg_prototype as P

f32 = np.float32


def pcg_hist(L, b, inv_h2, prec, tol=1e-6, maxit=120):
    x = np.zeros_like(b); r = b.copy()
    rhs2 = float((b.astype(np.float64) ** 2).sum()); thr = tol * tol * rhs2
    z = prec(r); p = z.copy(); abs_new = float((r.astype(np.float64) * z).sum())
    hist = []
    for it in range(maxit):
        q = P.applyA(L, p, inv_h2)
        alpha = f32(abs_new / float((p.astype(np.float64) * q).sum()))
        x = (x + alpha * p).astype(f32); r = (r - alpha * q).astype(f32)
        r2 = float((r.astype(np.float64) ** 2).sum()); hist.append((r2 / rhs2) ** 0.5)
        if r2 < thr:
            break
        z = prec(r); abs_old = abs_new; abs_new = float((r.astype(np.float64) * z).sum())
        beta = f32(abs_new / abs_old); p = (z + beta * p).astype(f32)
    return len(hist), hist


VARIANTS = [("round 1: plain transfers, coarsest 32 x 32", dict(nmin=32, renorm=False)),
            ("plain transfers, coarsest 4 x 4", dict(nmin=4, renorm=False)),
            ("renormalised transfers, coarsest 32 x 32", dict(nmin=32, renorm=True)),
            ("renormalised transfers, coarsest 4 x 4 (now the default)", dict(nmin=4, renorm=True))]


def run_pcg(lab, u, v, n):
    dx = f32(1) / f32(n)
    b = P.rhs_from(lab, u, v, dx)
    L = P.make_level(lab); h2 = f32(1) / (dx * dx)
    for name, kw in VARIANTS:
        mg = P.MG(lab, dx, pre=3, post=3, **kw)
        t = time.time()
        it, hist = pcg_hist(L, b, h2, lambda r: mg.vcycle(r))
        print(f"n={n} {name}: {len(mg.levels)} levels, {it} iterations ({time.time() - t:.0f}s)", flush=True)
        print("    " + " ".join(f"{h:.1e}" for h in hist), flush=True)


class MGExact(P.MG):
    """Exact (dense) solve on the coarsest level."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        Lc = self.levels[-1]; idx = np.flatnonzero(Lc['liq'].ravel()); self.cidx = idx
        A = np.zeros((len(idx), len(idx)))
        for k_, i in enumerate(idx):
            x = np.zeros(Lc['liq'].size, dtype=f32); x[i] = 1.0
            A[:, k_] = P.applyA(Lc, x.reshape(Lc['liq'].shape), self.h2[-1]).ravel()[idx]
        self.Ainv = np.linalg.inv(A)

    def vcycle(self, b, l=0):
        if l == len(self.levels) - 1:
            x = np.zeros(b.size, dtype=f32); x[self.cidx] = self.Ainv @ b.ravel()[self.cidx]
            return x.reshape(b.shape)
        return super().vcycle(b, l)


def max_eig(n, renorm):
    lab, u, v = P.tank(n); dx = f32(1) / f32(n); L = P.make_level(lab); h2 = f32(1) / (dx * dx)
    mg = MGExact(lab, dx, pre=3, post=3, nmin=32, renorm=renorm)
    rng = np.random.default_rng(0); v = np.where(L['liq'], rng.standard_normal(lab.shape), 0.0).astype(f32)
    lam = 0.0
    for _ in range(60):
        w = mg.vcycle(P.applyA(L, v, h2)); lam = float((w * v).sum() / (v * v).sum()); v = (w / np.linalg.norm(w)).astype(f32)
    return len(mg.levels), lam


if __name__ == "__main__":
    what, sizes = sys.argv[1], [int(a) for a in sys.argv[2:]]
    for n in sizes:
        if what == "pcg":
            run_pcg(*P.tank(n), n)
        elif what == "blobs":
            import scenes
            rng = np.random.default_rng(3)
            lab = scenes.random_labels(n, n, rng, p_solid=0.03)
            run_pcg(lab, scenes.random_field(n, n, rng), scenes.random_field(n, n, rng), n)
        else:
            for renorm in (False, True):
                print(f"n={n} renorm={renorm}: levels, largest eigenvalue of M A = {max_eig(n, renorm)}", flush=True)
