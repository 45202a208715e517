#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native PIC/FLIP hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload picflip4096|sl1024|...] [--no-cpu-baseline]

A "step" is one full pass of the hot path (FluidSolver::stepPICFLIP, reference
src/FluidSolver.cpp:211-251) over the whole grid and particle set.

Workload (N = 1): BASELINE.json configs[2] -- 4096^2 PIC/FLIP (pic_ratio 0.02), 4 particles
per cell in a 15/16-full tank (6.29e7 particles), swirl initial velocity, dt = 0.01*64/4096,
density = dt (dt/rho = 1 as in the reference's example), CG run to a 1e-6 relative residual.
This is the configuration BASELINE.json's north_star quotes its single-GPU target on.

value   cell-updates/s = size_x*size_y*steps / device time, state resident in HBM
e2e     same metric through the C ABI with HOST buffers: every step uploads the particle set
        from pinned host memory, runs fsb_step, and downloads the particle set again
roofline  the CG iteration of the persistent one-sweep solve kernel k_cg_solve1 (>99 % of the step):
        achieved = bytes the kernel is DESIGNED to move per iteration (23 B x the cells of the tiles it
        sweeps) / (CUDA-event time of the CG loop on the library's stream /
        iterations), against MEASURED_PEAKS.json hbm_gbs.  frac_textbook does the same with the 45 B
        per cell of SURVEY.md 8(d) (a stored q = Ap, x touched every iteration): it exceeds 1 because
        the kernel does not move those bytes.  traffic / frac_dram: DRAM bytes per iteration from the
        committed ncu capture of the same workload (profiles/), null when there is none.
        stage_roofline gives the algorithmic-byte figure for every other stage.
scale_cg8192  on every --gpus N line: BASELINE.json configs[3], the 8192^2 tank pressure solve to 1e-6
        sharded over the N ranks (us per iteration, iterations/s), so that the north-star scaling
        config is in the driver's record.
cpu_baseline  the reference's own sources (oracle/_ref) on one host core on the SAME scene at the SAME
        size: every stage of one step once, the CG capped at 20 iterations; time-to-1e-6 is that
        per-iteration rate times the reference's own iteration count for this scene (17 821, from the
        72-minute golden run, tests/golden/config2_picflip4096.npz), labelled as extrapolated

Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TEXTBOOK_BYTES_CG_PER_CELL = 45.0  # bytes per cell per CG iteration with a stored q
# bytes per swept cell and iteration the solve kernels are designed to move
DESIGN_BYTES_CG_PER_CELL = {"one-sweep": 23.0, "persistent": 32.0, "graph": 32.0}
# DRAM bytes per CG iteration (dram__bytes_read.sum + dram__bytes_write.sum of the persistent launch /
# its iterations) from the committed ncu --set full captures, keyed "<workload>/<n_gpus>/<mode>"
NCU_TRAFFIC_FILE = os.path.join("profiles", "r02_cg_traffic.json")
# the reference's own iteration count to 1e-6 on the first step of the picflip4096 scene
# (tests/golden/config2_picflip4096.npz, oracle/_ref, 4343 s of CPU time)
REFERENCE_ITERS = {"picflip4096": 17821}

WORKLOADS = {
    # name: (n, step kind, pic_ratio, cg tol, cg cap, particles per cell side)
    "picflip4096": dict(n=4096, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2),
    "picflip2048": dict(n=2048, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2),
    "picflip1024": dict(n=1024, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2),
    # BASELINE.json configs[1]: 1024^2 semi-Lagrangian + CG, synthetic dam-break (the FluidSource box of
    # examples/simple.cpp:29 scaled to the grid, 6.25 particles per cell: 2.27e6 particles)
    "sl1024": dict(n=1024, kind="sl", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2, scene="dam"),
    "sl4096": dict(n=4096, kind="sl", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2, scene="dam"),
    # BASELINE.json configs[4]: 16384^2 full PIC/FLIP step.  1.0e9 particles: the tank is filled on
    # the device by fsb_emit_source (the reference's FluidSource lattice, src/FluidDomain.cpp:34-50,
    # at delta/2 spacing starting a quarter cell in: 4 particles per cell at the stratum centres,
    # at rest), there is no host copy of the set and no e2e leg
    "picflip16384": dict(n=16384, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2,
                         emit=True),
    "picflip8192e": dict(n=8192, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2,
                         emit=True),
    "picflip256": dict(n=256, kind="picflip", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=2),
    # BASELINE.json configs[3]: pressure Poisson CG only, tank labels, swirl + gravity velocities
    "cg8192": dict(n=8192, kind="cg", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=0),
    "cg4096": dict(n=4096, kind="cg", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=0),
    "cg1024": dict(n=1024, kind="cg", pic_ratio=0.02, tol=1e-6, cap=400000, per_side=0),
    # one rank's share of cg8192 on 8 / 4 GPUs as a single-GPU workload (8192 columns x 1024 / 2048 rows,
    # same cell size): the per-rank sweep without any cross-GPU traffic (tuning aid, capped)
    "cg8192slab8": dict(n=8192, ny=1024, kind="cg", pic_ratio=0.02, tol=1e-6, cap=2000, per_side=0),
    "cg8192slab4": dict(n=8192, ny=2048, kind="cg", pic_ratio=0.02, tol=1e-6, cap=2000, per_side=0),
}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def tank_particles(n, per_side, seed=1234):
    import scenes
    return scenes.tank_particles(n, np.random.default_rng(seed), per_side)


def tank_fields(n, ny=None):
    """CG-only workloads: labels of the tank scene (SOLID border, LIQUID below 15/16, AIR cap) and a
    velocity field = analytic swirl + one gravity kick, sampled at the MAC face positions."""
    dx = 1.0 / n
    ny = n if ny is None else ny
    lab = np.full((ny, n), 1, dtype=np.uint8)
    lab[1:int(15.0 / 16.0 * ny), 1:-1] = 0
    lab[0, :] = lab[-1, :] = 2
    lab[:, 0] = lab[:, -1] = 2
    i, j = np.arange(n, dtype=np.float64), np.arange(ny, dtype=np.float64)
    xu, yu = (i * dx)[None, :], ((j + 0.5) * dx)[:, None]
    xv, yv = ((i + 0.5) * dx)[None, :], (j * dx)[:, None]
    u = (np.sin(np.pi * xu) * np.cos(np.pi * yu)).astype(np.float32)
    v = (-np.cos(np.pi * xv) * np.sin(np.pi * yv) - 9.82 * 0.01 * 64.0 / n).astype(np.float32)
    return lab, u, v


def step_kind(mod, name):
    return {"picflip": mod.STEP_PICFLIP, "sl": mod.STEP_SL, "flip": mod.STEP_FLIP,
            "pic": mod.STEP_PIC}[name]


# --------------------------------------------------------------------------- reference arm --
def run_cpu_reference(wl, name, steps, warmup, gpu_iters_hint=None, parts=None, budget_s=150.0):
    """Times the reference's CPU implementation (one core: the reference is single-threaded,
    src/FluidSolver.cpp:3,418-420) on a bounded sample of the SAME scene at the SAME size.

    Sample step 0: every stage of one step, once, stage by stage, with the CG capped at 20
    iterations (4096^2: about 28 s).  Further sample steps (the reference arm is asked for K of
    them): the pressure system assembled again and iterated for a few iterations, as many as the
    time budget allows per step.  Nothing is scaled in cells.  The converged-step time is
        t(all stages but the CG loop) + iterations_to_tol * t(per CG iteration)
    with iterations_to_tol the reference's OWN count where it is known (REFERENCE_ITERS: the golden
    run of this scene), else the GPU's count for the same scene -- labelled as extrapolated: only
    the iteration count is, every time in the formula was measured at full size."""
    import oracle_lib as ol
    import scenes
    kind_name = "reference" if ol.available("fsr") else "port"
    lib = ol.OracleLib("fsr" if kind_name == "reference" else "fso")
    n = wl["n"]
    dt = np.float32(0.01 * 64.0 / n)
    s = lib.sim(n, n, 1.0, 1.0, float(dt), wl["pic_ratio"])
    grav = float(np.float32(-9.82))
    if wl["kind"] == "cg":
        lab, u0, v0 = tank_fields(n)
        s.set_cell_types(lab)
        n_part = 0
    else:
        if parts is None:
            parts = scenes.tank_particles(n, np.random.default_rng(1234), wl["per_side"])
        s.set_particles(parts)
        n_part = parts.shape[0]
    cap0 = 20
    s.set_cg(cap0, wl["tol"])
    t_start = time.perf_counter()
    t0 = time.perf_counter()
    if wl["kind"] == "cg":
        s.set_grid(ol.U_FRONT, u0); s.set_grid(ol.V_FRONT, v0)
        t0 = time.perf_counter()
        t1 = t0; s.pressure_solve(float(dt), float(dt)); t2 = time.perf_counter()
    elif wl["kind"] == "sl":
        s.classify_cells(); s.advect_velocity_sl(float(dt)); s.add_acceleration(0.0, grav, float(dt))
        s.enforce_dirichlet()
        t1 = time.perf_counter(); s.pressure_solve(float(dt), float(dt)); t2 = time.perf_counter()
        s.enforce_dirichlet(); s.advect_particles_grid(float(dt))
    else:
        s.classify_cells(); s.p2g_spread(); s.save_previous()
        s.add_acceleration(0.0, grav, float(dt)); s.enforce_dirichlet(); s.extend_velocity(2)
        t1 = time.perf_counter(); s.pressure_solve(float(dt), float(dt)); t2 = time.perf_counter()
        s.enforce_dirichlet(); s.update_diff(); s.g2p(ol.G2P_PICFLIP, wl["pic_ratio"])
        s.advect_particles(float(dt), True)
    t3 = time.perf_counter()
    step_times = [t3 - t0]
    solve_samples = [(t2 - t1, max(1, s.cg_info()[0]))]
    # further sample steps: assembly + a few iterations of the next step's system
    per_step = max(0.0, budget_s - (t3 - t_start)) / max(1, steps - 1)
    for k in range(1, max(1, steps)):
        t_it = solve_samples[0][0] / (solve_samples[0][1] + 24.0)  # rough: assembly ~ 24 iterations
        cap = int(max(1, min(cap0, per_step / max(t_it, 1e-9) - 24)))
        s.set_cg(cap, wl["tol"])
        if wl["kind"] == "cg":
            s.set_grid(ol.U_FRONT, u0); s.set_grid(ol.V_FRONT, v0)
        ta = time.perf_counter(); s.pressure_solve(float(dt), float(dt)); tb = time.perf_counter()
        solve_samples.append((tb - ta, max(1, s.cg_info()[0])))
        step_times.append(tb - ta)
    # per-iteration time and assembly + patch time from the solves (least squares over the samples,
    # or a zero-iteration solve when there is only one)
    if len(set(c for _, c in solve_samples)) >= 2:
        A = np.array([[1.0, c] for _, c in solve_samples]); y = np.array([t for t, _ in solve_samples])
        (solve0, t_iter), *_ = np.linalg.lstsq(A, y, rcond=None)
        solve0, t_iter = float(max(solve0, 0.0)), float(max(t_iter, 1e-9))
    else:
        s.set_cg(0, wl["tol"])
        if wl["kind"] == "cg":
            s.set_grid(ol.U_FRONT, u0); s.set_grid(ol.V_FRONT, v0)
        ta = time.perf_counter(); s.pressure_solve(float(dt), float(dt)); solve0 = time.perf_counter() - ta
        t_iter = max(solve_samples[0][0] - solve0, 1e-9) / solve_samples[0][1]
    t_non_cg = (t3 - t0) - (t2 - t1) + solve0
    iters_to_tol = REFERENCE_ITERS.get(name) or gpu_iters_hint or int(3.2 * n)
    iters_src = ("the reference's own count for this scene (golden run)" if name in REFERENCE_ITERS
                 else "the GPU's count for this scene" if gpu_iters_hint else "the O(N) law 3.2 N")
    step_time_full = t_non_cg + iters_to_tol * t_iter
    what = "pressure solve" if wl["kind"] == "cg" else f"{wl['kind']} step"
    sample = (f"{kind_name} build, 1 thread, SAME scene and size ({n}^2, {n_part} particles): one {what} run "
              f"stage by stage with the CG capped at {cap0} iterations ({step_times[0]:.1f} s)"
              + (f", then {len(step_times) - 1} more sample steps = the pressure system assembled and iterated "
                 f"{solve_samples[-1][1]} times" if len(step_times) > 1 else "")
              + f": all stages but the CG loop {t_non_cg:.2f} s, {t_iter * 1e3:.1f} ms per CG iteration; "
              f"converged-step time = {t_non_cg:.1f} s + {iters_to_tol} x {t_iter * 1e3:.1f} ms = {step_time_full:.0f} s "
              f"with the iteration count taken from {iters_src} (time-to-1e-6 is EXTRAPOLATED in the "
              f"iteration count only; nothing is scaled in cells)")
    return {"value": n * n / step_time_full, "unit": "cell-updates/s", "cores": 1, "kind": kind_name,
            "sample": sample, "host_cores_available": os.cpu_count(),
            "cg_iters_per_s": 1.0 / t_iter, "sample_step_s": step_times,
            "non_cg_s": t_non_cg, "converged_step_s_extrapolated": step_time_full}


# ------------------------------------------------------------------------------- our arm --
def scale_cg8192(capi, sharding, dist, torch, local_rank, world, log):
    """The 8192^2 tank pressure solve (BASELINE.json configs[3]) to 1e-6, rows split over the ranks:
    one capped warm-up solve, one timed solve.  Device time of the CG loop, max over ranks."""
    n = 8192
    dt = float(np.float32(0.01 * 64.0 / n))
    lab, u0, v0 = tank_fields(n)
    g = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local_rank)
    g.set_cell_types(lab)
    del lab
    if world > 1:
        sharding.connect(g, dist, torch.device("cuda", local_rank))
    out = {}
    for cap, timed in ((200, False), (400000, True)):
        g.set_grid(capi.U_FRONT, u0); g.set_grid(capi.V_FRONT, v0)
        g.set_cg(cap, 1e-6)
        g.synchronize()
        if world > 1:
            dist.barrier()
        g.profile_enable(True)
        g.profile_read()
        g.pressure_solve(dt, dt)
        g.synchronize()
        prof = g.profile_read()
        g.profile_enable(False)
        if not timed:
            continue
        cg_ms = prof["cg"][0]
        if world > 1:
            t = torch.tensor([cg_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cg_ms = float(t.item())
        it, relres = g.cg_info()
        mode = {1: "graph", 4: "one-sweep"}.get(g.cg_launch_mode(), "persistent")
        out = {"workload": "8192^2 pressure Poisson solve (tank labels, swirl + gravity field), Jacobi-PCG to "
                           "1e-06, row slabs over the ranks", "n_gpus": world, "iterations": int(it),
               "relres": float(relres), "cg_ms": cg_ms, "us_per_iteration": 1e3 * cg_ms / max(it, 1),
               "cg_iters_per_s": it / (cg_ms * 1e-3) if cg_ms else None, "mode": mode,
               "swept_cells_per_rank": g.cg_swept_cells()}
        log(f"scale_cg8192: {out}")
    if world > 1:
        g.shard_disconnect()
        dist.barrier()
    g.close()
    return out


def small_configs(capi, local_rank, log):
    """BASELINE.json configs[0] and configs[1] on the device, as side measurements of the one-GPU line:
    the examples/simple.cpp scene (64^2, one source, dt 0.01, 100 x stepPICFLIP with the reference's own
    CG settings: cap 100, tolerance FLT_EPSILON) and the 1024^2 semi-Lagrangian dam-break with the CG
    run to 1e-6.  Parity of both against golden vectors of the compiled reference:
    tests/test_gpu_goldens_big.py."""
    import scenes
    out = {}
    g = capi.Sim(64, 64, 1.0, 1.0, 0.01, 0.05, device=local_rank)
    g.emit_source(*scenes.dam_break_args(64))
    for _ in range(5):
        g.step(capi.STEP_PICFLIP, 0.01)
    g.synchronize()
    g.timer_start()
    it = 0
    for _ in range(100):
        g.step(capi.STEP_PICFLIP, 0.01)
        it += g.cg_info()[0]
    ms = g.timer_stop()
    out["config0_simple64"] = {"workload": "examples/simple.cpp scene: 64^2, 7800 particles, PIC/FLIP 0.05, dt 0.01, "
                                           "100 steps, CG cap 100 / tolerance FLT_EPSILON (the reference's defaults)",
                               "ms_per_step": ms / 100, "cell_updates_per_s": 64 * 64 * 100 / (ms * 1e-3),
                               "cg_iters_per_step": it / 100}
    g.close()
    n = 1024
    dt = float(np.float32(0.01 * 64.0 / n))
    g = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local_rank)
    g.set_cg(400000, 1e-6)
    n_part = g.emit_source(*scenes.dam_break_args(n))
    for _ in range(3):
        g.step(capi.STEP_SL, dt)
    g.synchronize()
    g.profile_enable(True); g.profile_read()
    g.timer_start()
    it = 0
    for _ in range(5):
        g.step(capi.STEP_SL, dt)
        it += g.cg_info()[0]
    ms = g.timer_stop()
    prof = g.profile_read()
    out["config1_sl1024"] = {"workload": "1024^2 semi-Lagrangian step (RK3 back-trace, deterministic gather) + CG to "
                                         "1e-06, dam-break scene", "particles": int(n_part),
                             "ms_per_step": ms / 5, "cell_updates_per_s": n * n * 5 / (ms * 1e-3),
                             "cg_iters_per_step": it / 5, "cg_us_per_iteration": 1e3 * prof["cg"][0] / max(it, 1),
                             "stage_ms_per_step": {k: round(v[0] / 5, 4) for k, v in prof.items() if v[0] > 0}}
    g.close()
    log(f"small configs: {out}")
    return out


def config4_slabs(capi, sharding, dist, torch, local_rank, world, log, n=16384):
    """BASELINE.json configs[4]: the 16384^2 PIC/FLIP step on `world` GPUs -- particles partitioned by
    row slab (ghost rows and migration over NCCL, device buffers), the pressure CG sharded over the same
    slabs, the grid stages replicated.  The tank is filled on the device by fsb_emit_source (4 particles
    per cell, at rest); every rank emits the whole lattice, keeps its slab and frees nothing else.
    One warm-up step with the CG capped, one timed step to 1e-6."""
    dt = float(np.float32(0.01 * 64.0 / n))
    d = 1.0 / n
    g = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local_rank)
    n_all = g.emit_source(1.25 * d, 1.0 - d, 1.25 * d, 15.0 / 16.0, 1.25 * d, 1.25 * d, 0.0, 0.0)
    dev = torch.device("cuda", local_rank)
    slabs = sharding.DistSlabs(g, dist, dev)
    slabs.distribute()
    n_own = g.num_particles()
    sharding.connect(g, dist, dev)
    log(f"config4: {n_all} particles, {n_own} on this rank, rows {slabs.lo}..{slabs.hi}")
    g.set_cg(100, 1e-6)
    slabs.step(capi.STEP_PICFLIP, dt)
    g.synchronize(); dist.barrier()
    g.set_cg(400000, 1e-6)
    g.profile_enable(True); g.profile_read()
    t0 = time.perf_counter()
    g.timer_start()
    moved = slabs.step(capi.STEP_PICFLIP, dt)
    ms = g.timer_stop()
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    prof = g.profile_read(); g.profile_enable(False)
    it, relres = g.cg_info()
    t = torch.tensor([ms, prof["cg"][0], wall * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, cg_ms, wall_ms = (float(v) for v in t.tolist())
    out = {"workload": f"{n}^2 picflip full step on {world} GPUs: particle slabs (NCCL, device buffers) + sharded "
                       f"Jacobi-PCG to 1e-06, tank scene 4 particles/cell (device-emitted lattice, at rest)",
           "n_gpus": world, "particles": int(n_all), "particles_per_rank": int(n_own),
           "ms_per_step": ms, "wall_ms_per_step": wall_ms, "cell_updates_per_s": n * n / (wall_ms * 1e-3),
           "cg_iterations": int(it), "cg_relres": float(relres), "cg_us_per_iteration": 1e3 * cg_ms / max(it, 1),
           "migrated_from_this_rank": int(moved),
           "stage_ms": {k: round(v[0], 3) for k, v in prof.items()}}
    log(f"config4: {out}")
    g.shard_disconnect()
    dist.barrier()
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="picflip4096", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-optin", action="store_true", help="skip the opt-in multigrid side measurement")
    ap.add_argument("--no-scale", action="store_true", help="skip the 8192^2 CG (and, on 8 GPUs, the 16384^2 step) "
                                                            "side measurements")
    ap.add_argument("--config4", type=int, default=0, metavar="N",
                    help="run the particle-slab side measurement at N^2 on any multi-GPU world (debug)")
    ap.add_argument("--cg-cap", type=int, default=None, help="override the CG iteration cap (debug)")
    ap.add_argument("--precond", default="jacobi", choices=["jacobi", "mg"],
                    help="jacobi: the reference's preconditioner (headline); mg: the opt-in multigrid "
                         "V-cycle (same system and stopping rule, not the reference's algorithm)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.cg_cap is not None:
        wl["cap"] = args.cg_cap
    n = wl["n"]
    t_start = time.perf_counter()

    def log(msg):
        if args.verbose:
            print(f"[bench +{time.perf_counter() - t_start:7.2f}s] {msg}", file=sys.stderr, flush=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if wl["kind"] == "cg":
        cfg_name = (f"{n}^2 pressure Poisson solve only (tank labels, swirl + gravity field), "
                    f"Jacobi-PCG to {wl['tol']:g} relative residual")
    else:
        scene = ("dam-break scene 6.25 particles/cell" if wl.get("scene") == "dam"
                 else f"tank scene {wl['per_side']**2} particles/cell")
        cfg_name = (f"{n}^2 {wl['kind']} full step, {scene}"
                    f"{' (device-emitted lattice, at rest)' if wl.get('emit') else ''}, "
                    f"pic_ratio {wl['pic_ratio']}, CG to {wl['tol']:g} relative residual")

    if args.impl == "reference":
        if rank != 0:
            return 0
        hint = None
        hp = os.path.join(ROOT, "gpurun_out", "last_gpu_iters.json")
        if os.path.exists(hp):
            try:
                hint = json.load(open(hp)).get(args.workload)
            except Exception:
                hint = None
        cb = run_cpu_reference(wl, args.workload, args.steps, args.warmup, hint)
        # ms_per_step: the sample steps as they were actually timed (so that steps x ms_per_step is the
        # timed region of THIS run); value: the converged step, extrapolated in the iteration count
        timed = cb.pop("sample_step_s")
        line = {"impl": "reference", "metric": "cell_updates_per_s", "value": cb["value"],
                "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(timed) / max(1, len(timed)),
                "ms_per_converged_step_extrapolated": 1e3 * cb["converged_step_s_extrapolated"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": cfg_name},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # stdout carries exactly ONE JSON line: NCCL (C level) prints its version banner there when
    # NCCL_DEBUG is set on the box, so fd 1 is pointed at stderr for the whole run and the line is
    # written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from fluid_simulation_b200 import capi, sharding

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dt = float(np.float32(0.01 * 64.0 / n))
    ny = wl.get("ny", n)
    sim = capi.Sim(n, ny, 1.0, float(ny) / float(n), dt, wl["pic_ratio"], device=local_rank)
    sim.set_cg(wl["cap"], wl["tol"])
    if args.precond == "mg":
        sim.set_preconditioner(capi.PRECOND_MULTIGRID)
    cg_only = wl["kind"] == "cg"
    if cg_only:
        lab, u0, v0 = tank_fields(n, ny)
        hu = torch.empty((ny, n), dtype=torch.float32, pin_memory=True); hu.numpy()[:] = u0
        hv = torch.empty((ny, n), dtype=torch.float32, pin_memory=True); hv.numpy()[:] = v0
        hp = torch.empty((ny, n), dtype=torch.float32, pin_memory=True)
        sim.set_cell_types(lab)
        n_part = 0
        del u0, v0
        log("fields ready")
    elif wl.get("emit"):
        d = 1.0 / n
        n_part = sim.emit_source(1.25 * d, 1.0 - d, 1.25 * d, 15.0 / 16.0, 1.25 * d, 1.25 * d, 0.0, 0.0)
        host = None
        args.no_e2e = True
        kind = step_kind(capi, wl["kind"])
        log(f"scene emitted on the device: {n_part} particles")
    else:
        if wl.get("scene") == "dam":
            import scenes
            sim.emit_source(*scenes.dam_break_args(n))
            parts = sim.get_particles()
        else:
            parts = tank_particles(n, wl["per_side"])
        n_part = parts.shape[0]
        host = torch.empty((n_part, 4), dtype=torch.float32, pin_memory=True)
        host.numpy()[:] = parts
        del parts
        log(f"scene ready: {n_part} particles")
        sim.set_particles_ptr(host.data_ptr(), n_part)
        kind = step_kind(capi, wl["kind"])
    sim.synchronize()
    rows = (0, n)
    if world > 1:
        # every stage but the CG runs replicated (deterministic kernels, identical on all ranks);
        # the CG iterates on row slabs with peer-memory halos and mailbox reductions
        rows = sharding.connect(sim, dist, torch.device("cuda", local_rank))
    log(f"state uploaded, rows {rows}")

    def barrier():
        torch.cuda.synchronize()
        sim.synchronize()
        if world > 1:
            dist.barrier()

    def one_step():
        """One pass of the hot path with state resident in HBM."""
        if cg_only:
            # the solve consumes its input (it projects the field): restore it from HBM-resident
            # copies is not possible through the ABI, so re-upload; only the solve is timed below
            sim.set_grid_ptr(capi.U_FRONT, hu.data_ptr())
            sim.set_grid_ptr(capi.V_FRONT, hv.data_ptr())
            sim.pressure_solve(dt, dt)
        else:
            sim.step(kind, dt)

    # ---- warm-up
    for _ in range(args.warmup):
        one_step()
        sim.synchronize()
        log(f"warm-up step done, cg {sim.cg_info()}")
    barrier()

    # ---- timed region: K steps, state resident in HBM
    sim.profile_enable(True)
    sim.profile_read()
    launches0 = sim.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    iters_total = 0
    barrier()
    sim.timer_start()
    for _ in range(args.steps):
        one_step()
        iters_total += sim.cg_info()[0]
        log(f"timed step done, cg {sim.cg_info()}")
    ms = sim.timer_stop()
    barrier()
    clock_info = clocks.stop()
    launches = sim.launch_count() - launches0
    prof = sim.profile_read()
    sim.profile_enable(False)
    relres = sim.cg_info()[1]
    if cg_only:
        # device time of the solves only (build + CG loop + patch), not the input re-upload
        ms = prof["rhs"][0] + prof["cg"][0] + prof["patch"][0]
    if world > 1:
        t = torch.tensor([ms, prof["cg"][0]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, cg_ms_max = float(t[0].item()), float(t[1].item())
        prof["cg"] = (cg_ms_max, prof["cg"][1])
    ms_per_step = ms / args.steps
    # N > 1 is STRONG scaling of the same workload: total work fixed, CG rows split over N GPUs
    value = n * ny * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers in and out every step
    e2e = None
    if not args.no_e2e:
        barrier()
        e2e_iters = 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            if cg_only:
                sim.set_grid_ptr(capi.U_FRONT, hu.data_ptr())
                sim.set_grid_ptr(capi.V_FRONT, hv.data_ptr())
                sim.pressure_solve(dt, dt)
                sim.get_pressure_ptr(hp.data_ptr())
            else:
                sim.set_particles_ptr(host.data_ptr(), n_part)
                sim.step(kind, dt)
                sim.get_particles_ptr(host.data_ptr())
            e2e_iters += sim.cg_info()[0]
        barrier()
        t1 = time.perf_counter()
        et = t1 - t0
        if world > 1:
            t = torch.tensor([et], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            et = float(t.item())
        h2d = 2 * n * ny * 4 if cg_only else n_part * 16
        d2h = n * ny * 4 if cg_only else n_part * 16
        e2e = {"value": n * ny * args.steps / et, "unit": "cell-updates/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               # the e2e steps continue the same simulation: later steps, other iteration counts
               "cg_iters_per_step": e2e_iters / args.steps}

    cg_mode = {1: "graph", 3: "multigrid", 4: "one-sweep"}.get(sim.cg_launch_mode(), "persistent")
    swept_cells = sim.cg_swept_cells()  # per rank, per iteration (cells of the tiles the sweeps visit)

    # ---- side measurement, not the headline: the same steps with the opt-in multigrid preconditioner
    optin = None
    if world == 1 and args.precond == "jacobi" and not args.no_optin:
        try:
            sim.set_preconditioner(capi.PRECOND_MULTIGRID)
            one_step()  # builds the level hierarchy
            barrier()
            it_mg = 0
            sim.timer_start()
            for _ in range(args.steps):
                one_step()
                it_mg += sim.cg_info()[0]
            ms_mg = sim.timer_stop()
            mode_mg = sim.cg_launch_mode()
            optin = {"what": "same workload with fsb_set_preconditioner(FSB_PRECOND_MULTIGRID): same system "
                             "and stopping rule, NOT the reference's Jacobi-PCG (no parity claim, not the "
                             "headline)",
                     "ms_per_step": ms_mg / args.steps,
                     "cell_updates_per_s": n * ny * args.steps / (ms_mg * 1e-3),
                     "cg_iters_per_step": it_mg / args.steps, "relres": sim.cg_info()[1],
                     "multigrid_used": mode_mg == 3}
        except Exception as e:  # a side measurement must never cost the headline line
            optin = {"error": str(e)[:200]}
        finally:
            try:
                sim.set_preconditioner(capi.PRECOND_JACOBI)
            except Exception:
                pass
    # ---- side measurement on every line: BASELINE.json configs[3], the 8192^2 pressure solve to 1e-6
    # sharded over the `world` ranks (north_star: ">= 6x strong scaling from 1 to 8 GPUs")
    scale = None
    if not args.no_scale and args.workload != "cg8192" and args.precond == "jacobi":
        try:
            scale = scale_cg8192(capi, sharding, dist, torch, local_rank, world, log)
        except Exception as e:  # a side measurement must never cost the headline line
            scale = {"error": str(e)[:300]}
    # ---- side measurements on the one-GPU line: BASELINE.json configs[0] and configs[1]
    small = None
    if world == 1 and not args.no_scale and args.precond == "jacobi":
        try:
            small = small_configs(capi, local_rank, log)
        except Exception as e:
            small = {"error": str(e)[:300]}
    # ---- side measurement on the 8-GPU line: BASELINE.json configs[4] (16384^2 on the whole box)
    config4 = None
    if (world == 8 or args.config4) and world > 1 and not args.no_scale and args.precond == "jacobi":
        sim.shard_disconnect()
        sim.close()  # its 5 GB of particles and grids are not needed any more
        try:
            config4 = config4_slabs(capi, sharding, dist, torch, local_rank, world, log,
                                    n=args.config4 if args.config4 else 16384)
        except Exception as e:
            config4 = {"error": str(e)[:300]}
        dist.barrier()
        dist.destroy_process_group()
    elif world > 1:
        sim.shard_disconnect()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peak, peak_src = measured_peak_gbs()
    cg_ms, _ = prof["cg"]
    it_ms = cg_ms / max(iters_total, 1)
    # per GPU: each rank sweeps the active tiles of its 1/world of the rows per iteration
    design_b = DESIGN_BYTES_CG_PER_CELL.get(cg_mode)
    if not swept_cells:
        swept_cells = n * ny // world
    design_bytes = (design_b or TEXTBOOK_BYTES_CG_PER_CELL) * swept_cells
    textbook_bytes = TEXTBOOK_BYTES_CG_PER_CELL * n * ny / world
    achieved = design_bytes / (it_ms * 1e-3) / 1e9 if iters_total else 0.0
    achieved_textbook = textbook_bytes / (it_ms * 1e-3) / 1e9 if iters_total else 0.0
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, NCU_TRAFFIC_FILE)))
        ent = tr.get(f"{args.workload}/{world}/{cg_mode}") or (
            tr.get(f"cg{n}/{world}/{cg_mode}") if wl["kind"] != "cg" else None)
        if ent:
            traffic, traffic_src = float(ent["dram_bytes_per_iteration"]), ent["source"]
    except Exception:
        pass
    stages = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}
    # every stage against the HBM roofline: algorithmic bytes of SURVEY.md 8(d) (C cells, P
    # particles; the cell sort is overhead and has no algorithmic bytes) / CUDA-event time
    C, P = float(n * ny), float(n_part)
    # classification rides on the cell sort's counting pass (one read of the particle set for both),
    # so the two are timed together; the sort itself is overhead of the deterministic P2G
    alg = {"classify+sort": 8 * P + C, "p2g": 16 * P + 8 * C, "extend": 17 * C, "rhs": 13 * C,
           "grid_pre+patch": (17 + 21) * C if wl["kind"] not in ("cg", "sl") else 21 * C + 17 * C,
           "g2p": 32 * P + 17 * C, "advect_sl": 17 * C, "advect_part": 16 * P + 8 * C}
    stage_roofline = {}
    for name, nbytes in alg.items():
        t_ms = sum(stages.get(k, 0.0) for k in name.split("+"))
        if t_ms > 0:
            gbs = nbytes / (t_ms * 1e-3) / 1e9
            stage_roofline[name] = {"ms": round(t_ms, 4), "alg_gb": round(nbytes / 1e9, 4),
                                    "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
    kernel = {"multigrid": "OPT-IN multigrid-preconditioned CG iteration (one V-cycle + four sweeps; NOT the "
                           "reference's Jacobi iteration: no byte accounting applies)",
              "one-sweep": "k_cg_solve1 (persistent cooperative kernel, the whole solve is ONE launch; ONE sweep "
                           "and one reduction point per CG iteration): figures are per CG iteration",
              "persistent": "k_cg_solve (persistent cooperative kernel, the whole solve is ONE launch): "
                            "figures are per CG iteration = one direction sweep + one update sweep",
              "graph": "CG iteration = k_cg_direction + k_cg_update (2 launches in a CUDA graph)"}[cg_mode]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "what": (f"bytes the kernel is designed to move: {design_b:g} B x the {swept_cells} cells of the "
                         f"tiles it sweeps per iteration on one GPU" if design_b else "no byte accounting"),
                "achieved_textbook": achieved_textbook, "frac_textbook": achieved_textbook / peak,
                "what_textbook": "45 B x all cells (SURVEY.md 8d: stored q = Ap, x touched every iteration); "
                                 "above 1 because the kernel does not move those bytes",
                "traffic": traffic, "traffic_source": traffic_src,
                "frac_dram": (traffic / (it_ms * 1e-3) / 1e9 / peak) if traffic and iters_total else None,
                "kernel": kernel + (", per GPU, slab halos + sums over peer memory" if world > 1 else ""),
                "design_bytes_per_iteration": design_bytes,
                "textbook_bytes_per_iteration": textbook_bytes,
                "swept_cells_per_iteration": swept_cells,
                "iterations_per_launch": (iters_total / args.steps) if cg_mode != "graph" else 0.5,
                "avg_iteration_us": it_ms * 1e3, "peak_source": peak_src,
                "cg_share_of_step": cg_ms / ms if ms else None}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    try:
        json.dump({args.workload: int(round(iters_total / args.steps))},
                  open(os.path.join(ROOT, "gpurun_out", "last_gpu_iters.json"), "w"))
    except Exception:
        pass

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = run_cpu_reference(wl, args.workload, 1, 0, int(round(iters_total / args.steps)))
            cpu.pop("sample_step_s", None)
        except Exception as e:  # the reported baseline must never cost the headline line
            cpu = {"error": str(e)[:200]}

    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg_name, "particles": int(n_part), "dt": dt,
                   "preconditioner": ("multigrid V-cycle (opt-in)" if args.precond == "mg"
                                      else "Jacobi (the reference's)"),
                   "l2": (f"inputs larger than L2 ({n_part * 16 / 1e9:.1f} GB particles, "
                          f"{n * n * 4 / 1e6:.0f} MB per grid)")
                   if n >= 4096 else "working set may fit L2: latency-bound, see DESIGN.md",
                   "parallelism": (f"CG sharded over {world} row slabs (peer-memory halo stores + one mailbox "
                                   f"reduction per iteration over NVLink), other stages replicated")
                   if world > 1 else "single GPU"},
        "cg_iters_per_step": iters_total / args.steps,
        "cg_iters_per_s": iters_total / (cg_ms * 1e-3) if cg_ms else None,
        "cg_relres": relres,
        "stage_ms_per_step": stages,
        "stage_roofline": stage_roofline,
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "e2e": e2e,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "optin_multigrid": optin,
        "scale_cg8192": scale,
        "config4_picflip16384": config4,
        "small_configs": small,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
