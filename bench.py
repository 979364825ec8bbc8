#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native NBodySimulator.jl acceleration path.

Metric (BASELINE.json): gravity pair-interactions/s on the all-pairs Plummer sphere, 262,144 bodies,
fp64 (configs[1]) + LJ argon atom-steps/s (configs[2], the "lj" object of the same line).  One "step" =
one velocity-Verlet step of the whole system = one pass of the hot path (N(N-1) ordered pair
interactions) plus the O(N) update kernels.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

N > 1: every rank owns one context; torch.distributed (NCCL) only bootstraps the group (CUDA IPC handles), times
and gathers -- the step loop and the RHS drop-in run inside libnbody_b200 with all exchanges done by the
kernels over NVLink peer memory (nbx_group_*, include/nbody_b200.h).

Prints ONE JSON line on rank 0.  `value` is device-timed with the state resident in HBM (CUDA events
on the launching stream, L2 flushed between timed steps, max over ranks); `e2e` goes through the
public RHS drop-in nbx_accel with HOST buffers, copies inside the timed region; `roofline` compares
the dominant kernel with a DFMA peak measured live on the same device; `cpu_baseline` times the CPU
oracle (the restatement of the reference's Julia loops; Julia itself is not installed) on a bounded
sample on the host cores; `parity` compares a subsample of the accelerations the timed run left on
the device with that oracle, at every N.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 20.0          # SURVEY.md 8(d) convention for gravity
N_GRAVITY = 262144
DT_GRAVITY = 1.0e-4
LJ_CELLS = 64                 # FCC 64^3 x 4 = 1,048,576 atoms
METRIC = "gravity pair-interactions/s (all-pairs Plummer sphere, 262,144 bodies, fp64)"
LJ_METRIC = "LJ argon atom-steps/s (1,048,576 atoms, cell list, Berendsen, velocity Verlet)"
PARITY_TARGETS = 256


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads():
    return max(1, os.cpu_count() or 1)


def gravity_config():
    """Identical in both arms (the driver compares the dicts)."""
    return {"workload": "gravity_plummer_262144", "n_bodies": N_GRAVITY, "masses": "equal (1/N)", "softening": 0.0,
            "step": "one acceleration evaluation of all bodies (inside a velocity-Verlet step on the GPU arm)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (the timed region); when it is shorter than the sampling
        period and holds none, of all samples since start() (warm-up + timed steps: the same load), and says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.06)]
        window = "timed steps"
        if not rows:
            rows = [r for _, r in self.rows]
            window = "warm-up + timed steps (the timed region is shorter than the sampling period)"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_gravity_sample(u, ms, ntargets, nthreads, seed=1):
    """Times orc_accel_targets on `ntargets` targets against all sources; returns (pairs/s, seconds)."""
    from oracle import nbody_oracle as orc

    n = u.shape[1]
    s = orc.System(ms, gravity=dict(G=1.0))
    targets = np.random.Generator(np.random.Philox(seed)).choice(n, ntargets, replace=False)
    t0 = time.perf_counter()
    s.accel_targets(u, targets, nthreads)
    dt = time.perf_counter() - t0
    return ntargets * (n - 1) / dt, dt


def cpu_lj_sample(w, u, ntargets, nthreads, seed=1):
    """The reference's LJ loop has no neighbour structure (src/basic_potentials.jl:240-272: N-1 distance checks per atom):
    times `ntargets` targets against all N sources; returns (atom-steps/s of a full RHS at that rate, seconds)."""
    from oracle import nbody_oracle as orc

    n = u.shape[1]
    s = orc.System(w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    targets = np.random.Generator(np.random.Philox(seed)).choice(n, ntargets, replace=False)
    t0 = time.perf_counter()
    s.accel_targets(u, targets, nthreads)
    dt = time.perf_counter() - t0
    return ntargets / dt, dt   # atoms whose acceleration is complete per second == atom-steps/s of the force loop


def lj_inputs(cells=LJ_CELLS):
    import nbody_b200.workloads as wl

    w = wl.fcc_argon_reduced(cells)
    rng = np.random.Generator(np.random.Philox(2))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    return w, u


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia is absent from the image).  The reference is
    single-threaded: `value` is the 1-thread rate; the OpenMP-over-targets figure on all host
    threads is reported under "all_threads" and labelled as not the reference."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import nbody_b200.workloads as wl
    from oracle import nbody_oracle as orc

    orc.build()
    n = N_GRAVITY
    u, v, ms = wl.plummer(n)
    probe, _ = cpu_gravity_sample(u, ms, 32, 1, seed=99)
    ntargets = int(min(n, max(32, probe * 2.0 / (n - 1))))          # ~2 s of 1-thread work per step
    for w_ in range(args.warmup):
        cpu_gravity_sample(u, ms, max(16, ntargets // 8), 1, seed=w_)
    t_total, pairs = 0.0, 0.0
    for k in range(args.steps):
        _, dt = cpu_gravity_sample(u, ms, ntargets, 1, seed=1000 + k)
        t_total += dt
        pairs += ntargets * (n - 1)
    value = pairs / t_total
    threads = host_threads()
    rate_all, dt_all = cpu_gravity_sample(u, ms, 64 * threads, threads, seed=7)
    wl_lj, u_lj = lj_inputs()
    lj_rate, lj_dt = cpu_lj_sample(wl_lj, u_lj, 256, 1, seed=3)
    lj_all, _ = cpu_lj_sample(wl_lj, u_lj, 64 * threads, threads, seed=4)
    sample = f"{ntargets} targets x {n} sources per step (of {n} targets), 1 thread, scaled by time"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pair-interactions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (n * (n - 1) / value),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": gravity_config(),
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": 1, "kind": "port", "sample": sample},
        "all_threads": {"value": rate_all, "unit": "pair-interactions/s", "cores": threads,
                        "note": "OpenMP over targets -- NOT the reference (NBodySimulator.jl runs the loop on one thread)",
                        "sample": f"{64 * threads} targets x {n} sources ({dt_all:.1f} s)"},
        "lj": {"metric": LJ_METRIC, "value": lj_rate, "unit": "atom-steps/s", "cores": 1, "kind": "port",
               "sample": f"256 targets x {u_lj.shape[1]} sources ({lj_dt:.1f} s): the reference tests all N-1 partners of every atom "
                         "(no neighbour structure), force loop only",
               "all_threads": {"value": lj_all, "cores": threads}},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# second half of BASELINE.json's metric: LJ argon atom-steps/s (config 3)
# ------------------------------------------------------------------------------------------------
LJ_FLOP_PER_ATOM_STEP = 24.0 * 38.5      # algorithmic minimum, in-cutoff pairs only
LJ_BYTES_PER_ATOM_STEP = 200.0           # fused VV 144 B + cell rebuild 56 B


def _lj_context(local, w, n):
    from nbody_b200 import _lib

    ctx = _lib.Context(local)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
    return ctx


def lj_secondary(local, world, steps, warmup, cells=LJ_CELLS, weak=False):
    """1,048,576-atom FCC argon box, cubic PBC, R = 2.25 sigma, Berendsen, velocity Verlet on the device, through
    nbx_step_vv.  One GPU: the whole box in one context.  N GPUs: x-slabs inside the library (slab_enqueue: halo
    positions by peer-memory stores, rebuild decision from a peer-memory all-reduce, CUDA graph of two steps).
    Collective: every rank calls it; the returned dict is complete on rank 0 (times are max over ranks)."""
    import torch
    import torch.distributed as dist

    from nbody_b200 import _lib
    from nbody_b200.parallel import join_group_dist
    from oracle import nbody_oracle as orc

    rank = env_int("RANK", 0)
    w, u = lj_inputs(cells)
    n = u.shape[1]
    ctx = _lj_context(local, w, n)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    ctx.upload(u, w["v"])
    if world > 1:
        join_group_dist(ctx)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    steps = max(steps, 200)            # long enough to amortise list rebuilds / graph capture
    run = lambda k: ctx.step_vv(w["dt"], k)  # noqa: E731
    run(max(warmup, 3) + 40)
    barrier()
    reb0 = ctx.info("verlet_rebuilds")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n * steps / (ms * 1e-3)
    rebuilds = ctx.info("verlet_rebuilds") - reb0
    graph = ctx.info("graph_cached") if world > 1 else ctx.info("graph_if_nodes")
    _, _, T_after = ctx.energy(potential=False)
    # phase shares from a short run with the library's event timers on (eager launches)
    ctx.timing_reset()
    ctx.timing_enable(True)
    k_t = 20
    run(k_t)
    barrier()
    ctx.timing_enable(False)
    pair_ms, _ = ctx.timing_get(_lib.T_PAIR_CELLS)
    build_ms, _ = ctx.timing_get(_lib.T_CELL_BUILD)
    int_ms, _ = ctx.timing_get(_lib.T_INTEGRATE)
    own = ctx.info("slab_own") if world > 1 else n
    ghosts = ctx.info("slab_ghost") if world > 1 else 0
    # ---- parity of the path that was timed: resident accelerations after one thermostat-free step vs the CPU oracle ----
    ctx.thermostat(_lib.THERMO_NONE)
    run(1)
    if world == 1:
        uu, _, aa = ctx.download(want_dv=True)
        parts = [(np.arange(n), uu, aa)]
    else:
        gid, uu, _, aa = ctx.slab_download()
        parts = [None] * world
        dist.all_gather_object(parts, (gid, uu, aa))
    full, acc = np.zeros((3, n), order="F"), np.zeros((3, n), order="F")
    seen = np.zeros(n, dtype=np.int64)
    for g, x, a in parts:
        full[:, g] = x
        acc[:, g] = a
        seen[g] += 1
    ctx.close()
    parity = None
    if rank == 0:
        pick = np.random.Generator(np.random.Philox(11)).choice(n, PARITY_TARGETS, replace=False)
        ref = orc.System(w["ms"], bc=("cubic", w["L"]), lj=w["lj"]).accel_targets(full, pick, host_threads())
        err = np.linalg.norm(acc[:, pick] - ref, axis=0) / np.maximum(np.linalg.norm(ref, axis=0), 1e-300)
        parity = {"max_rel_err_per_body": float(err.max()), "targets": PARITY_TARGETS, "tolerance": 1e-12,
                  "ownership_is_a_partition": bool((seen == 1).all()),
                  "what": "after the timed run: one more step without the thermostat term, then the resident accelerations of a "
                          "subsample drawn over ALL ranks' atoms vs the CPU oracle's N-1 partner loop at the same positions"}
    out = {
        "metric": LJ_METRIC if not weak else f"LJ argon atom-steps/s ({n:,} atoms, weak scaling point)",
        "value": value, "unit": "atom-steps/s", "ms_per_step": ms / steps, "steps": steps, "n_atoms": n,
        "n_gpus": world, "scaling": "weak" if weak else "strong",
        "parallelism": "1 GPU" if world == 1 else f"x-slabs x{world} inside libnbody_b200 (nbx_step_vv on a group): halo positions and "
                       "the rebuild decision travel by peer-memory stores + flags from the step kernels; no NCCL call, no host "
                       "decision per step; CUDA graph of two steps with the rebuild chain in an IF node",
        "temperature_after": T_after,
        "inputs": "25 MB of positions per 1M atoms (L2-resident), Verlet lists streamed from HBM; L2 not flushed between steps",
        "ms_per_step_pair_kernel": pair_ms / k_t, "ms_per_step_cell_build": build_ms / k_t,
        "ms_per_step_integrate": int_ms / k_t,
        "neighbour_structure": f"Verlet lists (skin 0.08 R) over the cell list, rebuilt when a particle moved skin/2 (decided on the "
                               f"device{', collectively' if world > 1 else ''}): {rebuilds} rebuilds in the {steps} timed steps",
        "graph": bool(graph), "parity": parity,
    }
    if world > 1:
        out["rank0_own"], out["rank0_ghosts"] = own, ghosts
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    return out


# ------------------------------------------------------------------------------------------------
# the remaining BASELINE.json configs (4: SPC/Fw water x 32,768 molecules; 5: 65,536 charged / magnetic bodies,
# Langevin SDE variant), state resident, CUDA events on the library's stream -- reported beside the headline
# ------------------------------------------------------------------------------------------------
def other_configs(local, world, steps=5):
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import join_group_dist

    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    out = {}

    def timed(run, k):
        run(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(k)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def grouped(ctx, u, v):
        ctx.set_stream(side.cuda_stream)
        ctx.upload(u, v)
        if world > 1:
            join_group_dist(ctx)
        return ctx.info("group_mode") if world > 1 else 0

    modes = {0: "1 GPU", 1: "pair sharding", 2: "target blocks", 3: "x-slabs"}

    def parity_of(ctx, system, molecules=None, ntargets=96):
        """One GPU: the accelerations left resident by the timed run against the CPU oracle at the resident positions, for a
        subsample (whole molecules for water).  (Groups: the headline and the LJ run carry the multi-GPU parity figures.)"""
        if world > 1:
            return None
        try:
            from oracle import nbody_oracle as orc

            u_now, _, a_now = ctx.download(want_v=False, want_dv=True)
            rng = np.random.Generator(np.random.Philox(21))
            if molecules is not None:
                mols = np.sort(rng.choice(molecules, ntargets // 3, replace=False))
                cols = (3 * mols[:, None] + np.arange(3)[None, :]).ravel()
                ref = orc.System(**system).accel_molecules(u_now, mols, host_threads())
            else:
                cols = np.sort(rng.choice(u_now.shape[1], ntargets, replace=False))
                ref = orc.System(**system).accel_targets(u_now, cols, host_threads())
            norms = np.linalg.norm(ref, axis=0)
            floor = 1e-3 * float(np.sqrt(np.mean(norms ** 2)))   # the tests' metric: bodies whose net acceleration cancels below
            err = np.linalg.norm(a_now[:, cols] - ref, axis=0) / np.maximum(norms, max(floor, 1e-300))  # 1e-3 RMS are judged against it
            return {"max_rel_err_per_body": float(err.max()), "median_rel_err_per_body": float(np.median(err)),
                    "targets": int(len(cols)), "tolerance": 1e-12,
                    "what": "resident accelerations after the timed steps vs the CPU oracle (fp64 restatement) at the resident "
                            "positions; heavily cancelling sums are refereed in long double by tests/test_gpu_parity.py"}
        except Exception as e:  # a diagnostic must not cost the line
            return {"error": repr(e)}

    try:
        for tag, rel, what in (("config4_water_cutoff_0.9162nm", 0.9162, "cell lists for O-O Lennard-Jones and Coulomb"),
                               ("config4_water_cutoff_0.49L", None, "Coulomb cutoff 4.886 nm = 0.49 L (no cell list possible): every unordered pair "
                                                                    "once, the reference's periodic predicate (Newton's-third-law "
                                                                    "kernel; groups shard the ring offsets)")):
            w = wl.water_omm(32, Rel=rel)
            ctx = _lib.Context(local)
            ctx.system(w["ms"], qs=w["qs"], water=True)
            ctx.boundary(_lib.BC_CUBIC, [w["L"]])
            ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
            ctx.add_coulomb(w["coulomb"]["k"], w["coulomb"]["R"])
            ctx.add_spcfw(w["spcfw"]["rOH"], w["spcfw"]["aHOH"], w["spcfw"]["kb"], w["spcfw"]["ka"])
            mode = grouped(ctx, w["u"], w["v"])
            ms = timed(lambda k: ctx.step_vv(w["dt"], k), steps if rel is None else 10 * steps)
            out[tag] = {"metric": "SPC/Fw water molecule-steps/s (32,768 molecules, LJ + Coulomb cutoff + bonds/angles, velocity Verlet)",
                        "value": w["nmol"] / (ms * 1e-3), "unit": "molecule-steps/s", "ms_per_step": ms, "n_atoms": 3 * w["nmol"],
                        "coulomb_cutoff_nm": w["coulomb"]["R"], "path": what, "n_gpus": world, "decomposition": modes[mode],
                        "parity": parity_of(ctx, dict(ms=w["ms"], qs=w["qs"], water=True, bc=("cubic", w["L"]), lj=w["lj"],
                                                      coulomb=w["coulomb"], spcfw=w["spcfw"]), molecules=w["nmol"])}
            ctx.close()
        # config 2 at its largest size (the 8-GPU target of the north star): 1.1e12 pairs per evaluation
        n1 = 1048576
        u1, v1, m1 = wl.plummer(n1)
        ctx = _lib.Context(local)
        ctx.system(m1)
        ctx.add_gravity(1.0)
        mode = grouped(ctx, u1, v1)
        ms = timed(lambda k: ctx.step_vv(1e-4, k), 2)
        out["config2_gravity_1048576"] = {"metric": "gravity pair-interactions/s (all-pairs Plummer sphere, 1,048,576 bodies, fp64)",
                                          "value": float(n1) * float(n1 - 1) / (ms * 1e-3), "unit": "pair-interactions/s",
                                          "ms_per_step": ms, "n_bodies": n1, "n_gpus": world, "decomposition": modes[mode],
                                          "tflops_at_20_flop_per_pair": 20.0 * float(n1) * float(n1 - 1) / (ms * 1e-3) / 1e12}
        ctx.close()
        n = 65536
        for tag, gen in (("config5a_coulomb", wl.charged_lattice), ("config5b_dipole", wl.dipole_lattice)):
            w = gen(n)
            ctx = _lib.Context(local)
            ctx.system(w["ms"], qs=w.get("qs"), mm=w.get("mm"))
            if "coulomb" in w:
                ctx.add_coulomb(w["coulomb"]["k"], float("inf"))
            else:
                ctx.add_dipole(w["dipole"]["mu_4pi"])
            # Langevin SDE variant (src/nbody_to_ode.jl:567-598, Euler-Maruyama as test/thermostat_test.jl:85-89)
            ctx.thermostat(_lib.THERMO_LANGEVIN, 90.0, 10.0, 1.38e-23, n, 0)
            mode = grouped(ctx, w["u"], w["v"])
            dt = 1e-9
            ms = timed(lambda k: ctx.step_em(dt, k), steps)
            osys = dict(ms=w["ms"], qs=w.get("qs"), mm=w.get("mm"))
            if "coulomb" in w:
                osys["coulomb"] = dict(k=w["coulomb"]["k"], R=float("inf"))
            else:
                osys["dipole"] = w["dipole"]
            out[tag] = {"metric": "pair-interactions/s (65,536 bodies, all-pairs, Langevin thermostat, Euler-Maruyama steps)",
                        "value": float(n) * float(n - 1) / (ms * 1e-3), "unit": "pair-interactions/s", "ms_per_step": ms,
                        "n_bodies": n, "n_gpus": world, "decomposition": modes[mode],
                        "parity": parity_of(ctx, osys)}   # (a(x_end) is left resident by nbx_step_em)
            ctx.close()
        if world == 1:
            # config 1, the reference's own CPU-runnable case (examples/liquid_argon.jl as shipped: 216 atoms, R = L/2, velocity
            # Verlet, no thermostat): the only configuration whose WHOLE workload the CPU port runs here
            from oracle import nbody_oracle as orc

            w = wl.liquid_argon_si(216)
            ctx = _lib.Context(local)
            ctx.system(w["ms"])
            ctx.boundary(_lib.BC_CUBIC, [w["L"]])
            ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
            grouped(ctx, w["u"], w["v"])
            ms = timed(lambda k: ctx.step_vv(w["dt"], k), 3000)
            ctx.close()
            sysc = orc.System(w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
            t0 = time.perf_counter()
            orc.velocity_verlet(sysc, w["u"], w["v"], w["dt"], 300, 1)
            cpu_ms = (time.perf_counter() - t0) / 300 * 1e3
            out["config1_liquid_argon_216"] = {"metric": "LJ argon atom-steps/s (examples/liquid_argon.jl as shipped: 216 atoms, R = L/2, velocity Verlet)",
                                               "value": 216 / (ms * 1e-3), "unit": "atom-steps/s", "ms_per_step": ms, "n_atoms": 216,
                                               "n_gpus": 1, "decomposition": "1 GPU (a graph of two steps; the step is launch-bound)",
                                               "cpu_port_1_thread": {"value": 216 / (cpu_ms * 1e-3), "unit": "atom-steps/s",
                                                                     "ms_per_step": cpu_ms, "sample": "300 of the example's 30,000 steps"}}
    except Exception as e:  # never lose the headline line
        out["error"] = repr(e)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    return out


def lj_rooflines(out, hbm_peak, fp64_peak, src):
    value = out["value"]
    out["roofline_fp64"] = {"achieved": LJ_FLOP_PER_ATOM_STEP * value / 1e12, "peak": fp64_peak * out["n_gpus"],
                            "unit": "TFLOP/s",
                            "frac": LJ_FLOP_PER_ATOM_STEP * value / 1e12 / (fp64_peak * out["n_gpus"]) if fp64_peak else None}
    out["roofline_hbm"] = {"achieved": LJ_BYTES_PER_ATOM_STEP * value / 1e9, "peak": hbm_peak * out["n_gpus"], "unit": "GB/s",
                           "frac": LJ_BYTES_PER_ATOM_STEP * value / 1e9 / (hbm_peak * out["n_gpus"]) if hbm_peak else None,
                           "peak_source": src}
    return out


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def captured_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the newest committed
    `ncu --set full` summary under profiles/ (None when there is none: the figure is never hard-coded)."""
    import csv
    import glob

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "prof_gravity_r*.summary.csv"))):
        rd = wr = None
        try:
            with open(path) as f:
                for row in csv.DictReader(f):
                    if not row.get("kernel", "").startswith(kernel_prefix) or row.get("launch") != "0":
                        continue
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row.get("unit"), None)
                    if scale is None:
                        continue
                    if row["metric"] == "dram__bytes_read.sum":
                        rd = float(row["value"]) * scale
                    if row["metric"] == "dram__bytes_write.sum":
                        wr = float(row["value"]) * scale
        except (OSError, ValueError, KeyError):
            continue
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.relpath(path, ROOT))
    return best


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import join_group_dist
    from oracle import nbody_oracle as orc

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = N_GRAVITY
    u, v, ms = wl.plummer(n)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)

    def gravity_context(masses, mode):
        ctx = _lib.Context(local)
        ctx.system(masses)
        ctx.add_gravity(1.0)
        ctx.set_stream(side.cuda_stream)
        ctx.set_option("pin_host", 1)   # the caller's u / dv buffers are page-locked once, at the first nbx_accel that sees them
        ctx.upload(u, v)
        if world > 1:
            join_group_dist(ctx, mode=mode)
        return ctx

    gmode = {"pairs": 1, "targets": 2}[args.mode]
    ctx = gravity_context(ms, gmode)
    lo, hi = (ctx.info("shard_lo"), ctx.info("shard_hi")) if world > 1 else (0, n)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(c, k):
        """k steps, one nbx_step_vv call each, L2 flushed (untimed) in between; summed device time, max over ranks."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        for j in range(k):
            flush.zero_()                       # untimed: evict the step's working set from L2
            ev[j][0].record()
            c.step_vv(DT_GRAVITY, 1)
            ev[j][1].record()
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # (nvidia-smi needs a few hundred ms to deliver its first sample: started before the warm-up)
    ctx.step_vv(DT_GRAVITY, max(args.warmup, 3))
    barrier()
    ctx.timing_reset()
    ctx.timing_enable(True)
    t_begin = time.perf_counter()
    ms_total = timed_steps(ctx, args.steps)
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    ctx.timing_enable(False)
    pairs_per_step = float(n) * float(n - 1)
    value = pairs_per_step * args.steps / (ms_total * 1e-3)
    k_ms, k_cnt = ctx.timing_get(_lib.T_PAIR_ALLPAIRS)
    i_ms, i_cnt = ctx.timing_get(_lib.T_INTEGRATE)

    # ---- parity of what the timed run left on the device: accelerations of a subsample of rank 0's targets vs the oracle
    parity = None
    u_now, _, a_now = ctx.download(want_v=False, want_dv=True)
    if world > 1:
        blocks = [None] * world
        dist.all_gather_object(blocks, (lo, hi, np.ascontiguousarray(a_now[:, lo:hi])))
        for blo, bhi, blk in blocks:
            a_now[:, blo:bhi] = blk
    if rank == 0:
        pick = np.random.Generator(np.random.Philox(12)).choice(n, PARITY_TARGETS, replace=False)
        ref = orc.System(ms, gravity=dict(G=1.0)).accel_targets(u_now, pick, host_threads())
        err = np.linalg.norm(a_now[:, pick] - ref, axis=0) / np.maximum(np.linalg.norm(ref, axis=0), 1e-300)
        parity = {"max_rel_err_per_body": float(err.max()), "median_rel_err_per_body": float(np.median(err)),
                  "targets": PARITY_TARGETS, "tolerance": 1e-12,
                  "what": "accelerations resident after the timed steps (a subsample drawn over ALL ranks' blocks) vs the CPU oracle "
                          "at the same positions"}

    # ---- end to end through the RHS drop-in with host buffers -----------------------------------
    dv = np.zeros((3, n), order="F")
    rhs = lambda: ctx.accel(u, out=dv)  # noqa: E731
    for _ in range(2):
        rhs()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rhs()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_per_step * args.steps / float(t.item())
    h2d = int(3 * (hi - lo) * 8)
    d2h = int(3 * (hi - lo) * 8)
    peak_tf, eff_mhz = ctx.measure_fp64_peak() if rank == 0 else (0.0, 0.0)
    grid, chunks = ctx.info("allpairs_grid"), ctx.info("allpairs_chunks")
    ctx.close()

    # ---- general masses (the headline rides the equal-mass variant: 18 instead of 20 FP64 instructions per unordered pair)
    general = None
    if not args.no_lj:
        msg = ms * (0.5 + np.random.Generator(np.random.Philox(4)).random(n))
        cg = gravity_context(msg, gmode)
        cg.step_vv(DT_GRAVITY, 3)
        g_ms = timed_steps(cg, max(3, args.steps // 2))
        general = {"value": pairs_per_step * max(3, args.steps // 2) / (g_ms * 1e-3), "unit": "pair-interactions/s",
                   "ms_per_step": g_ms / max(3, args.steps // 2), "masses": "unequal (0.5 .. 1.5) / N"}
        cg.close()

    # ---- second half of the metric: LJ argon atom-steps/s (all ranks take part) -------------------
    lj = None
    if not args.no_lj:
        try:
            lj = lj_secondary(local, world, max(args.steps, 5), args.warmup)
        except Exception as e:  # the headline line must still be printed
            lj = {"error": repr(e)}
    others = None
    if not args.no_lj:
        others = other_configs(local, world)
        if world > 1:   # weak-scaling point of the LJ path: 1,048,576 atoms per GPU where the lattice allows (cells^3 x 4)
            cells = {2: 80, 4: 101, 8: 128}.get(world)
            if cells:
                try:
                    wk = lj_secondary(local, world, 100, 3, cells=cells, weak=True)
                except Exception as e:
                    wk = {"error": repr(e)}
                others["lj_weak_scaling"] = wk

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    pairs_mode = world == 1 or args.mode == "pairs"
    kernel_ms = k_ms / max(k_cnt, 1)
    instr_pp = 9 if pairs_mode else 16  # FP64-pipe instructions per ORDERED pair
    share = 1.0 / world
    if pairs_mode and world > 1:   # rank 0's ring offsets k = 0, world, 2 world, ... of the NT-tile half ring; the half offset k = K of an
        NT = -(-n // 1024)         # even ring is dealt out over all ranks by tile (csrc/nbx_sympairs.cu)
        K = NT // 2
        if NT % 2 == 0 and K > 0:
            share = (len(range(0, K, world)) + 0.5 / world) / (K + 0.5)
        else:
            share = len(range(0, K + 1, world)) / (K + 1.0)
    pairs_per_launch = pairs_per_step * share
    achieved_tf = FLOP_PER_PAIR * pairs_per_launch / (kernel_ms * 1e-3) / 1e12
    traffic = captured_traffic("void sym_kernel") if world == 1 else None
    roofline = {
        "bound": "fp64", "kernel": "sym_kernel<8,2,uniform,2,4> (Newton 3rd law, 18 FP64 instr per unordered pair = 9 per ordered pair)" if pairs_mode else "allpairs_kernel<GravPolicy> (16 FP64 instr per ordered pair)", "achieved": achieved_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf > 0 else None,
        "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
        "peak_source": "DFMA saturation microbenchmark measured live on this device (MEASURED_PEAKS.json has no "
                       "FP64 figure); nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2",
        "peak_effective_sm_mhz": eff_mhz, "flop_per_pair": FLOP_PER_PAIR, "frac_of_nominal_37.2": achieved_tf / 37.2,
        "fp64_instr_per_pair": instr_pp, "pipe_bound_frac_of_peak": FLOP_PER_PAIR / (instr_pp * 2.0),
        "kernel_ms": kernel_ms, "kernel_launches": k_cnt, "kernel_share_of_step": k_ms / ms_total, "rank0_share_of_pairs": share,
        "integrate_and_exchange_ms_per_step": i_ms / max(args.steps, 1),
    }

    # ---- CPU baseline on the host cores (bounded sample) ---------------------------------------------
    rate1, dt1 = cpu_gravity_sample(u, ms, 64, 1, seed=5)
    nt = int(min(n, max(64, rate1 * 12.0 / (n - 1))))
    rate, dt_cpu = cpu_gravity_sample(u, ms, nt, 1, seed=6)
    cpu = {"value": rate, "unit": "pair-interactions/s", "cores": 1, "kind": "port",
           "sample": f"{nt} targets x {n} sources ({dt_cpu:.1f} s, 1 thread: the reference is single-threaded)",
           "host_threads_available": host_threads()}

    # kernels of one timed step (the launch lists under profiles/ show the same sequence)
    per_step = (["vv_pos_kernel", "sym_kernel", "sym_reduce_kernel", "vv_vel_kernel", "final_sum_kernel"] if world == 1 else
                ["vv_pos_push_kernel", "comm_wait_kernel", "sym_kernel", "sym_reduce_kernel", "acc_push_kernel", "acc_sum_kernel",
                 "vv_vel_kernel", "final_sum_kernel"] if pairs_mode else
                ["vv_pos_push_kernel", "comm_wait_kernel", "allpairs_kernel", "allpairs_reduce_kernel", "vv_vel_kernel", "final_sum_kernel"])
    out = {
        "metric": METRIC, "value": value, "unit": "pair-interactions/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": gravity_config(),
        "details": {"integrator": "velocity_verlet", "dt": DT_GRAVITY, "l2": "flushed between timed steps (256 MB write, untimed)",
                    "parallelism": ("1 GPU" if world == 1 else
                                    (f"pair sharding x{world} inside libnbody_b200 (nbx_step_vv on a group): positions all-gathered by the "
                                     "update kernel's peer stores, partial accelerations pushed to their owners and added in rank order; "
                                     "no NCCL call per step" if pairs_mode else
                                     f"target-block sharding x{world}, positions all-gathered by the update kernel's peer stores")),
                    "allpairs_grid": grid, "allpairs_chunks": chunks},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pair-interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "nbx_accel (RHS drop-in, host pointers)" + ("" if world == 1 else
                       " on a group member: own block of u up, all-gather over NVLink, own columns of dv back"),
                "host_buffers": "the caller's arrays, page-locked once by the library (option pin_host)",
                "bytes": "per rank"},
        "gpu_launches": len(per_step) * args.steps, "kernels_per_step": per_step,
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "general_masses": general,
    }
    if lj is not None:
        if "error" not in lj:
            hbm, src = measured_hbm_peak()
            lj_rooflines(lj, hbm, peak_tf, src)
        out["lj"] = lj
    if others is not None:
        out["other_configs"] = others
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-lj", action="store_true", help="headline only: skip the LJ argon half, the general-mass run and the other configs")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "targets"], help="multi-GPU decomposition of the gravity workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
