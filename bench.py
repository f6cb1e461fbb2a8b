#!/usr/bin/env python
"""Benchmark of the 2-D VOF per-timestep hot path (BASELINE.json: timesteps/s and Jacobi
Gcell-updates/s at 8192^2; HBM GB/s as % of peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192] [--preroll 1000]
    python bench.py --nx-global 32768 --gpus 8            # BASELINE config 4: fixed 32768^2, strong scaling
    python bench.py --dim 3 --n 512 --gpus 8              # BASELINE config 5: 3-D dam break, 512^3 over plane slabs

One "step" = one pass of the loop body 2dvof.py:513-528 (props, normals/curvature, advection,
BC, 10 Jacobi sweeps, projection, BC, FCT x/y, post-process, BC) over the whole grid.
Workload (configs[2] of BASELINE.json, the configuration the metric is quoted on): dropping
liquid (-ic 3) at 8192^2 per GPU, fp32, constant-dx scaling (L = 0.1 * n / 200 so that
nu*dt/dx^2 stays at the reference's stable value; SURVEY.md 7 risk 3), synthetic = generated
by the solver's own set_init_F.  For N > 1 the domain is (8192 N) x 8192, row-slab decomposed
(weak scaling), one halo exchange per step.

The kernels adapt to the data (exact short-cuts on bulk gas / liquid rows), so the state matters: the first steps
after set_init_F are the cheapest input there is (u == 0 almost everywhere, F exactly 0 or 1).  The timed region
therefore runs on a DEVELOPED state: `--preroll` (default 1000) un-timed steps first, after which the pressure
front has crossed the whole grid and u, v are non-zero everywhere; the early-state time is reported beside it
(`early_state`), and `kernels_general_path` times the data-adaptive kernels on a random F field (no bulk at all).

Prints ONE JSON line (rank 0).  `value` = Jacobi cell-updates/s of the whole job with state
resident in HBM (= n_jacobi * cells * steps / time, the definition in BASELINE.md 3.4);
`timesteps_per_s` rides along.  `e2e` = the same through vof2d_streamer_step_host with pinned HOST
buffers (H2D of u,v,p,F + step + D2H of u,v,p,F inside the timed region).  `parity` compares the GPU fields
with the CPU leg's oracle state at the same step (N = 1) or a slab run with a single-GPU run (N > 1).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_JACOBI = 10
IC_NAMES = {1: "dam break", 2: "rising bubble", 3: "dropping liquid"}
METRIC = "Jacobi Gcell-updates/s over whole timesteps (10 sweeps/step), 8192^2 cells per GPU"
METRIC3 = "Jacobi Gcell-updates/s over whole timesteps (10 sweeps/step), 3-D dam break (3dvof.py path)"
UNIT = "Gcell-updates/s"
STEP_BYTES_3D = 236      # advect 28 + rhs 20 + Jacobi 12 x 10 + project 32 + FCT 12 x 3
# algorithmic bytes per cell per launch (SURVEY.md 8d / DESIGN.md): fp32 arrays read + written once
ALGO_BYTES = {"kappa": 8, "advect": 24, "rhs": 16, "jacobi": 12, "project": 24, "fct_x": 12, "fct_y": 12, "props": 12}
STEP_BYTES_UNBLOCKED = 8 + 24 + 16 + 12 * N_JACOBI + 24 + 24          # 216 B/cell: one HBM pass per sweep


def workload_config(n, ic, dim=2):
    """The `config` both arms print (same keys, same values, so the driver can tell they ran the same thing)."""
    if dim == 3:
        return {"workload": f"3-D dam break (3dvof.py, -ic 1) at {n}^3, constant-dx scaling, {N_JACOBI} Jacobi sweeps/step",
                "grid": [n, n, n], "ic": 1, "n_jacobi": N_JACOBI, "dt": 4e-6, "constants": "constant dx: L = 0.1 * n / 200"}
    return {"workload": f"{IC_NAMES[ic]} (-ic {ic}) at {n}^2 per GPU, constant-dx scaling, {N_JACOBI} Jacobi sweeps/step",
            "grid": [n, n], "ic": ic, "n_jacobi": N_JACOBI, "dt": 4e-6, "constants": "constant dx: L = 0.1 * n / 200"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm_sorted = sorted(sm)
            out.update(sm_mhz=sm_sorted[len(sm_sorted) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement of the reference (oracle "port"), all host threads
# --------------------------------------------------------------------------------------
def time_cpu_port(n, steps, warmup, ic=3, dim=2, keep=False):
    """Times the C/OpenMP oracle on every core this process may use.  keep=True also returns the oracle (its state
    after warmup + steps steps is what `parity` compares the GPU fields with)."""
    from oracle import c_oracle
    threads = c_oracle.use_all_host_cores()     # torchrun hands its workers OMP_NUM_THREADS=1
    if dim == 3:
        from oracle.vof3d_oracle import Vof3DParams
        o = c_oracle.Vof3DCOracle(Vof3DParams.scaled(n))
        ic, cells = 1, n ** 3
    else:
        from oracle.vof2d_oracle import Vof2DParams
        o = c_oracle.Vof2DCOracle(Vof2DParams.scaled(n))
        cells = n * n
    o.set_init_F(ic)
    o.run(warmup)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    r = {"seconds": dt, "steps_per_s": steps / dt, "gcell_updates_per_s": N_JACOBI * cells * steps / dt / 1e9,
         "threads": threads, "n": n, "steps": steps, "warmup": warmup}
    return (r, o) if keep else r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the full grid when the whole run stays within a few minutes (~0.5 s/step at 8192^2 on 16 cores),
    # else a smaller grid of the same workload (throughput per cell is size-independent far beyond the caches)
    n = args.n
    if args.dim == 3:
        n = args.n if (args.steps + args.warmup) <= 12 else min(args.n, 256)
    elif (args.steps + args.warmup) > 40:
        n = min(args.n, 4096)
    r = time_cpu_port(n, args.steps, args.warmup, ic=args.ic, dim=args.dim)
    what = "3dvof.py" if args.dim == 3 else "2dvof.py"
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) from set_init_F at {n}^{args.dim}, C/OpenMP restatement of {what} with the "
              f"reference's loop structure (taichi 1.4.1 not installable: py3.12, offline), {r['threads']} threads; its cost does not "
              f"depend on the state (no data-adaptive paths), so the early steps are a fair sample of the developed flow")
    line = {
        "impl": "reference", "metric": METRIC3 if args.dim == 3 else METRIC, "value": r["gcell_updates_per_s"], "unit": UNIT,
        "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "timesteps_per_s": r["steps_per_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.n, args.ic, args.dim),
        "run": {"sample_grid": [n] * args.dim, "threads": r["threads"]},
        "cpu_baseline": {"value": r["gcell_updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": r["gcell_updates_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this rank's threads (and, by first touch, its pinned host buffers) on the CPUs NVML lists
    as local to its GPU, so that the e2e copies of 8 ranks do not all cross one socket's memory controllers.
    Returns the number of CPUs bound to, or None when NVML gives no affinity."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch
    from taichi_2d_vof_b200 import VofSolver2D, VofSolver3D, reference_params, reference_params3d
    from taichi_2d_vof_b200.slab import SlabSolver2D, slab_parity_check

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None      # pinned host buffers on the GPU's own NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    three_d = args.dim == 3
    n = args.n
    if three_d:                                # BASELINE config 5: a fixed n^3 grid over plane slabs (strong scaling)
        nx_global, ny, nz, scaling = n, n, n, ("strong" if world > 1 else "weak")
        L = 0.1 * n / 200.0
        cells_per_row = ny * nz

        def params_fn(slab, halo, device):
            return reference_params3d(nx=n, ny=n, nz=n, Lx=L, Ly=L, Lz=L, n_jacobi=N_JACOBI, slab=slab, halo=halo, device=device)
        solver_cls, fields, ic = VofSolver3D, ("F", "u", "v", "w", "p"), 1
    else:
        ny, nx_global, scaling = n, n * world, "weak"     # weak scaling: n rows per GPU
        if args.nx_global:                     # BASELINE config 4: fixed global grid (32768^2 over 2/4/8 GPUs)
            nx_global, ny, scaling = args.nx_global, (args.ny or args.nx_global), "strong"
        L_y, L_x = 0.1 * ny / 200.0, 0.1 * nx_global / 200.0
        cells_per_row = ny

        def params_fn(slab, halo, device):
            return reference_params(nx=nx_global, ny=ny, Lx=L_x, Ly=L_y, n_jacobi=N_JACOBI, slab=slab, halo=halo, device=device)
        solver_cls, fields, ic = VofSolver2D, ("F", "u", "v", "p"), args.ic

    slab = SlabSolver2D(params_fn, nx_global, rank, world, dist=dist, n_jacobi=N_JACOBI, device=local, transport=args.transport,
                        solver_cls=solver_cls, halo_fields=fields)
    slab.check_every = 0                       # the health check synchronises: once, after the timed region
    s = slab.solver
    slab.set_init_F(ic)
    stream = slab.stream
    can_profile = hasattr(s, "profile")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(k, profile_every=0):
        """k steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        if can_profile:
            s.profile(profile_every)
        l0 = s.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(k):
            slab.step()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = s.launch_count() - l0
        prof = s.profile_read() if (can_profile and profile_every) else {}
        if can_profile:
            s.profile(False)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
            dist.all_reduce(lt)
            launches = int(lt.item())
        return ms, launches, prof

    cells_per_gpu = (slab.hi - slab.lo + 1) * cells_per_row
    cells_total = nx_global * cells_per_row

    def kernel_table(prof, ms_step, sampled_steps):
        kern = {}
        for name, (tot_ms, nspan) in prof.items():
            if name not in ALGO_BYTES:
                continue
            per = tot_ms / nspan
            sweeps = (N_JACOBI * sampled_steps / nspan) if name == "jacobi" else 1.0
            bytes_cell = ALGO_BYTES[name]
            algo = bytes_cell * cells_per_gpu * sweeps
            kern[name] = {"launches_per_step": nspan / sampled_steps, "ms_per_launch": per, "algo_bytes_per_cell": bytes_cell,
                          "algo_GBps": algo / (per * 1e-3) / 1e9, "share_of_step": (tot_ms / sampled_steps) / ms_step}
            if name == "jacobi":
                kern[name]["sweeps_per_launch"] = sweeps
                kern[name]["gcell_updates_per_s"] = cells_per_gpu * sweeps / (per * 1e-3) / 1e9
        return kern

    # ---- early state: the first steps after set_init_F (what round 1 timed)
    for _ in range(args.warmup):
        slab.step()
    k_early = min(args.steps, 12)
    prof_every = 4 if k_early >= 8 else 1
    ms_e, _, prof_e = timed_steps(k_early, prof_every)
    early = {"steps": k_early, "after_steps": args.warmup, "ms_per_step": ms_e / k_early,
             "kernels_ms": {k: v[0] / v[1] for k, v in prof_e.items() if k in ALGO_BYTES}}

    # ---- pre-roll to a developed flow (un-timed), then the timed region
    t_pre = time.perf_counter()
    slab.run(args.preroll)
    barrier()
    t_pre = time.perf_counter() - t_pre
    for _ in range(args.warmup):
        slab.step()
    prof_every = 4 if args.steps >= 8 else 1
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches, prof = timed_steps(args.steps, prof_every)
    clk = clocks.stop() if rank == 0 else None
    slab.check()                               # a timed-out or out-of-step halo exchange voids the run

    d = slab.diagnostics(residual=False)       # all-reduced over the ranks
    finite = bool(np.isfinite(d["mass"]) and np.isfinite(d["max_cfl"]))
    sec = ms / 1e3
    value = N_JACOBI * cells_total * args.steps / sec / 1e9
    peak, peak_src = measured_peak()
    ms_step = ms / args.steps
    steps_sampled = len(range(0, args.steps, prof_every))
    kern = kernel_table(prof, ms_step, steps_sampled)
    tj = load_traffic()

    # ---- roofline of the dominant kernel (Jacobi).  One launch of the temporally blocked kernel does T sweeps per HBM
    # pass, so three fractions are reported: `frac` = the un-blocked algorithmic convention (12 B x cell-updates per
    # launch / time; legitimately > 1), `per_pass_frac` = what one pass must move (12 B x cells) / time, and `dram_frac`
    # = the bytes DRAM really moved per launch (ncu) / time.  `step` = the same for the whole step.
    roof = None
    if "jacobi" in kern:
        a = kern["jacobi"]["algo_GBps"]
        T = kern["jacobi"]["sweeps_per_launch"]
        per = kern["jacobi"]["ms_per_launch"]
        roof = {"kernel": "k_jacobi_pk<%d> (%g sweeps per HBM pass; packed fp32x2 register pipeline, third generation)" % (round(T), T), "bound": "hbm",
                "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "peak_source": peak_src,
                "algo_bytes_per_cell_update": 12, "cell_updates_per_launch": cells_per_gpu * T, "ms_per_launch": per,
                "frac_of_nominal_8TBps": a / 8000.0,
                "per_pass_GBps": 12 * cells_per_gpu / (per * 1e-3) / 1e9, "per_pass_frac": 12 * cells_per_gpu / (per * 1e-3) / 1e9 / peak,
                "traffic": None, "traffic_source": tj.get("source"),
                "note": "frac uses the un-blocked convention (12 B per cell-update): temporal blocking moves ~1/T of those bytes, "
                        "so it may exceed 1; per_pass_frac and dram_frac are the physical fractions"}
        tb = tj.get("jacobi_bytes_per_launch", tj.get("jacobi_tb_bytes_per_launch"))
        if tb and not three_d and cells_per_gpu == 8192 * 8192:
            roof["traffic"] = tb
            roof["dram_GBps"] = tb / (per * 1e-3) / 1e9
            roof["dram_frac"] = roof["dram_GBps"] / peak
        npass = kern["jacobi"]["launches_per_step"]
        blocked = (STEP_BYTES_UNBLOCKED - 12 * N_JACOBI) + 12 * npass - (12 if "rhs" not in kern else 0)
        roof["step"] = {"algo_bytes_per_cell_unblocked": STEP_BYTES_UNBLOCKED, "algo_bytes_per_cell_blocked": blocked,
                        "GBps_unblocked": STEP_BYTES_UNBLOCKED * cells_per_gpu / (ms_step * 1e-3) / 1e9,
                        "GBps_blocked": blocked * cells_per_gpu / (ms_step * 1e-3) / 1e9}
        roof["step"]["frac_unblocked"] = roof["step"]["GBps_unblocked"] / peak
        roof["step"]["frac_blocked"] = roof["step"]["GBps_blocked"] / peak

    if three_d:     # no per-kernel spans in the 3-D context: the whole step against its algorithmic bytes (one pass per sweep)
        gbps = STEP_BYTES_3D * cells_per_gpu / (ms_step * 1e-3) / 1e9
        roof = {"kernel": "whole 3-D step (k3_jacobi5 x 10 is ~half of it; see profiles/)", "bound": "hbm", "achieved": gbps, "peak": peak,
                "unit": "GB/s", "frac": gbps / peak, "peak_source": peak_src, "algo_bytes_per_cell_step": STEP_BYTES_3D,
                "frac_of_nominal_8TBps": gbps / 8000.0, "traffic": None}

    # ---- opt-in tolerance mode of the pressure sweeps (VOF_OPT_FAST_MATH: FMA contraction + reciprocal multiply, within the
    # north-star tolerances but NOT bit-exact -- tests/test_round2_gpu.py): a separate key, never the headline
    tolerance = None
    if rank == 0 and world == 1 and not three_d and can_profile:
        from taichi_2d_vof_b200 import _lib as _vl
        keep = {k: getattr(s, k).to_numpy() for k in ("u", "v", "p", "F")}
        s.set_option(_vl.VOF_OPT_FAST_MATH, 1)
        for _ in range(3):
            slab.step()
        k_tol = max(4, min(args.steps, 12))
        ms_t, _, prof_t = timed_steps(k_tol, 1)
        s.set_option(_vl.VOF_OPT_FAST_MATH, 0)
        for k, a_ in keep.items():                      # the exact state goes on (e2e, diagnostics)
            getattr(s, k).from_numpy(a_)
        jt = prof_t.get("jacobi")
        tolerance = {"option": "VOF_OPT_FAST_MATH = 1", "ms_per_step": ms_t / k_tol, "steps": k_tol,
                     "value": N_JACOBI * cells_total * k_tol / (ms_t * 1e-3) / 1e9, "unit": UNIT,
                     "jacobi_ms_per_launch": (jt[0] / jt[1]) if jt else None,
                     "jacobi_per_pass_GBps": (12 * cells_per_gpu / (jt[0] / jt[1] * 1e-3) / 1e9) if jt else None,
                     "note": "5 instead of 8 fp32 operations per cell-update in the blocked Jacobi; rel. L-inf vs the exact path <= 1e-5 after "
                             "one step, <= 1e-3 after 100 (tests/test_round2_gpu.py); not used for `value`"}

    # ---- e2e: the same step through the host-buffer C-ABI call (pinned memory, developed state)
    e2e = None
    if not args.no_e2e and not three_d:
        shape = (s.nrows, ny + 2)
        host = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(4)]
        arrs = [h.numpy() for h in host]
        for a_, k in zip(arrs, ("u", "v", "p", "F")):
            a_[...] = getattr(s, k).to_numpy()
        k_e2e = max(3, min(args.steps, 5))

        def timed(fn):
            fn()   # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(k_e2e):
                fn()
            barrier()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        def slab_roundtrip():             # N > 1: same transfers, with the halo exchange between upload and step
            for a_, k in zip(arrs, ("u", "v", "p", "F")):
                getattr(s, k).from_numpy(a_)
            slab.step()
            for a_, k in zip(arrs, ("u", "v", "p", "F")):
                getattr(s, k).to_numpy(out=a_)

        nbytes = 4 * shape[0] * shape[1] * 4
        if world == 1:
            # whole-array call first (explains the streamed number), then the streamed call = the e2e headline
            dt_plain = timed(lambda: s.step_host(*arrs))
            from taichi_2d_vof_b200 import VofStreamer2D
            n_slabs = max(1, min(args.e2e_slabs, nx_global // 32))
            st = VofStreamer2D(params_fn(None, 0, local), n_slabs=n_slabs)
            # separate pinned output arrays, swapped with the inputs after every step (the call's `out=` form): a slab's
            # download then need not wait for the next slab's upload of the rows it overwrites
            host2 = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(4)]
            pp = [arrs, [h.numpy() for h in host2]]

            def streamed():
                st.step_host(*pp[0], out=pp[1])
                pp[0], pp[1] = pp[1], pp[0]

            dt = timed(streamed)
            up_rows = sum(min(nx_global + 1, nx_global * (k + 1) // n_slabs + st.halo) - max(0, 1 + nx_global * k // n_slabs - st.halo) + 1
                          for k in range(n_slabs)) if n_slabs > 1 else shape[0]
            e2e = {"value": N_JACOBI * cells_total * k_e2e / dt / 1e9, "unit": UNIT, "timesteps_per_s": k_e2e / dt,
                   "h2d_bytes_per_step": 4 * up_rows * shape[1] * 4, "d2h_bytes_per_step": nbytes, "steps": k_e2e,
                   "api": f"vof2d_streamer_step_host (pinned host u,v,p,F in, separate pinned arrays out; {n_slabs} row slabs, halo {st.halo}, "
                          "upload / step / download overlapped on three streams, dense staging blocks on the device)",
                   "unstreamed": {"value": N_JACOBI * cells_total * k_e2e / dt_plain / 1e9, "timesteps_per_s": k_e2e / dt_plain,
                                  "api": "vof2d_step_host (whole arrays up, step, whole arrays down)"}}
            st.close()
        else:
            dt = timed(slab_roundtrip)
            e2e = {"value": N_JACOBI * cells_total * k_e2e / dt / 1e9, "unit": UNIT, "timesteps_per_s": k_e2e / dt,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": k_e2e,
                   "api": "per-rank slab: from_numpy(u,v,p,F), halo exchange + step, to_numpy(u,v,p,F) (pinned host)"}

    # ---- the data-adaptive kernels on a field with no bulk at all (random F, CFL ~ 0.1 velocities): general path everywhere
    general = None
    if rank == 0 and world == 1 and not three_d and not args.no_general:
        rng = np.random.default_rng(0)
        shape = (s.nrows, ny + 2)
        vel = 0.1 * s.P.dx / s.P.dt
        s.F.from_numpy(rng.random(shape, dtype=np.float32))
        s.u.from_numpy((rng.random(shape, dtype=np.float32) - 0.5) * 2 * vel)
        s.v.from_numpy((rng.random(shape, dtype=np.float32) - 0.5) * 2 * vel)
        general = {"state": "F ~ U(0,1) in every cell, u, v ~ U(-1,1) * 0.1 dx/dt"}
        for name, fn in (("kappa", s.get_normal_young), ("fct_x", s.fct_x_sweep), ("fct_y", s.fct_y_sweep)):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t_ms = sorted(ts)[1]
            general[name] = {"ms_per_launch": t_ms, "algo_GBps": ALGO_BYTES[name] * cells_per_gpu / (t_ms * 1e-3) / 1e9,
                             "frac": ALGO_BYTES[name] * cells_per_gpu / (t_ms * 1e-3) / 1e9 / peak}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only), bounded sample; its end state pins `parity`
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = min(n, 256) if three_d else min(n, 8192)
        k_cpu = 5 if three_d else 24
        r, orc = time_cpu_port(n_cpu, k_cpu, 1, ic=ic, dim=args.dim, keep=True)
        cpu = {"value": r["gcell_updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "timesteps_per_s": r["steps_per_s"],
               "sample": f"{k_cpu} steps (+1 warm-up) of the same workload from set_init_F at {r['n']}^{args.dim}, C/OpenMP restatement of the "
                         f"reference's loop structure, {r['threads']} threads ({r['seconds']:.1f} s)"}
        # the GPU path over the same k_cpu + 1 steps from the same initial condition, every element of every field
        if three_d:
            from taichi_2d_vof_b200 import scaled_params3d
            g = VofSolver3D(scaled_params3d(n_cpu, device=local))
        else:
            from taichi_2d_vof_b200 import scaled_params
            g = VofSolver2D(scaled_params(n_cpu, device=local))
        g.set_init_F(ic)
        g.run(k_cpu + 1)
        diffs, worst = {}, 0.0
        for k in fields:
            a_, b_ = getattr(g, k).to_numpy(), getattr(orc, k)
            diffs[k] = int(np.count_nonzero(a_ != b_))
            if diffs[k]:
                worst = max(worst, float(np.max(np.abs(a_.astype(np.float64) - b_))))
        parity = {"against": f"the C oracle of the cpu_baseline leg after {k_cpu + 1} steps at {n_cpu}^{args.dim} (oracle == the reference's "
                             "own source executed under the taichi stand-in, tests/test_reference_pin_cpu.py)",
                  "fields": list(fields), "cells_differing": diffs, "max_abs_diff": worst, "identical": all(v == 0 for v in diffs.values())}
        g.close()
        del orc
    elif world > 1 and not args.no_parity:
        parity = slab_parity_check(dist, rank, world, local, transport=args.transport, three_d=three_d, n=96)

    if rank == 0:
        cfg = workload_config(n, ic, args.dim)
        line = {
            "metric": METRIC3 if three_d else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "timesteps_per_s": args.steps / sec, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"state": f"developed flow: {args.preroll} un-timed steps (+{2 * args.warmup + k_early} warm-up / early-state steps) after set_init_F "
                             f"({t_pre:.1f} s of GPU time); fused step (vof{args.dim}d_step) + one halo exchange per step",
                    "grid_per_gpu": [slab.hi - slab.lo + 1, ny] + ([nz] if three_d else []),
                    "global_grid": [nx_global, ny] + ([nz] if three_d else []),
                    "decomposition": ("row slabs along i, deep halo %d rows, 1 exchange/step (one fused peer-store kernel), transport %s"
                                      % (s.halo, slab.transport)) if world > 1 else "single GPU",
                    "l2": "inputs exceed L2 (live fp32 fields of %.0f MB each >> 126 MB); no flush needed" % (s.nrows * (ny + 2) * (nz + 2 if three_d else 1) * 4 / 1e6),
                    "rows_processed_per_launch": s.nrows, "cpus_bound_per_rank": numa},
            "early_state": early,
            "roofline": roof, "kernels": kern, "kernels_general_path": general,
            "kernel_event_sampling": f"CUDA events around every kernel of every {prof_every}th step of the timed region" if prof_every > 1 else "CUDA events around every kernel of the timed region",
            "cpu_baseline": cpu, "e2e": e2e, "tolerance_mode": tolerance, "parity": parity, "gpu_launches": launches,
            "clocks": clk, "state_finite": finite, "mass": d["mass"], "mass_per_rank": d.get("mass_per_rank"), "max_cfl": d["max_cfl"],
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--dim", type=int, choices=[2, 3], default=2, help="3: the 3dvof.py path (BASELINE config 5), a fixed n^3 grid over the GPUs")
    ap.add_argument("--n", "--size", dest="n", type=int, default=0, help="cells per side per GPU (default: the metric's 8192; 512 with --dim 3)")
    ap.add_argument("--nx-global", type=int, default=0, help="fixed global rows (strong scaling) instead of --n rows per GPU")
    ap.add_argument("--ny", type=int, default=0, help="columns with --nx-global (default: square)")
    ap.add_argument("--ic", type=int, choices=[1, 2, 3], default=3)
    ap.add_argument("--preroll", type=int, default=-1, help="un-timed steps before the timed region (default 1000; 100 with --dim 3)")
    ap.add_argument("--transport", choices=["p2p", "nccl"], default="p2p", help="halo exchange: NVLink peer stores + device flags, or NCCL send/recv")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=16, help="row slabs of the streamed host-buffer step (N = 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-general", action="store_true", help="skip the general-path (random F) kernel timings")
    ap.add_argument("--no-parity", action="store_true", help="skip the slab-vs-single-GPU check of multi-GPU runs")
    args = ap.parse_args()
    if args.n == 0:
        args.n = 512 if args.dim == 3 else 8192
    if args.preroll < 0:
        args.preroll = 100 if args.dim == 3 else 1000
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
