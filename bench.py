#!/usr/bin/env python
"""Benchmark of the 2-D VOF per-timestep hot path (BASELINE.json: timesteps/s and Jacobi
Gcell-updates/s at 8192^2; HBM GB/s as % of peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 8192]

One "step" = one pass of the loop body 2dvof.py:513-528 (props, normals/curvature, advection,
BC, 10 Jacobi sweeps, projection, BC, FCT x/y, post-process, BC) over the whole grid.
Workload (configs[2] of BASELINE.json, the configuration the metric is quoted on): dropping
liquid (-ic 3) at 8192^2 per GPU, fp32, constant-dx scaling (L = 0.1 * n / 200 so that
nu*dt/dx^2 stays at the reference's stable value; SURVEY.md 7 risk 3), synthetic = generated
by the solver's own set_init_F.  For N > 1 the domain is (8192 N) x 8192, row-slab decomposed
(weak scaling), one halo exchange per step.

Prints ONE JSON line (rank 0).  `value` = Jacobi cell-updates/s of the whole job with state
resident in HBM (= n_jacobi * cells * steps / time, the definition in BASELINE.md 3.4);
`timesteps_per_s` rides along.  `e2e` = the same through vof2d_step_host with pinned HOST
buffers (H2D of u,v,p,F + step + D2H of u,v,p,F inside the timed region).
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
UNIT = "Gcell-updates/s"
# algorithmic bytes per cell per launch (SURVEY.md 8d / DESIGN.md): fp32 arrays read + written once
ALGO_BYTES = {"kappa": 8, "advect": 24, "rhs": 16, "jacobi": 12, "project": 24, "fct_x": 12, "fct_y": 12, "props": 12}


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
def time_cpu_port(n, steps, warmup, ic=3):
    from oracle.c_oracle import Vof2DCOracle
    from oracle.vof2d_oracle import Vof2DParams
    P = Vof2DParams.scaled(n)
    o = Vof2DCOracle(P)
    o.set_init_F(ic)
    o.run(warmup)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    cells = n * n
    return {"seconds": dt, "steps_per_s": steps / dt, "gcell_updates_per_s": N_JACOBI * cells * steps / dt / 1e9,
            "threads": Vof2DCOracle.threads(), "n": n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the full 8192^2 grid when the whole run stays within a few minutes
    # (~2 s/step on 16 cores), else a 4096^2 grid of the same workload (throughput per cell is
    # size-independent once the working set is far beyond the caches)
    n = args.n if (args.steps + args.warmup) <= 40 else min(args.n, 4096)
    r = time_cpu_port(n, args.steps, args.warmup)
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) of -ic 3 at {n}^2, C/OpenMP restatement of 2dvof.py "
              f"(taichi 1.4.1 not installable: py3.12, offline), {r['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gcell_updates_per_s"], "unit": UNIT,
        "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "timesteps_per_s": r["steps_per_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"dropping liquid (-ic 3) at {n}^2, constant-dx scaling, {N_JACOBI} Jacobi sweeps/step",
                   "grid": [n, n]},
        "cpu_baseline": {"value": r["gcell_updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": r["gcell_updates_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    from taichi_2d_vof_b200 import VofSolver2D, scaled_params
    from taichi_2d_vof_b200.slab import SlabSolver2D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    n = args.n
    ny = n
    nx_global = n * world                      # weak scaling: 8192 rows per GPU
    scaling = "weak"
    if args.nx_global:                         # fixed global grid (BASELINE config 4: 32768^2 over 2/4/8 GPUs)
        nx_global, ny, scaling = args.nx_global, (args.ny or args.nx_global), "strong"
    L_y, L_x = 0.1 * ny / 200.0, 0.1 * nx_global / 200.0

    def params_fn(slab, halo, device):
        from taichi_2d_vof_b200 import reference_params
        return reference_params(nx=nx_global, ny=ny, Lx=L_x, Ly=L_y, n_jacobi=N_JACOBI, slab=slab, halo=halo, device=device)

    slab = SlabSolver2D(params_fn, nx_global, rank, world, dist=dist, n_jacobi=N_JACOBI, device=local, transport=args.transport)
    s = slab.solver
    slab.set_init_F(args.ic)
    stream = slab.stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        slab.step()
    barrier()

    # ---- timed region: K steps, state resident in HBM, per-kernel events on the ctx stream
    prof_every = 4 if args.steps >= 8 else 1      # kernel events on every 4th step of the timed region: per-kernel
    s.profile(prof_every)                         # durations are measured live, 3/4 of the steps run uninstrumented
    l0 = s.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        slab.step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = s.launch_count() - l0
    prof = s.profile_read()
    s.profile(False)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())

    d = s.diagnostics()
    finite = np.isfinite(d["mass"]) and np.isfinite(d["max_cfl"])

    cells_per_gpu = (slab.hi - slab.lo + 1) * ny
    cells_total = nx_global * ny
    sec = ms / 1e3
    value = N_JACOBI * cells_total * args.steps / sec / 1e9
    peak, peak_src = measured_peak()

    # ---- roofline of the dominant kernel (Jacobi) + the others, from the live CUDA-event spans.
    # A "span" is one launch group of a kernel kind; the temporally blocked Jacobi does T = 5 sweeps per launch,
    # so its algorithmic bytes per launch are 12 B x cells x T.
    rows_local = s.nrows          # rows a launch actually processes (owned + redundant halo rows)
    kern = {}
    steps_sampled = len(range(0, args.steps, prof_every))
    for name, (tot_ms, nspan) in prof.items():
        if name not in ALGO_BYTES:
            continue
        per = tot_ms / nspan
        sweeps = (N_JACOBI * steps_sampled / nspan) if name == "jacobi" else 1.0
        algo = ALGO_BYTES[name] * cells_per_gpu * sweeps
        kern[name] = {"launches_per_step": nspan / steps_sampled, "ms_per_launch": per,
                      "algo_GBps": algo / (per * 1e-3) / 1e9, "share_of_step": (tot_ms / steps_sampled) / (ms / args.steps)}
        if name == "jacobi":
            kern[name]["sweeps_per_launch"] = sweeps
            kern[name]["gcell_updates_per_s"] = cells_per_gpu * sweeps / (per * 1e-3) / 1e9
    dom = "jacobi"
    roof = None
    if dom in kern:
        a = kern[dom]["algo_GBps"]
        T = kern[dom]["sweeps_per_launch"]
        roof = {"kernel": "k_jacobi_tb<%d> (%g sweeps per HBM pass, register-pipelined)" % (round(T), T), "bound": "hbm",
                "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "peak_source": peak_src,
                "algo_bytes_per_cell_update": 12, "cell_updates_per_launch": cells_per_gpu * T,
                "ms_per_launch": kern[dom]["ms_per_launch"], "traffic": None, "frac_of_nominal_8TBps": a / 8000.0,
                "note": "temporal blocking moves ~1/T of the un-blocked bytes, so the algorithmic rate may exceed the copy peak; "
                        "`traffic` is the DRAM bytes one launch really moved (ncu), see profiles/"}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f)
            roof["traffic"] = tj.get("jacobi_tb_bytes_per_launch")
            if roof["traffic"]:
                roof["dram_GBps"] = roof["traffic"] / (kern[dom]["ms_per_launch"] * 1e-3) / 1e9
                roof["dram_frac"] = roof["dram_GBps"] / peak
        except Exception:
            pass

    # ---- e2e: the same step through the host-buffer C-ABI call (rank-local slab, pinned memory)
    e2e = None
    if not args.no_e2e:
        shape = (s.nrows, ny + 2)
        host = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(4)]
        arrs = [h.numpy() for h in host]
        for a, k in zip(arrs, ("u", "v", "p", "F")):
            a[...] = getattr(s, k).to_numpy()
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
            for a, k in zip(arrs, ("u", "v", "p", "F")):
                getattr(s, k).from_numpy(a)
            slab.step()
            for a, k in zip(arrs, ("u", "v", "p", "F")):
                getattr(s, k).to_numpy(out=a)

        nbytes = 4 * shape[0] * shape[1] * 4
        if world == 1:
            # whole-array call first (explains the streamed number), then the streamed call = the e2e headline
            dt_plain = timed(lambda: s.step_host(*arrs))
            from taichi_2d_vof_b200 import VofStreamer2D
            n_slabs = max(1, min(args.e2e_slabs, nx_global // 32))
            st = VofStreamer2D(params_fn(None, 0, local), n_slabs=n_slabs)
            dt = timed(lambda: st.step_host(*arrs))
            up_rows = sum(min(nx_global + 1, nx_global * (k + 1) // n_slabs + st.halo) - max(0, 1 + nx_global * k // n_slabs - st.halo) + 1
                          for k in range(n_slabs)) if n_slabs > 1 else shape[0]
            e2e = {"value": N_JACOBI * cells_total * k_e2e / dt / 1e9, "unit": UNIT, "timesteps_per_s": k_e2e / dt,
                   "h2d_bytes_per_step": 4 * up_rows * shape[1] * 4, "d2h_bytes_per_step": nbytes, "steps": k_e2e,
                   "api": f"vof2d_streamer_step_host (pinned host u,v,p,F in and out; {n_slabs} row slabs, halo {st.halo}, "
                          "upload / step / download overlapped on three streams)",
                   "unstreamed": {"value": N_JACOBI * cells_total * k_e2e / dt_plain / 1e9, "timesteps_per_s": k_e2e / dt_plain,
                                  "api": "vof2d_step_host (whole arrays up, step, whole arrays down)"}}
            st.close()
        else:
            dt = timed(slab_roundtrip)
            e2e = {"value": N_JACOBI * cells_total * k_e2e / dt / 1e9, "unit": UNIT, "timesteps_per_s": k_e2e / dt,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": k_e2e,
                   "api": "per-rank slab: from_numpy(u,v,p,F), halo exchange + step, to_numpy(u,v,p,F) (pinned host)"}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = time_cpu_port(n if n <= 8192 else 8192, 24, 1)
        cpu = {"value": r["gcell_updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "timesteps_per_s": r["steps_per_s"],
               "sample": f"24 steps (+1 warm-up) of the same -ic 3 workload at {r['n']}^2, C/OpenMP restatement of 2dvof.py with the "
                         f"reference's loop structure, {r['threads']} threads ({r['seconds']:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "timesteps_per_s": args.steps / sec, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{IC_NAMES[args.ic]} (-ic {args.ic}), {slab.hi - slab.lo + 1} x {ny} cells per GPU (global {nx_global} x {ny}), constant-dx scaling "
                                   f"L = 0.1*n/200, dt = 4e-6, {N_JACOBI} Jacobi sweeps/step, fused step (vof2d_step)",
                       "grid_per_gpu": [slab.hi - slab.lo + 1, ny], "global_grid": [nx_global, ny],
                       "decomposition": ("row slabs along i, deep halo %d rows, 1 exchange/step, transport %s" % (s.halo, args.transport)) if world > 1 else "single GPU",
                       "l2": "inputs exceed L2 (10 live fp32 fields x %.0f MB >> 126 MB)" % (s.nrows * (ny + 2) * 4 / 1e6),
                       "rows_processed_per_launch": rows_local},
            "roofline": roof, "kernels": kern, "kernel_event_sampling": f"CUDA events around every kernel of every {prof_every}th step of the timed region" if prof_every > 1 else "CUDA events around every kernel of the timed region", "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk, "state_finite": bool(finite), "mass": d["mass"], "max_cfl": d["max_cfl"],
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
    ap.add_argument("--n", type=int, default=8192, help="cells per side per GPU (default: the metric's 8192)")
    ap.add_argument("--nx-global", type=int, default=0, help="fixed global rows (strong scaling) instead of --n rows per GPU")
    ap.add_argument("--ny", type=int, default=0, help="columns with --nx-global (default: square)")
    ap.add_argument("--ic", type=int, choices=[1, 2, 3], default=3)
    ap.add_argument("--transport", choices=["p2p", "nccl"], default="p2p", help="halo exchange: NVLink peer stores + device flags, or NCCL send/recv")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=16, help="row slabs of the streamed host-buffer step (N = 1)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
