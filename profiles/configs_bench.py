"""Timings of the BASELINE.json configs that are not the headline bench line (configs 1, 2, 5), GPU vs the CPU
oracle port on this box's host cores.  `python profiles/configs_bench.py [which ...]`  (which in: c1 c2 c5)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.c_oracle import Vof2DCOracle, Vof3DCOracle
from oracle.vof2d_oracle import Vof2DParams
from oracle.vof3d_oracle import Vof3DParams
from taichi_2d_vof_b200 import VofSolver2D, VofSolver3D, reference_params, scaled_params3d

which = sys.argv[1:] or ["c1", "c2", "c5"]
out = {}


def gpu_time(fn, sync):
    fn(); sync()
    t0 = time.perf_counter(); fn(); sync()
    return time.perf_counter() - t0


if "c1" in which:   # dam break, 200^2, 100 steps, headless (PR1 reference)
    s = VofSolver2D(reference_params()); s.set_init_F(1)
    s.run(100); s.synchronize()
    t = gpu_time(lambda: s.run(1000), s.synchronize) / 1000
    o = Vof2DCOracle(Vof2DParams()); o.set_init_F(1); o.run(20)
    t0 = time.perf_counter(); o.run(300); tc = (time.perf_counter() - t0) / 300
    out["c1_dam_break_200x200"] = {"gpu_steps_per_s": 1 / t, "cpu_steps_per_s": 1 / tc, "cpu_threads": Vof2DCOracle.threads(),
                                    "gpu_mode": "vof2d_run (CUDA-graph replay of 2-step graphs)"}
if "c2" in which:   # rising bubble, 2048^2, reference constants unchanged
    P = Vof2DParams(nx=2048, ny=2048)
    s = VofSolver2D(reference_params(nx=2048, ny=2048)); s.set_init_F(2)
    s.run(10); s.synchronize()
    t = gpu_time(lambda: s.run(100), s.synchronize) / 100
    o = Vof2DCOracle(P); o.set_init_F(2); o.run(2)
    t0 = time.perf_counter(); o.run(20); tc = (time.perf_counter() - t0) / 20
    d = s.diagnostics()
    out["c2_rising_bubble_2048x2048"] = {"gpu_steps_per_s": 1 / t, "cpu_steps_per_s": 1 / tc, "cpu_threads": Vof2DCOracle.threads(),
                                          "gpu_gcell_updates_per_s": 10 * 2048 * 2048 / t / 1e9, "state_finite": bool(np.isfinite(d["mass"]))}
if "c5" in which:   # 3-D dam break at 512^3 (single GPU here; slabs over 8 GPUs use the same kernels)
    n = 512
    s = VofSolver3D(scaled_params3d(n)); s.set_init_F(1)
    s.run(3); s.synchronize()
    t = gpu_time(lambda: s.run(10), s.synchronize) / 10
    nc = 256
    o = Vof3DCOracle(Vof3DParams.scaled(nc)); o.set_init_F(1); o.run(1)
    t0 = time.perf_counter(); o.run(3); tc = (time.perf_counter() - t0) / 3
    out["c5_dam_break_3d_512"] = {"gpu_steps_per_s": 1 / t, "gpu_gcell_updates_per_s": 10 * n ** 3 / t / 1e9,
                                  "cpu_gcell_updates_per_s": 10 * nc ** 3 / tc / 1e9, "cpu_sample": f"{nc}^3, 3 steps",
                                  "gpu_mcell_steps_per_s": n ** 3 / t / 1e6, "mass": s.mass()}
print(json.dumps(out, indent=1))
