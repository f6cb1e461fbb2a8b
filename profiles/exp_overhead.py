"""How much of the bench's ms/step is instrumentation?  8192^2, -ic 3: eager steps, eager steps with per-kernel events, graph replay."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, scaled_params
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(8192), stream=stream); s.set_init_F(3)
for _ in range(6):
    s.step()
s.synchronize()
def timed(fn, k=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(stream); fn(k); b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
def eager(k):
    for _ in range(k):
        s.step()
for rep in range(2):
    print(f"eager steps:            {timed(eager):.4f} ms/step")
    s.profile(True); print(f"eager + kernel events:  {timed(eager):.4f} ms/step"); s.profile(False)
    s.run(2); print(f"graph replay (run):     {timed(lambda k: s.run(k)):.4f} ms/step", flush=True)
