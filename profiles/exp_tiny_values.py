import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from taichi_2d_vof_b200 import VofSolver2D, scaled_params
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(8192), stream=stream); s.set_init_F(3)
def t_jac():
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); s.solve_p_jacobi(5); b.record(stream); torch.cuda.synchronize(); return a.elapsed_time(b)
for n in (1,2,3,5,10,20,30):
    while s.istep < n: s.step()
    p = s.p.to_numpy()[1:-1,1:-1]; ap = np.abs(p)
    tiny = np.count_nonzero((ap>0)&(ap<7.9e-31)); sub = np.count_nonzero((ap>0)&(ap<1.1754944e-38)); nz=np.count_nonzero(ap)
    rows_tiny = np.count_nonzero(((ap>0)&(ap<7.9e-31)).any(axis=1)); rows_tiny_c = np.count_nonzero(((ap>0)&(ap<7.9e-31)).any(axis=0))
    ms = t_jac()
    print(f"step {n}: nonzero {nz/p.size:.3f} tiny {tiny/p.size:.4f} subnormal {sub/p.size:.4f} rows-with-tiny {rows_tiny} cols-with-tiny {rows_tiny_c}  jacobi(5)+rhs {ms:.3f} ms")
