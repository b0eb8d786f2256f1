"""upwind/ssprk3 throughput case (ne=1000, 4 sines, R=4096) for profiling: python profiles/run_upwind.py [R] [K]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
eng = EnsembleIntegrator(1000, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.02, seed=3, scheme='upwind', timestepping='ssprk3')
eng.set_state(build_initial_u(1000, 40.0, 'proper_peak'))
eng.step(K); torch.cuda.synchronize()
for _ in range(3):
    eng.step(K); print('kernel ms', eng.last_kernel_ms(), 'rate', R * K / eng.last_kernel_ms() * 1e3, eng.layout)
