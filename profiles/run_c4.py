"""BASELINE config 4 physics (ne=16384, constant family, clusters of 8) for profiling: python profiles/run_c4.py [R] [K]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
R = int(sys.argv[1]) if len(sys.argv) > 1 else 72
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = EnsembleIntegrator(16384, 40.0, 5e-4, R, num_Xis=1, Xi_family='constant', sigma=0.2, seed=3)
eng.set_state(build_initial_u(16384, 40.0, 'new_peak', peak_width=1 / 6))
eng.step(K); torch.cuda.synchronize()
for _ in range(3):
    eng.step(K); print('kernel ms', eng.last_kernel_ms(), 'rate', R * K / eng.last_kernel_ms() * 1e3, eng.layout)
