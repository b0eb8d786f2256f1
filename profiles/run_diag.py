"""pk_diag_kernel on a C4-sized table (4096 rows of 16384 nodes) for profiling: python profiles/run_diag.py"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
eng = EnsembleIntegrator(16384, 40.0, 5e-4, 4096, num_Xis=1, Xi_family='constant', sigma=0.2, seed=3)
eng.set_state(build_initial_u(16384, 40.0, 'new_peak', peak_width=1 / 6))
eng.u.add_(torch.randn_like(eng.u) * 1e-3)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.diagnostics(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('diagnostics of 4096 x 16384: %.3f ms = %.0f GB/s' % (ms, 4096 * 16384 * 8 / ms / 1e6))
