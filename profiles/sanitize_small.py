"""Small invocations of every kernel family, meant for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_small.py
NOTE (round 1): compute-sanitizer is closed on this GPU pool ("runs under it have left GPUs needing a reset"), so the
race / bounds evidence of this repo is the CPU emulation harness (tests/emu: identical results for ascending and
descending thread order between barriers, block-local-store checks for the cluster layout) plus the parity tests."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u, ensemble as ens
for ne, R, kw in ((1000, 6, dict(num_Xis=4, sigma=0.2)), (200, 5, dict(num_Xis=2, sigma=0.2)),
                  (300, 3, dict(num_Xis=1, Xi_family='constant', sigma=0.3)),
                  (1000, 4, dict(num_Xis=3, sigma=0.1, scheme='test')),
                  (1000, 4, dict(num_Xis=3, sigma=0.02, scheme='upwind', timestepping='ssprk3'))):
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, R, seed=1, **kw)
    eng.set_state(build_initial_u(ne, 40.0, 'proper_peak'))
    eng.step(7); eng.step(3)
    eng.diagnostics(want_m=not eng.upwind)
    if not eng.upwind:
        eng.peakon_suite(); eng.mu()
        ens.EnsembleStatistics(eng, nbins=16, lo=0.0, hi=40.0).update(False)
    torch.cuda.synchronize()
    print('ok', ne, kw, float(eng.u.abs().max()))
    eng.close()
os.environ['PK_FORCE_CLUSTER'] = '2'
eng = EnsembleIntegrator(1000, 40.0, 1e-3, 2, num_Xis=2, sigma=0.2, seed=1)
eng.set_state(build_initial_u(1000, 40.0, 'proper_peak')); eng.step(3); torch.cuda.synchronize()
print('ok cluster', float(eng.u.abs().max()))
