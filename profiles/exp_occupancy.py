"""
Occupancy / lock-step probe of pk_step_kernel on the bench workload (ne=1000, 4 sines):
  * R = 148, 296, 444 (1, 2, 3 resident CTAs per SM, one wave), 888, 1024, 1332 at K=50
  * K sweep at R=444: CTAs of one wave start together, so the first steps run in near lock-step
    (shared instruction fetch); a lower per-step time at small K than at large K is evidence that the
    instruction supply, not the fp64 pipe, limits the drifted steady state.
    python profiles/exp_occupancy.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u

u0 = build_initial_u(1000, 40.0, 'proper_peak')


def run(R, K, reps=5):
    eng = EnsembleIntegrator(1000, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.2, seed=3)
    eng.set_state(u0)
    eng.step(K); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        eng.step(K); ms.append(eng.last_kernel_ms())
    eng.close()
    return min(ms)


for R in (148, 296, 444, 888, 1024, 1332, 2048):
    ms = run(R, 50)
    print(json.dumps(dict(R=R, K=50, ms=ms, rate=R * 50 / ms * 1e3, us_per_step=ms * 1e3 / 50)), flush=True)
prev = None
for K in (1, 2, 4, 8, 16, 32, 64, 128):
    ms = run(444, K, reps=7)
    d = dict(R=444, K=K, ms=ms, us_per_step=ms * 1e3 / K)
    if prev:
        d['incremental_us_per_step'] = (ms - prev[1]) * 1e3 / (K - prev[0])
    prev = (K, ms)
    print(json.dumps(d), flush=True)
