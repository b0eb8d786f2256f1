"""e2e through pk_step_host (pinned host buffers, C2 shard) for 1..4 pipeline chunks: python profiles/exp_host_chunks.py"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path.insert(0, %r)
import numpy as np, torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
R, NE, K = 1024, 1000, 50
eng = EnsembleIntegrator(NE, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.2, seed=2026)
u0 = build_initial_u(NE, 40.0, 'proper_peak')
pin = lambda shape, dt_: torch.zeros(shape, dtype=dt_).pin_memory().numpy()
hu = pin((R, NE), torch.float64); hu[:] = u0
hfs = pin((R,), torch.int32); hfs[:] = -1
hil, hit = pin((R,), torch.int32), pin((R,), torch.int32)
for _ in range(3):
    eng.step_host(K, hu, None, None, None, None, hfs, hil, hit)
t0 = time.perf_counter()
n = 30
for _ in range(n):
    eng.step_host(K, hu, None, None, None, None, hfs, hil, hit)
dt = time.perf_counter() - t0
eng.set_state(u0); eng.step(K); torch.cuda.synchronize()
ms = []
for _ in range(5):
    eng.step(K); ms.append(eng.last_kernel_ms())
print(json.dumps({'nchunk': os.environ.get('PK_HOST_NCHUNK', 'default'), 'e2e': R * K * n / dt, 'kernel_only': R * K / min(ms) * 1e3}))
''' % ROOT
for nc in ('1', '2', '3', '4'):
    env = dict(os.environ, PK_HOST_NCHUNK=nc)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-300:], flush=True)
