"""
Throughput of the BASELINE configs other than the headline one (C3: ne=2000, C4: ne=16384,
C5: ne=20000), device-resident, CUDA events, Philox noise.  R is reduced to a few waves so that the
run stays short; realisations are independent, so the rate scales with full waves.
    python profiles/bench_configs.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u

HBM = 6465.2e9
CASES = [
    dict(name='C2 ne=1000 4 sines', ne=1000, R=1024, K=50, dt=1e-3, fam='sines', nxi=4, sigma=0.2, ic='proper_peak'),
    dict(name='C3 ne=2000 4 sines', ne=2000, R=592, K=20, dt=1e-3, fam='sines', nxi=4, sigma=0.2, ic='periodic_peakon'),
    dict(name='C4 ne=16384 constant', ne=16384, R=144, K=10, dt=5e-4, fam='constant', nxi=1, sigma=0.2, ic='new_peak'),
    dict(name='C5 ne=20000 constant sigma=1', ne=20000, R=72, K=10, dt=1e-3, fam='constant', nxi=1, sigma=1.0, ic='periodic_peakon'),
    dict(name='ne=1000 deterministic (C1)', ne=1000, R=1024, K=50, dt=1e-3, fam='sines', nxi=1, sigma=0.0, ic='one_peak'),
]
CASES += [
    dict(name='upwind/ssprk3 ne=1000 4 sines', ne=1000, R=4096, K=50, dt=1e-3, fam='sines', nxi=4, sigma=0.02, ic='proper_peak', scheme='upwind'),
    dict(name='upwind/ssprk3 ne=20 7 sines (test_dumping)', ne=20, R=65536, K=50, dt=1e-3, fam='sines', nxi=7, sigma=0.25, ic='one_peak', scheme='upwind'),
]
for c in CASES:
    up = c.get('scheme') == 'upwind'
    eng = EnsembleIntegrator(c['ne'], 40.0, c['dt'], c['R'], num_Xis=c['nxi'], Xi_family=c['fam'], sigma=c['sigma'], seed=3,
                             scheme=c.get('scheme', 'hydrodynamic'), timestepping='ssprk3' if up else 'midpoint')
    eng.set_state(build_initial_u(c['ne'], 40.0, c['ic'], peak_width=1 / 6))
    eng.step(c['K']); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        eng.step(c['K']); ms.append(eng.last_kernel_ms())
    t = min(ms) * 1e-3
    rate = c['R'] * c['K'] / t
    balg = (32 if up else 16) * c['ne'] + 8 * c['nxi']
    its = eng.its_total.double().mean().item() / (4 * c['K'])
    print(json.dumps(dict(config=c['name'], R=c['R'], steps_per_launch=c['K'], kernel_ms=min(ms), rate=rate,
                          node_steps_per_s=rate * c['ne'], hbm_frac=rate * balg / HBM, newton_its_per_step=its,
                          failed=int((eng.fail_step >= 0).sum().item()), layout=eng.layout)))
    eng.close()
