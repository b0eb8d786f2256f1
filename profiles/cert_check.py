"""
Checking build of the Newton convergence certificate: the library is compiled with -DPK_CERT_CHECK
(profiles/build_variant.py certcheck -DPK_CERT_CHECK, PK_WITH_M10=1), in which every time the certificate decides
SNES's `|R| <= rtol |R0|` test without evaluating R, the residual IS evaluated and the realisation fails (reason 9) if the
test would not have passed, or (reason 8) if |R| exceeds the bound the certificate computed.  Runs the BASELINE
configurations, the test scheme and forced-cluster layouts and reports failures (there must be none) and how often the
certificate fired (Newton totals against a PK_NO_CERTIFICATE run must be identical).
    PK_LIB_PATH=peakondrake_b200/libpeakon_b200_certcheck.so python profiles/cert_check.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u

CASES = [
    ('C1', dict(ne=1000, R=8, K=200, dt=1e-3, ic='one_peak', fam='sines', nxi=1, sigma=0.0)),
    ('C2', dict(ne=1000, R=256, K=300, dt=1e-3, ic='proper_peak', fam='sines', nxi=4, sigma=0.2)),
    ('C3', dict(ne=2000, R=128, K=200, dt=1e-3, ic='periodic_peakon', fam='sines', nxi=4, sigma=0.2)),
    ('C4', dict(ne=16384, R=32, K=100, dt=5e-4, ic='new_peak', fam='constant', nxi=1, sigma=0.2, pw=1 / 6)),
    ('C5 dt=1e-2', dict(ne=20000, R=16, K=10, dt=1e-2, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0)),
    ('C5 dt=1e-3', dict(ne=20000, R=16, K=100, dt=1e-3, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0)),
    ('C5 dt=1e-4', dict(ne=20000, R=16, K=200, dt=1e-4, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0)),
    ('test scheme', dict(ne=1000, R=32, K=100, dt=1e-3, ic='proper_peak', fam='sines', nxi=4, sigma=0.2, scheme='test')),
    ('forced cluster of 4', dict(ne=1000, R=32, K=100, dt=1e-3, ic='proper_peak', fam='sines', nxi=4, sigma=0.2, force=4)),
    ('no_gradient', dict(ne=1000, R=32, K=100, dt=1e-3, ic='proper_peak', fam='sines', nxi=4, sigma=0.2, scheme='no_gradient')),
]


def run(c, nocert):
    os.environ['PK_NO_CERTIFICATE'] = '1' if nocert else '0'
    if c.get('force'):
        os.environ['PK_FORCE_CLUSTER'] = str(c['force'])
    else:
        os.environ.pop('PK_FORCE_CLUSTER', None)
    eng = EnsembleIntegrator(c['ne'], 40.0, c['dt'], c['R'], num_Xis=c['nxi'], Xi_family=c['fam'], sigma=c['sigma'], seed=11,
                             scheme=c.get('scheme', 'hydrodynamic'))
    eng.set_state(build_initial_u(c['ne'], 40.0, c['ic'], peak_width=c.get('pw', 1.0)))
    eng.step(c['K'])
    torch.cuda.synchronize()
    res = dict(failed=int((eng.fail_step >= 0).sum().item()), its=eng.its_total.cpu().numpy().copy(), u=eng.u.cpu().numpy().copy())
    eng.close()
    return res


for name, c in CASES:
    a, b = run(c, False), run(c, True)
    print(json.dumps({'case': name, 'failed_with_check': a['failed'], 'failed_without_certificate': b['failed'],
                      'newton_totals_equal': bool(np.array_equal(a['its'], b['its'])), 'states_equal': bool(np.array_equal(a['u'], b['u'])),
                      'mean_its_per_step': float(a['its'].mean() / c['K'])}), flush=True)
