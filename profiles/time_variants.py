"""
Times the headline workload (C2: ne=1000, 4 sines, R=1024, K=50) and BASELINE config 4 physics (ne=16384, clusters of 8,
R=128) for every library given on the command line, each in its own process:
  python profiles/time_variants.py default nodump nocoop ...      (names as built by profiles/build_variant.py)
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
def run(ne, R, K, fam, nxi, dt, ic, pw=1.0, reps=6):
    eng = EnsembleIntegrator(ne, 40.0, dt, R, num_Xis=nxi, Xi_family=fam, sigma=0.2, seed=3)
    eng.set_state(build_initial_u(ne, 40.0, ic, peak_width=pw))
    eng.step(K); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        eng.step(K); ms.append(eng.last_kernel_ms())
    eng.close()
    ms.sort()
    return R * K / ms[len(ms) // 2] * 1e3, R * K / ms[0] * 1e3
out = {'lib': os.environ.get('PK_LIB_PATH', 'default')}
out['C2_median'], out['C2_best'] = run(1000, 1024, 50, 'sines', 4, 1e-3, 'proper_peak')
if os.environ.get('WITH_C4', '1') == '1':
    out['C4_median'], out['C4_best'] = run(16384, 128, 10, 'constant', 1, 5e-4, 'new_peak', 1 / 6, reps=4)
    out['C4_R1024_median'], _ = run(16384, 1024, 10, 'constant', 1, 5e-4, 'new_peak', 1 / 6, reps=3)
    out['C5_R1024_median'], _ = run(20000, 1024, 10, 'constant', 1, 1e-3, 'periodic_peakon', reps=3)
    # fused per-step dumps (pk_step_dump) on the C4 mesh
    eng = EnsembleIntegrator(16384, 40.0, 5e-4, 1024, num_Xis=1, Xi_family='constant', sigma=0.2, seed=3)
    eng.set_state(build_initial_u(16384, 40.0, 'new_peak', peak_width=1 / 6))
    eng.step_dump(10, 1, suite=True); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        eng.step_dump(10, 1, suite=True); ms.append(eng.last_kernel_ms())
    out['C4_R1024_dump_every_step'] = 1024 * 10 / sorted(ms)[1] * 1e3
    eng.close()
print(json.dumps(out), flush=True)
''' % ROOT
for name in sys.argv[1:]:
    env = dict(os.environ)
    if name != 'default':
        env['PK_LIB_PATH'] = os.path.join(ROOT, 'peakondrake_b200', 'libpeakon_b200_%s.so' % name)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    print(name, r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
