"""
Cluster-kernel layouts on the C4 (ne=16384) and C5 (ne=20000) meshes: more, thinner threads (4096 x 4 rows in CTAs of 512 or
in clusters of 16) against the default 2048 x 8 / 2048 x 10.  Each layout in its own process; the final state of every layout
is compared with the default's (same Philox increments).
  python profiles/exp_cluster_layouts.py
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, numpy as np
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
ne, dt, ic, sigma = (16384, 5e-4, 'new_peak', 0.2) if os.environ['MESH'] == 'c4' else (20000, 1e-3, 'periodic_peakon', 1.0)
R, K = 1024, 10
eng = EnsembleIntegrator(ne, 40.0, dt, R, num_Xis=1, Xi_family='constant', sigma=sigma, seed=3)
eng.set_state(build_initial_u(ne, 40.0, ic, peak_width=1 / 6))
eng.step(K); torch.cuda.synchronize()
ms = []
for _ in range(3):
    eng.step(K); ms.append(eng.last_kernel_ms())
u = eng.u[:8].cpu().numpy()
np.save(os.environ['OUT'], u)
print(json.dumps(dict(layout=eng.layout, rate=R * K / sorted(ms)[1] * 1e3, its=float(eng.its_total.double().mean().item()),
                      failed=int((eng.fail_step >= 0).sum().item()))), flush=True)
eng.close()
''' % ROOT
import numpy as np
LIB = lambda n: {'PK_LIB_PATH': os.path.join(ROOT, 'peakondrake_b200', 'libpeakon_b200_%s.so' % n)}
CASES = [('c4', 'default', {}), ('c4', 'generic, 1 cluster/SM set', {'PK_CLUSTER_GENERIC': '1', 'PK_CLUSTER_MINB': '1'}),
         ('c4', 'generic, clusters of 16 x 128 threads', {'PK_CLUSTER_GENERIC': '1', 'PK_FORCE_CLUSTER': '16'}),
         ('c5', 'generic, clusters of 16 x 128 threads', {'PK_CLUSTER_GENERIC': '1', 'PK_FORCE_CLUSTER': '16'}), ('c4', 'one cluster per SM set (255 registers, two row buffers)', {'PK_CLUSTER_MINB': '1'}),
         ('c5', 'one cluster per SM set (255 registers, two row buffers)', {'PK_CLUSTER_MINB': '1'}),
         ('c4', 'fast reciprocal', LIB('fastr')), ('c5', 'fast reciprocal', LIB('fastr')),
         ('c4', 'relaxed cluster barrier', LIB('relaxed')), ('c5', 'relaxed cluster barrier', LIB('relaxed')),
         ('c4', 'edges pushed AND cluster barriers kept', LIB('edgebar')), ('c5', 'edges pushed AND cluster barriers kept', LIB('edgebar')), ('c4', 'runtime shared-memory offsets (round-2 midpoint)', {'PK_CLUSTER_GENERIC': '1'}),
         ('c5', 'runtime shared-memory offsets', {'PK_CLUSTER_GENERIC': '1'}), ('c4', '4096x4 in 8 CTAs of 512', {'PK_FORCE_TM': '4096,5', 'PK_FORCE_CLUSTER': '8'}),
         ('c4', '4096x4 in 16 CTAs of 256', {'PK_FORCE_TM': '4096,5'}),
         ('c5', 'default (2048x10)', {}), ('c5', '4096x5 in 16 CTAs of 256', {'PK_NO_WIDE_CHUNKS': '1'}),
         ('c5', '4096x5 in 8 CTAs of 512', {'PK_NO_WIDE_CHUNKS': '1', 'PK_FORCE_CLUSTER': '8'})]
if os.environ.get('QUICK', '0') == '1':
    CASES = [c for c in CASES if 'CTAs of' not in c[1]]
CASES = [c for c in CASES if 'PK_LIB_PATH' not in c[2] or os.path.exists(c[2]['PK_LIB_PATH'])]
CASES.sort(key=lambda c: c[0])
ref = {}
for mesh, name, envx in CASES:
    env = dict(os.environ); env.update(envx); env['MESH'] = mesh; env['OUT'] = '/tmp/exp_cl_%s.npy' % mesh
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip() or r.stderr.strip()[-600:]
    diff = None
    if r.returncode == 0:
        u = np.load(env['OUT'])
        if mesh not in ref: ref[mesh] = u
        diff = float(np.max(np.linalg.norm(u - ref[mesh], axis=1) / np.linalg.norm(ref[mesh], axis=1)))
    print(mesh, name, line, 'rel diff vs default', diff, flush=True)
