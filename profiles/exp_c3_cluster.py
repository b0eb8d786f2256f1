import sys, json, os; sys.path.insert(0, ".")
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
ne = 2000
u0 = build_initial_u(ne, 40.0, 'periodic_peakon')
for R in (148, 296, 592, 1184):
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.2, seed=3)
    eng.set_state(u0); eng.step(20); torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        eng.step(20); ms.append(eng.last_kernel_ms())
    print("FORCE=%s" % os.environ.get("PK_FORCE_CLUSTER"), json.dumps(dict(ne=ne, R=R, ms=min(ms), node_steps=R*20*ne/min(ms)*1e3, lay=eng.layout)), flush=True)
    eng.close()
