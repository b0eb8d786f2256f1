import sys, os, json; sys.path.insert(0, ".")
from peakondrake_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), 'peakondrake_b200', 'libpeakon_b200_x.so')
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
for fam, nxi, sig, sch in (('sines', 1, 0.0, 'hydrodynamic'), ('constant', 1, 0.2, 'hydrodynamic'), ('sines', 4, 0.2, 'no_gradient')):
    for R in (592, 1024, 2048):
        eng = EnsembleIntegrator(1000, 40.0, 1e-3, R, num_Xis=nxi, Xi_family=fam, sigma=sig, seed=3, scheme=sch)
        eng.set_state(build_initial_u(1000, 40.0, 'one_peak')); eng.step(50); torch.cuda.synchronize()
        ms = []
        for _ in range(4):
            eng.step(50); ms.append(eng.last_kernel_ms())
        print("PK_FS=%s" % os.environ.get("PK_FS"), fam, sch, json.dumps(dict(R=R, ms=min(ms), rate=R*50/min(ms)*1e3, ctas=eng.layout.get('resident_ctas'))), flush=True)
        eng.close()
