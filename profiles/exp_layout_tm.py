"""Layout experiment (threads per realisation T x rows per thread M) on the bench physics:
  python profiles/exp_layout_tm.py; PK_FORCE_TM=256,5 NE=1280 python profiles/exp_layout_tm.py
The M=4 runs of profiles/r1c_exp_layout_tm.log used a temporary build with an M=4 translation unit
(uniform 4-row chunks only); result: T=256/M=4 is 15% slower than T=128/M=8 at ne=1024, T=128/M=4 is 17%
faster than T=64/M=8 at ne=512."""
import sys, json, os; sys.path.insert(0, ".")
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
ne = int(os.environ.get("NE", 1024))
u0 = build_initial_u(ne, 40.0, 'proper_peak')
for R in (296, 444, 1024, 2048):
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.2, seed=3)
    eng.set_state(u0); eng.step(50); torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        eng.step(50); ms.append(eng.last_kernel_ms())
    print(os.environ.get("PK_FORCE_TM"), json.dumps(dict(ne=ne, R=R, ms=min(ms), rate=R*50/min(ms)*1e3, lay=eng.layout, its=eng.its_total.double().mean().item()/250)), flush=True)
    eng.close()
