"""A/B of the per-thread code variants (Thread<..., VAR> in pk_core.h) on the bench workload; needs a library
built with -DPK_EXPERIMENTAL_VARIANTS:  PK_FS=0|1|2|4 python profiles/exp_fs.py
  0 default (unrolled Newton pass, rolled residual-only pass)   1 FS: factors in shared memory, rolled loops, 2 CTAs/SM
  2 rolled assembly, register factors picked by a switch        4 everything unrolled
Result (profiles/r1d_exp_variants.log): R=1024: 13.97e6 / 11.26e6 / 10.86e6 / 13.23e6 realisation-steps/s."""
import sys, json, os; sys.path.insert(0, ".")
import torch
from peakondrake_b200 import EnsembleIntegrator, build_initial_u
ne = int(os.environ.get("NE", 1000))
u0 = build_initial_u(ne, 40.0, 'proper_peak')
ref = None
for R in (148, 296, 444, 1024, 2048):
    eng = EnsembleIntegrator(ne, 40.0, 1e-3, R, num_Xis=4, Xi_family='sines', sigma=0.2, seed=3)
    eng.set_state(u0); eng.step(50); torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        eng.step(50); ms.append(eng.last_kernel_ms())
    print("PK_FS=%s" % os.environ.get("PK_FS"), json.dumps(dict(ne=ne, R=R, ms=min(ms), rate=R*50/min(ms)*1e3, lay=eng.layout,
          its=eng.its_total.double().mean().item()/250, chk=float(eng.u[0].sum()))), flush=True)
    eng.close()
