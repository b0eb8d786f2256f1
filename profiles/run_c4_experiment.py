"""BASELINE config 4 at full size through experiment() on one GPU: python profiles/run_c4_experiment.py [dir]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from peakondrake_b200 import experiment, outputting as out
d = sys.argv[1] if len(sys.argv) > 1 else '/tmp/pk_c4'
os.makedirs(d, exist_ok=True); os.chdir(d)
for tag in ('warm', 'timed'):
    R = 64 if tag == 'warm' else 4096
    torch.cuda.synchronize(); t0 = time.time()
    experiment('c4_' + tag, 40., 0.1, resolutions=16384, dts=5e-4, sigmas=0.2, seeds=range(R), schemes='hydrodynamic',
               timesteppings='midpoint', alphasq=1.0, c0=0., gamma=0., allow_fail=True, rng='philox', philox_seed=2026,
               ics='new_peak', num_Xis=1, Xi_family='constant', peak_width=1 / 6,
               diagnostics=['max_du', 'min_du', 'max_jump_local', 'peakon_suite'], ndump=1)
    torch.cuda.synchronize(); dt_ = time.time() - t0
    f = out.read_data_file('results/c4_%s/data.nc' % tag)
    print(json.dumps(dict(config='C4 through experiment()', realisations=R, steps=200, seconds=dt_, realisation_steps_per_s=R * 200 / dt_,
                          dumps=len(f['time']), min_du_min=float(np.nanmin(f['min_du'])), failed=int(np.isfinite(f['failed_time']).sum()))), flush=True)
