"""
The BASELINE configurations at their FULL sizes through the public surface (`experiment()`), on all the GPUs of the job:

  python profiles/run_full_configs.py [--dir DIR]                                     one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
         profiles/run_full_configs.py                                                 8 GPUs (seeds sharded over the ranks)

C2: 1024 realisations, ne=1000, 4 sines, 1000 steps, dumps every 50 steps
C3: 65 536 realisations, ne=2000, periodic peakon, 4 sines, 1000 steps, moments + histograms all-reduced every 100 steps
C4: 4096 realisations, ne=16 384, new_peak width 1/6, constant Xi, dt=5e-4, 200 steps, min/max u_x, jump, peakon suite EVERY step
C5: 8192 realisations, ne=20 000, periodic peakon, constant Xi sigma=1, tmax=0.1: the true peakon path (SDE at dt=1e-6 per
    seed) and the seven members of the dt sweep with nXi_updates=int(dt/1e-6), u_error_with_sde against the per-seed path
In-kernel Philox increments (keyed by the global realisation index).  Prints one JSON line per configuration (rank 0):
wall seconds of the experiment() call, realisation-steps/s through the drop-in, failures, a few statistics.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dir', default='/tmp/pk_full')
    ap.add_argument('--scale', type=float, default=1.0, help='fraction of the realisation counts (1.0 = BASELINE sizes)')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from peakondrake_b200 import experiment, outputting as out
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank() if world > 1 else 0
    os.makedirs(args.dir, exist_ok=True)
    os.chdir(args.dir)
    tag = 'w%d_' % world
    sc = lambda n: max(world, int(round(n * args.scale)))
    common = dict(schemes='hydrodynamic', timesteppings='midpoint', alphasq=1.0, c0=0., gamma=0., allow_fail=True,
                  rng='philox', philox_seed=2026)

    def timed(name, nreal, nsteps, fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt_ = time.time() - t0
        if rank == 0:
            return dict(config=name, gpus=world, realisations=nreal, steps=nsteps, seconds=dt_, realisation_steps_per_s=nreal * nsteps / dt_)

    def report(rec, code, extra):
        if rank == 0:
            d = out.read_data_file('results/%s/data.nc' % code)
            rec.update(extra(d))
            rec['failed'] = int(np.isfinite(d['failed_time']).sum())
            print(json.dumps(rec), flush=True)

    # ---- C2 ----
    R = sc(1024)
    rec = timed('C2', R, 1000, lambda: experiment(tag + 'c2', 40., 1.0, resolutions=1000, dts=1e-3, sigmas=0.2, seeds=range(R),
                                                  ics='proper_peak', num_Xis=4, Xi_family='sines',
                                                  diagnostics=['energy', 'max_du', 'min_du', 'q_pde'], ndump=50, ensemble_stats=True, **common))
    report(rec, tag + 'c2', lambda d: dict(dumps=len(d['time']), ens_count_last=float(d['ens_count'][-1]),
                                           min_du_ensemble_min=float(d['ens_min_du_min'][-1])))
    # ---- C3 ----
    R = sc(65536)
    rec = timed('C3', R, 1000, lambda: experiment(tag + 'c3', 40., 1.0, resolutions=2000, dts=1e-3, sigmas=0.2, seeds=range(R),
                                                  ics='periodic_peakon', num_Xis=4, Xi_family='sines',
                                                  diagnostics=['q_pde', 'min_du'], ndump=100, ensemble_stats=True, **common))
    report(rec, tag + 'c3', lambda d: dict(dumps=len(d['time']), ens_count_last=float(d['ens_count'][-1]),
                                           hist_q_total=float(d['ens_hist_q_pde'][-1].sum()), var_u_max=float(d['ens_var_u'][-1].max())))
    # ---- C4 ----
    R = sc(4096)
    rec = timed('C4', R, 200, lambda: experiment(tag + 'c4', 40., 0.1, resolutions=16384, dts=5e-4, sigmas=0.2, seeds=range(R),
                                                 ics='new_peak', num_Xis=1, Xi_family='constant', peak_width=1 / 6,
                                                 diagnostics=['max_du', 'min_du', 'max_jump_local', 'peakon_suite'], ndump=1, **common))
    report(rec, tag + 'c4', lambda d: dict(dumps=len(d['time']), min_du_min=float(np.nanmin(d['min_du'])),
                                           peakon_nu_mean_last=float(np.nanmean(d['peakon_nu'][-1]))))
    # ---- C5 ----
    R = sc(8192)
    c5 = dict(resolutions=20000, sigmas=1.0, seeds=range(R), ics='periodic_peakon', num_Xis=1, Xi_family='constant', periodic=True)
    c5.update(common)
    nt = 100000
    rec = timed('C5 true peakon paths (SDE, dt=1e-6)', R, nt,
                lambda: experiment(tag + 'true_peakon_data', 40., 0.1, dts=1e-6, diagnostics=['l2_u'], fields_to_output=[], ndump=nt,
                                   field_ndump=-1, peakon_equations=True, only_peakons=True, **c5))     # (true_peakon_sde.py also asks for
                # .pvd dumps, field_ndump=nt: one 2 MB ASCII .vtu per realisation at t=0 -- 16 GB for 8192 seeds; left out here)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    rms = {}
    for i, dt in enumerate([1e-2, 5e-3, 2.5e-3, 1e-3, 5e-4, 2.5e-4, 1e-4]):
        nst = int(round(0.1 / dt))
        rec = timed('C5 dt=%g' % dt, R, nst,
                    lambda: experiment(tag + 'strong_%d' % i, 40., 0.1, dts=dt, diagnostics=['u_error_with_sde', 'p_pde', 'q_pde'],
                                       ndump=nst, true_peakon_data=tag + 'true_peakon_data', nXi_updates=int(round(dt / 1e-6)), **c5))
        def extra(d, dt=dt):
            e = d['u_error_with_sde'][1]
            rms[dt] = float(np.sqrt(np.nanmean(e ** 2)))
            return dict(rms_u_error_with_sde=rms[dt], mean_u_error_with_sde=float(np.nanmean(e)))
        report(rec, tag + 'strong_%d' % i, extra)
    if rank == 0:
        dts = sorted(rms)
        slope = float(np.polyfit(np.log(dts), np.log([rms[d] for d in dts]), 1)[0])
        print(json.dumps(dict(config='C5 strong order over %d realisations' % R, fitted_order=slope)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
