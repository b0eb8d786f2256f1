"""
Multi-GPU check of the drop-in (not a pytest file: needs N GPUs and a torchrun launch).

  python tests/mgpu_experiment.py --single --dir /tmp/mg           one process, one GPU: the reference files
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
         tests/mgpu_experiment.py --dir /tmp/mg                    N ranks: same experiments, then rank 0 compares

BASELINE config 3 at its parameters (ne=2000, periodic_peakon, 4 sines, statistics every 100 steps) with a reduced
realisation count, in-kernel Philox keyed by the global realisation id: the per-realisation rows of data.nc must be
bit-identical to the single-GPU file, the all-reduced statistics equal to rounding of the sums; plus a run on the
reference's MT19937 streams with the fused per-step dumps and the device peakon SDE.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_all(tag, R):
    from peakondrake_b200 import experiment
    t0 = time.time()
    experiment(tag + '_c3', 40., 0.2, resolutions=2000, dts=1e-3, sigmas=0.2, seeds=range(R), schemes='hydrodynamic',
               timesteppings='midpoint', ics='periodic_peakon', num_Xis=4, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['q_pde', 'min_du', 'energy', 'u_field'], ndump=100, allow_fail=True, ensemble_stats=True,
               rng='philox', philox_seed=2026)
    t1 = time.time()
    experiment(tag + '_fused', 40., 0.03, resolutions=1000, dts=1e-3, sigmas=0.2, seeds=range(1, R + 1), schemes='hydrodynamic',
               timesteppings='midpoint', ics='proper_peak', num_Xis=4, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['max_du', 'min_du', 'peakon_suite', 'energy'], ndump=1, allow_fail=True, peakon_equations=True)
    return t1 - t0, time.time() - t1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--single', action='store_true')
    ap.add_argument('--dir', default='/tmp/mg')
    ap.add_argument('--R', type=int, default=256)
    args = ap.parse_args()
    import torch
    os.makedirs(args.dir, exist_ok=True)
    os.chdir(args.dir)
    if args.single:
        torch.cuda.set_device(0)
        ta, tb = run_all('one', args.R)
        print(json.dumps({'single': True, 'R': args.R, 'c3_s': ta, 'fused_s': tb}), flush=True)
        return
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ta, tb = run_all('par', args.R)
    dist.barrier()
    if rank == 0:
        from peakondrake_b200 import outputting as out
        res = {'world': world, 'R': args.R, 'c3_s': ta, 'fused_s': tb, 'compared': False}
        if os.path.exists('results/one_c3/data.nc'):
            for code in ('c3', 'fused'):
                a = out.read_data_file('results/one_%s/data.nc' % code)
                b = out.read_data_file('results/par_%s/data.nc' % code)
                assert sorted(a) == sorted(b), code
                for k in a:
                    if k == 'wallclock_time':
                        continue
                    sums = k.startswith('ens_') and not k.startswith('ens_hist') and 'count' not in k and not k.endswith(('_min', '_max'))
                    if sums:
                        assert np.allclose(a[k], b[k], rtol=1e-12, atol=1e-14, equal_nan=True), (code, k)
                    else:
                        assert np.array_equal(a[k], b[k], equal_nan=True), (code, k)
            res['compared'] = True
            res['c3_ens_count'] = [float(x) for x in out.read_data_file('results/par_c3/data.nc')['ens_count']]
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
