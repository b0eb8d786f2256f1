#!/usr/bin/env python
"""
bench.py -- headline benchmark: realisation*steps/s of the batched stochastic Camassa-Holm
timestep (BASELINE.json metric), workload = BASELINE config 2 per GPU:
  hydrodynamic/midpoint, ne=1000, Ld=40, dt=1e-3, 4 'sines' noise modes sigma=0.2, ic 'proper_peak',
  1024 realisations per GPU (weak scaling), Brownian increments drawn in-kernel (Philox).

A bench "step" is ONE pass of the hot path over the batch: one kernel launch that advances every
realisation of the shard by STEPS_PER_LAUNCH timesteps (the reference's dump cadence, ndump=50);
its persistent CTAs pull (realisation, block of timesteps) items from a work queue.

  python bench.py [--gpus N] [--steps K] [--warmup W]            own arm (GPU)
  python bench.py --impl reference ...                            CPU arm: the oracle port of the
                                                                  reference's Firedrake path on all host cores
Under torchrun (N>1) every rank drives one GPU; NCCL is used for the barrier, the max-over-ranks
timing and the per-launch ensemble statistics all-reduce (the only exchange this path has).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NE, LD, DT, NXI, SIGMA, IC = 1000, 40.0, 1e-3, 4, 0.2, 'proper_peak'
R_PER_GPU = 1024
STEPS_PER_LAUNCH = 50
B_ALG = 16 * NE + 8 * NXI      # algorithmic bytes per realisation-step
METRIC = 'realisation_steps_per_sec'
UNIT = 'realisation*steps/s'
WORKLOAD = ('C2: stochastic CH hydrodynamic/midpoint, ne=1000, Ld=40, dt=1e-3, 4 sines sigma=0.2, '
            'ic proper_peak, %d realisations/GPU, %d timesteps per launch' % (R_PER_GPU, STEPS_PER_LAUNCH))


def shared_config(world):
    """`config` is the SAME object in both arms (the driver compares them); arm-specific facts go to `details`."""
    return {'workload': WORKLOAD, 'realisations_total': world * R_PER_GPU, 'ne': NE, 'dt': DT,
            'timesteps_per_launch': STEPS_PER_LAUNCH, 'noise_modes': NXI, 'sigma': SIGMA, 'ic': IC,
            'l2': 'GPU arm: flushed between timed iterations (256 MB memset); CPU arm: not applicable'}


def measured_traffic():
    """dram bytes per launch of the headline kernel, from the file profiles/ncu_traffic.py writes out of an
    `ncu --set full` capture of this very command (never a literal here); None if no capture has been summarised."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            t = json.load(f)
        return float(t['pk_step_kernel']['dram_bytes_per_launch']), t['pk_step_kernel'].get('source')
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [s for (ts, s) in self.samples
                  if (self.t0 is None or ts >= self.t0) and (self.t1 is None or ts <= self.t1 + 0.05)]
        if not inside:      # region shorter than the sampling period: fall back to every sample taken under load
            inside = [s for (_, s) in self.samples]
        for s in inside:
            f = [x.strip() for x in s.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port, one worker per core, whole realisations each
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init():
    """Per-worker set-up (untimed): mesh, IC, noise tables, one warm-up step."""
    from oracle import ch_oracle as O
    mesh = O.Mesh1D(NE, LD)
    _W['O'] = O
    _W['tabs'] = O.noise_tables(mesh, 'sines', NXI, SIGMA)
    _W['orc'] = O.CHOracle(NE, LD, DT)
    _W['u'] = O.initial_u(mesh, IC)
    _W['seed'] = os.getpid()
    _cpu_worker(1)


def _cpu_worker(nsteps):
    """Advance this worker's realisation by nsteps steps of the bench workload; returns its busy time."""
    O, tabs, orc = _W['O'], _W['tabs'], _W['orc']
    dW = O.reference_dW(_W['seed'], nsteps, NXI, DT)
    _W['seed'] += 7919
    u = _W['u']
    t0 = time.perf_counter()
    for n in range(nsteps):
        u = orc.solve(u, dW[n] @ tabs[0], dW[n] @ tabs[1], dW[n] @ tabs[2])
    _W['u'] = u
    return time.perf_counter() - t0


class CpuArm(object):
    """The oracle port of the reference's Firedrake path on all host cores: one process per core (the
    reference's own job-array parallelism), each integrating whole realisations of the bench workload."""

    def __init__(self, workers=None):
        import multiprocessing as mp
        self.workers = workers or (os.cpu_count() or 1)
        os.environ.setdefault('OMP_NUM_THREADS', '1')
        self.pool = mp.get_context('fork').Pool(self.workers, initializer=_cpu_init)
        self.pool.map(_cpu_worker, [1] * self.workers)      # every worker is up and warm

    def step(self, nsteps):
        """One bench step = nsteps timesteps of `workers` realisations; returns (rate, wall seconds)."""
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [nsteps] * self.workers, chunksize=1)
        wall = time.perf_counter() - t0
        return self.workers * nsteps / wall, wall

    def sample(self, nsteps, wall):
        return ('%d realisations (one per core) x %d steps of the bench workload, oracle (NumPy/SciPy '
                'monolithic sparse-LU Newton), wall %.1f s' % (self.workers, nsteps, wall))

    def close(self):
        self.pool.close(); self.pool.join()


def cpu_rate(target_seconds=15.0):
    """cpu_baseline of the own arm: a bounded sample (about target_seconds of wall on every core)."""
    arm = CpuArm()
    r0, _ = arm.step(20)
    nsteps = max(20, int(target_seconds * r0 / arm.workers))
    rate, wall = arm.step(nsteps)
    sample = arm.sample(nsteps, wall)
    cores = arm.workers
    arm.close()
    return rate, cores, sample


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm()
    r0, _ = arm.step(20)
    # each bench step is a bounded sample sized so that warm-up + K steps end within ~3 minutes
    budget = 170.0 / max(1, args.steps + args.warmup)
    nsteps = max(4, int(budget * r0 / arm.workers))
    for _ in range(args.warmup):
        arm.step(nsteps)
    rates, walls = [], []
    for _ in range(args.steps):
        r, w = arm.step(nsteps)
        rates.append(r); walls.append(w)
    units = arm.workers * nsteps
    rate = units * len(rates) / sum(walls)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': len(rates), 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(walls) / len(walls),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': shared_config(args.gpus),
        'details': {'note': 'CPU arm: Firedrake is not installable offline; this is the oracle port of the same weak form '
                            '(BASELINE.md section 4), each step a bounded sample extrapolated linearly (realisations are independent)'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': arm.workers, 'kind': 'port',
                         'sample': arm.sample(nsteps, sum(walls) / len(walls)) + ' per bench step'},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    arm.close()
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def extra_configs(rank, world, local, peak):
    """The other BASELINE configurations and the upwind scheme, measured in the same process after the headline
    (VERDICT r1 #4): per-GPU shard at full size, device-timed (CUDA events around the whole leg, max over ranks),
    step-kernel time from the library's own event pairs."""
    import torch
    import torch.distributed as dist
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    from peakondrake_b200 import ensemble as ens
    dev = torch.device('cuda', local)
    comm = ens.Comm.world(dev)
    out = {}

    def leg(name, ne, R, dt, ic, fam, nxi, sigma, launches, steps_per_launch, scheme='hydrodynamic', peak_width=1.0,
            nupd=1, per_launch=None, b_alg=None, note='', dump_every=0):
        eng = EnsembleIntegrator(ne, LD, dt, R, scheme=scheme, timestepping='ssprk3' if scheme == 'upwind' else 'midpoint',
                                 num_Xis=nxi, Xi_family=fam, sigma=sigma, nXi_updates=nupd, seed=2026,
                                 rid_offset=rank * R, device=local)
        eng.set_state(build_initial_u(ne, LD, ic, peak_width=peak_width))
        ctx = dict(eng=eng, comm=comm)
        if per_launch is not None:
            per_launch(ctx, True)
        def advance():
            if dump_every:      # per-dump state snapshots inside the launch + ONE batched diagnostics / peakon-suite call,
                                # D2H of the [dumps, R, 14+4+2+3] tables per launch (what run_batch does)
                _, dg, di, sv, si, _, _ = eng.step_snap(steps_per_launch, dump_every, suite=True)
                ctx['host'] = [t.to('cpu', non_blocking=True) for t in (dg, di, sv, si)]
            else:
                eng.step(steps_per_launch)
        advance()                                       # warm-up launch
        if per_launch is not None:
            per_launch(ctx, False)
        torch.cuda.synchronize()
        eng.kernel_ms_since()
        n0 = eng.launch_count
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(launches):
            advance()
            if per_launch is not None:
                per_launch(ctx, False)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        kms, kn = eng.kernel_ms_since()
        units = world * R * launches * steps_per_launch
        rate = units / (ms.item() * 1e-3)
        b = b_alg if b_alg is not None else (16 * ne + 8 * nxi)
        failed = int((eng.fail_step >= 0).sum().item())
        out[name] = {'workload': note, 'ne': ne, 'realisations_per_gpu': R, 'timesteps': launches * steps_per_launch,
                     'value': rate, 'unit': UNIT, 'ms_total': ms.item(),
                     'step_kernel_ms_per_launch': (kms / kn) if kn else None,
                     'step_kernel_rate': (R * steps_per_launch * kn / (kms * 1e-3)) if kn and kms > 0 else None,
                     'algorithmic_bytes_per_realisation_step': b, 'hbm_frac': rate / world * b / (peak * 1e9),
                     'gpu_launches': eng.launch_count - n0, 'failed_realisations': failed, 'layout': eng.layout}
        eng.close()
        del eng
        torch.cuda.empty_cache()

    # C3: 8192 realisations per GPU, ne=2000, 4 sines, moments + histograms all-reduced every 100 steps
    def c3_stats(ctx, first):
        if first:
            ctx['stats'] = ens.EnsembleStatistics(ctx['eng'], nbins=64, lo=0.0, hi=LD)
            return
        ctx['stats'].update(comm=ctx['comm'])
    leg('C3', 2000, 8192, 1e-3, 'periodic_peakon', 'sines', 4, 0.2, 3, 100, per_launch=c3_stats,
        note='65536 realisations over 8 GPUs = 8192/GPU, ne=2000, periodic_peakon, 4 sines sigma=0.2, dt=1e-3; per-x moments, '
             'scalar moments/extrema and the q_pde / min_du histograms all-reduced every 100 steps')

    # C4: 4096 realisations, ne=16384, steep front, diagnostics every step
    leg('C4', 16384, 4096, 5e-4, 'new_peak', 'constant', 1, 0.2, 3, 10, peak_width=1 / 6, dump_every=1,
        b_alg=16 * 16384 + 8 + 8 * 18,
        note='4096 realisations, ne=16384, new_peak width 1/6, constant Xi sigma=0.2, dt=5e-4, min_du / max_du / '
             'max_jump_local / peakon_suite EVERY step (pk_step_snap: one step launch per 10 steps with a state snapshot per '
             'step, one batched pk_diagnostics_rows + pk_peakon_suite_rows over the 10 x R snapshots, tables copied to the host)')
    # C5: strong-convergence sweep shard, ne=20000, sigma=1 (the dt=1e-3 member, nXi_updates=1000)
    leg('C5', 20000, 8192, 1e-3, 'periodic_peakon', 'constant', 1, 1.0, 2, 20, nupd=1000,
        note='8192 realisations, ne=20000, periodic_peakon, constant Xi sigma=1, dt=1e-3 with nXi_updates=1000 (the '
             'dt-sweep member in the middle of peakon_strong_dt.py), in-kernel Philox sub-updates')
    # upwind / ssprk3 (the genuinely HBM-shaped scheme), ne=1000, 4 sines
    leg('upwind_ssprk3', 1000, 4096, 1e-3, 'proper_peak', 'sines', 4, 0.02, 20, 50, scheme='upwind', b_alg=32 * 1000 + 8 * 4,
        note='upwind/ssprk3 (equations.py:28-61), ne=1000, 4 sines sigma=0.02, 4096 realisations, 50 steps per launch')
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u
    from peakondrake_b200 import ensemble as ens

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    comm = ens.Comm.world(dev)

    R = R_PER_GPU
    u0 = build_initial_u(NE, LD, IC)
    eng = EnsembleIntegrator(NE, LD, DT, R, num_Xis=NXI, Xi_family='sines', sigma=SIGMA, seed=2026,
                             rid_offset=rank * R, device=local)
    eng.set_state(u0)
    stats = ens.EnsembleStatistics(eng, nbins=64, lo=0.0, hi=LD)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > L2 (126 MB)
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    # The ensemble statistics of a launch (local fused reductions + the packed all-reduces) run on a side stream,
    # overlapping the NEXT step kernel.  They work on a SNAPSHOT of (u, fail_step) taken on the main stream, and the
    # next snapshot waits for the side stream: nothing the statistics read is being overwritten.
    snap_u = torch.empty_like(eng.u)
    snap_fs = torch.empty_like(eng.fail_step)
    snap_ready = torch.cuda.Event()
    side_done = torch.cuda.Event()
    side_done.record(side)

    def one_step():
        flush.zero_()                         # L2 flush between timed iterations
        eng.step(STEPS_PER_LAUNCH)            # (CUDA events around the kernel are recorded on the launching stream)
        main.wait_event(side_done)            # the previous statistics are done with the snapshot
        snap_u.copy_(eng.u); snap_fs.copy_(eng.fail_step)
        snap_ready.record(main)
        side.wait_event(snap_ready)
        with torch.cuda.stream(side):
            stats.update(comm=comm, u=snap_u, fail_step=snap_fs)
            side_done.record(side)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    eng.kernel_ms_since()                     # drop the warm-up launches from the kernel timer
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    launches0 = eng.launch_count
    kms, klaunches = 0.0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        one_step()
        if (i + 1) % 48 == 0 or i + 1 == args.steps:      # the event ring holds 64 launches
            ms_, n_ = eng.kernel_ms_since()
            kms += ms_; klaunches += n_
    main.wait_stream(side)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_end()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    gpu_launches = eng.launch_count - launches0          # counted by the library at every kernel launch
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the host-buffer C-ABI call, pinned host arrays, H2D + kernel + D2H every step ----
    pin = lambda shape, dt_: torch.zeros(shape, dtype=dt_).pin_memory().numpy()
    hu = pin((R, NE), torch.float64); hu[:] = u0
    hfs = pin((R,), torch.int32); hfs[:] = -1
    hil, hit = pin((R,), torch.int32), pin((R,), torch.int32)
    # u in / u out per step; the Newton unknown w1 = (u1, dF, dG) is solver-internal state in the reference
    # (equations.py:71,89) and stays resident in the context (u1 = dF = dG = None)
    for _ in range(min(2, args.warmup)):
        eng.step_host(STEPS_PER_LAUNCH, hu, None, None, None, None, hfs, hil, hit)
    if world > 1:
        dist.barrier()
    te0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 20))
    for _ in range(e2e_steps):
        eng.step_host(STEPS_PER_LAUNCH, hu, None, None, None, None, hfs, hil, hit)
        energy_check = float(hu[0].max())       # the host reads the step's result
    e2e_wall = time.perf_counter() - te0
    h2d = R * NE * 8 + 2 * R * 4
    d2h = R * NE * 8 + 3 * R * 4

    tm = torch.tensor([dev_ms, e2e_wall * 1e3, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = [float(x) for x in tm.cpu()]
    units_per_step = world * R * STEPS_PER_LAUNCH
    value = units_per_step * args.steps / (dev_ms * 1e-3)
    e2e_value = units_per_step * e2e_steps / (e2e_ms * 1e-3)
    failed = int((eng.fail_step >= 0).sum().item())
    its_per_step = float(eng.its_total.double().mean().item()) / ((args.warmup + args.steps) * STEPS_PER_LAUNCH)
    layout = eng.layout
    eng.close()
    del eng, flush, snap_u
    torch.cuda.empty_cache()

    peak, which = measured_peak()
    configs = None
    if args.configs:
        try:
            configs = extra_configs(rank, world, local, peak)
        except Exception as e:          # the headline line must survive a failing extra leg
            configs = {'error': repr(e)}

    if rank == 0:
        k_ms = kms / max(1, klaunches)
        achieved = R * STEPS_PER_LAUNCH * B_ALG / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic()
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': shared_config(world),
            'details': {'noise': 'Philox4x32-10 in-kernel', 'layout': layout, 'failed_realisations': failed,
                       'mean_newton_its_per_step': its_per_step,
                       'statistics': 'per launch: fused diagnostics + per-x moments + scalar moments/extrema + histograms on a '
                                     'snapshot of the state, side stream, two packed all-reduces (sum, min)'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps,
                    'api': 'pk_step_host: u in/out through pinned host buffers (chunked H2D/kernel/D2H pipeline, counters in one '
                           'staging block per chunk); the Newton warm start (u1, dF, dG) is solver-internal state as in the '
                           'reference and stays in the context'},
            'gpu_launches': gpu_launches,
            'gpu_launches_note': 'pk_launch_count: per step pk_step_kernel, pk_diag_kernel, pk_moments_partial, pk_moments_final, '
                                 'pk_ensemble_scalars_kernel',
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': which,
                         'traffic_note': 'dram__bytes_read+write per launch, read from profiles/traffic.json (%s); the state is '
                                         'loaded once per launch and stepped on-chip for 50 timesteps' % traffic_src,
                         'kernel': 'pk_step_kernel', 'kernel_ms': k_ms,
                         'algorithmic_bytes_per_realisation_step': B_ALG,
                         'note': 'not HBM-bound by construction (SURVEY.md 8d): ~600 fp64 instructions per node-step against '
                                 '16 algorithmic bytes; the binding resource is the fp64 pipe (DESIGN.md section 4)'},
            'wall_ms_per_step': wall_ms / args.steps,
        }
        if configs is not None:
            line['configs'] = configs
        if args.cpu_baseline:
            rate, cores, sample = cpu_rate(15.0)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-cpu-baseline', dest='cpu_baseline', action='store_false')
    ap.add_argument('--no-configs', dest='configs', action='store_false',
                    help='skip the extra legs (C3, C4, C5, upwind) measured after the headline')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'own' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        if int(os.environ.get('WORLD_SIZE', '1')) > 1:
            args.cpu_baseline = False     # cpu_baseline is reported at N=1 only
        run_gpu(args)


if __name__ == '__main__':
    main()
