"""
Multi-rank host logic on CPU (gloo, world_size 2): the ensemble is partitioned into contiguous blocks
of global realisation ids, stepping needs no communication, and the statistics of a dump are ONE
packed all-reduce.  The packing/reduction is exercised here with the emulated kernel
producing each rank's shard; the NCCL path uses the same `allreduce_packed`.
"""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ch_oracle as O


def _pack(u, nbins, Ld, h):
    """Host restatement of the packed layout of pk_ensemble_stats (see include/peakon_b200.h)."""
    ne = u.shape[1]
    du = (np.roll(u, -1, axis=1) - u) / h
    q = (np.argmax(u, axis=1) + 0.5) * h
    hq = np.histogram(np.clip(q, 0, Ld * (1 - 1e-12)), bins=nbins, range=(0, Ld))[0]
    hd = np.histogram(np.clip(du.min(axis=1), -10, -1e-12), bins=nbins, range=(-10, 0))[0]
    return np.concatenate([u.sum(0), (u * u).sum(0), [u.shape[0]], hq, hd]).astype(np.float64)


def _worker(rank, world, port, R_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_util import emu_step
    from peakondrake_b200.ensemble import partition, allreduce_packed
    ne, Ld, dt, n, nxi = 40, 40.0, 1e-3, 3, 2
    off, cnt = partition(R_total, world, rank)
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(O.Mesh1D(ne, Ld), 'sines', nxi, 0.3)]
    u = np.tile(u0, (cnt, 1)); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    emu_step(ne, Ld, dt, 2, u, u1, dF, dG, tabs, None, n, nxi=nxi, seed=77, rid0=off)   # Philox keyed by GLOBAL id
    packed = torch.from_numpy(_pack(u, 8, Ld, Ld / ne))
    allreduce_packed(packed)
    q.put((rank, off, cnt, u, packed.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_packed_allreduce():
    from peakondrake_b200.ensemble import partition
    # partition covers every id exactly once, for awkward sizes too
    for R, W in ((7, 2), (1024, 8), (5, 8), (65536, 8)):
        ids = []
        for r in range(W):
            o, c = partition(R, W, r)
            ids += list(range(o, o + c))
        assert ids == list(range(R))
    R_total, world = 7, 2
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_util import lib
    lib()          # build the emulation harness HERE if it is stale (minutes of g++), not concurrently in both workers
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, R_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    u_all = np.concatenate([g[3] for g in got], axis=0)
    assert u_all.shape[0] == R_total
    # single-rank run of the whole ensemble gives the same realisations (sharding-invariant noise)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_util import emu_step
    ne, Ld, dt, n, nxi = 40, 40.0, 1e-3, 3, 2
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(O.Mesh1D(ne, Ld), 'sines', nxi, 0.3)]
    u = np.tile(u0, (R_total, 1)); z = [np.zeros_like(u) for _ in range(3)]
    emu_step(ne, Ld, dt, 2, u, z[0], z[1], z[2], tabs, None, n, nxi=nxi, seed=77, rid0=0)
    assert np.array_equal(u, u_all)
    ref = _pack(u, 8, Ld, Ld / ne)
    for g in got:
        assert np.allclose(g[4], ref, rtol=1e-14, atol=1e-14)
        assert g[4][2 * ne] == R_total and g[4][2 * ne + 1:2 * ne + 9].sum() == R_total
