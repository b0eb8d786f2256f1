"""
Ensemble statistics across realisations (and across GPUs).

The reference runs one seed per process and merges results off-line (job arrays,
experiments/publication/stochastic/mu_stochastic_sine.py:11-23; the running mean `Eu` of
experiment.py:224-227; the histograms of plotting/plot_stochastic_histograms.py:61-67).  Here the
ensemble lives on the GPUs, so the statistics of a dump are one fused device reduction per GPU plus
ONE packed NCCL all-reduce -- the only collective on this path.
"""
import ctypes
import torch

from peakondrake_b200 import _lib


def partition(R_total, world_size, rank):
    """Contiguous block of global realisation ids owned by `rank`: (offset, count).
    Philox counters are keyed by the GLOBAL id, so results do not depend on world_size."""
    base, extra = divmod(int(R_total), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def allreduce_packed(packed, group=None, op='sum'):
    """The collective of this path: reduce a packed statistics buffer over ranks (NCCL on GPU tensors, gloo on
    CPU tensors in the tests).  op: 'sum' (moments, histograms, counts) or 'min' (extrema; maxima travel negated)."""
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM if op == 'sum' else dist.ReduceOp.MIN, group=group)
    return packed


class Comm(object):
    """
    The ranks of a multi-GPU run (one process per GPU, torch.distributed; NCCL on the GPU box, gloo in the CPU
    tests).  `Comm.world()` is None in a single-process run.  Realisations are never exchanged while stepping; the
    two operations used are the packed statistics all-reduce and, once per dump, the gather of the per-realisation
    diagnostics rows to rank 0, which owns `data.nc` ("per-realisation series by all-gather").
    """

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device if self.backend == 'nccl' else torch.device('cpu')

    @staticmethod
    def world(device=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return Comm(None, device)
        return None

    def shard(self, n_total):
        return partition(n_total, self.size, self.rank)

    def gather_rows(self, rows, n_total):
        """rows: float64 array [n_local, ...] of this rank's contiguous block of realisations.  Returns the
        [n_total, ...] array in global realisation order on rank 0, None elsewhere."""
        import numpy as np
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        tail = rows.shape[1:]
        width = int(np.prod(tail)) if tail else 1
        cap = partition(n_total, self.size, 0)[1]            # rank 0 holds the largest block
        buf = torch.zeros(cap, width, dtype=torch.float64)
        buf[:rows.shape[0]] = torch.from_numpy(rows.reshape(rows.shape[0], width))
        buf = buf.to(self.device)
        parts = [torch.empty_like(buf) for _ in range(self.size)]
        self.dist.all_gather(parts, buf, group=self.group)
        if self.rank != 0:
            return None
        out = [parts[r][:partition(n_total, self.size, r)[1]].cpu().numpy() for r in range(self.size)]
        return np.concatenate(out, axis=0).reshape((n_total,) + tail)

    def allreduce(self, t, op='sum'):
        return allreduce_packed(t, self.group, op)

    def any(self, flag):
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=self.device)
        self.allreduce(t)
        return bool(t.item() > 0)

    def barrier(self):
        self.dist.barrier(group=self.group)


class EnsembleStatistics(object):
    """
    Statistics of one dump over the realisations (of all GPUs):
      packed  = [ sum_r u (ne) | sum_r u^2 (ne) | healthy count | histogram q_pde | histogram min_du ]   (SUM)
      sums    = [ sum_r v_k (NDIAG) | sum_r v_k^2 (NDIAG) | healthy count | failed count ]               (SUM)
      ext     = [ min_r v_k (NDIAG) | -max_r v_k (NDIAG) ]                                               (MIN)
    v_k = the scalar diagnostics of pk_diagnostics.  Local part: fused device reductions (pk_ensemble_stats,
    pk_ensemble_scalars); across GPUs: two all-reduces (the SUM buffers are adjacent in one allocation, so they
    travel together).
    """

    def __init__(self, engine, nbins=64, lo=0.0, hi=None, du_lo=-10.0, du_hi=0.0):
        self.eng = engine
        self.nbins = int(nbins)
        self.q_lo, self.q_hi = float(lo), float(engine.Ld if hi is None else hi)
        self.d_lo, self.d_hi = float(du_lo), float(du_hi)
        nd = _lib.NDIAG
        self.n_packed = 2 * engine.ne + 1 + 2 * self.nbins
        self.all_sums = torch.zeros(self.n_packed + 2 * nd + 2, dtype=torch.float64, device=engine.device)
        self.packed = self.all_sums[:self.n_packed]
        self.sums = self.all_sums[self.n_packed:]
        self.ext = torch.zeros(2 * nd, dtype=torch.float64, device=engine.device)
        self.launches = 0

    def update(self, allreduce=False, comm=None, diagnostics_done=False, u=None, fail_step=None, **diag_kw):
        """Recompute the statistics of the engine's current state (on the current stream).  u / fail_step: a snapshot
        of the state to work on instead (lets the statistics run on a side stream while the next step kernel already
        overwrites the live state; the context's statistics scratch is only touched from that one stream)."""
        eng = self.eng
        if diagnostics_done:
            vals, idx = eng.diag, eng.didx
        else:
            if u is not None:
                diag_kw = dict(diag_kw, u=u, fail_step=fail_step)
            vals, idx, _ = eng.diagnostics(**diag_kw)
            self.launches += 1
        stream = ctypes.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream) if eng.device.type == 'cuda' else None
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        eng.ensemble_stats_into(self, vals, idx, stream, u=u, fail_step=fail_step)
        self.launches += 3      # pk_moments_partial + pk_moments_final + pk_ensemble_scalars
        if comm is not None:
            comm.allreduce(self.all_sums, 'sum')
            comm.allreduce(self.ext, 'min')
        elif allreduce:
            allreduce_packed(self.all_sums)
            allreduce_packed(self.ext, op='min')
        return self.packed

    # views into the packed buffers
    @property
    def count(self):
        return self.packed[2 * self.eng.ne]

    @property
    def fail_count(self):
        return self.sums[2 * _lib.NDIAG + 1]

    @property
    def mean_u(self):
        return self.packed[:self.eng.ne] / self.count

    @property
    def var_u(self):
        ne = self.eng.ne
        m = self.mean_u
        return self.packed[ne:2 * ne] / self.count - m * m

    @property
    def hist_q(self):
        o = 2 * self.eng.ne + 1
        return self.packed[o:o + self.nbins]

    @property
    def hist_min_du(self):
        o = 2 * self.eng.ne + 1 + self.nbins
        return self.packed[o:o + self.nbins]

    def scalar_stats(self, col):
        """(mean, variance, min, max) over the healthy realisations of diagnostics column `col` (host floats)."""
        nd = _lib.NDIAG
        s = self.sums.cpu().numpy()
        e = self.ext.cpu().numpy()
        n = s[2 * nd]
        if n <= 0:
            nan = float('nan')
            return nan, nan, nan, nan
        mean = s[col] / n
        return mean, s[nd + col] / n - mean * mean, e[col], -e[nd + col]
