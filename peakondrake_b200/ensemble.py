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


def allreduce_packed(packed, group=None):
    """The one collective of this path: sum the packed statistics buffer over ranks (NCCL on GPU
    tensors, gloo on CPU tensors in the tests)."""
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


class EnsembleStatistics(object):
    def __init__(self, engine, nbins=64, lo=0.0, hi=None, du_lo=-10.0, du_hi=0.0):
        self.eng = engine
        self.nbins = int(nbins)
        self.q_lo, self.q_hi = float(lo), float(engine.Ld if hi is None else hi)
        self.d_lo, self.d_hi = float(du_lo), float(du_hi)
        self.packed = torch.zeros(2 * engine.ne + 1 + 2 * self.nbins, dtype=torch.float64, device=engine.device)
        self.launches = 0

    def update(self, allreduce=False):
        """Recompute the statistics of the engine's current state (on the current stream)."""
        eng = self.eng
        vals, idx, _ = eng.diagnostics()
        stream = ctypes.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = eng.lib.pk_ensemble_stats(eng.ctx, p(eng.u), p(eng.fail_step), p(vals), p(idx), self.nbins,
                                       self.q_lo, self.q_hi, self.d_lo, self.d_hi, p(self.packed), stream)
        _lib.check(eng.lib, eng.ctx, rc)
        self.launches += 3      # pk_diag_kernel + pk_moments_partial + pk_moments_final
        if allreduce:
            allreduce_packed(self.packed)
        return self.packed

    # views into the packed buffer
    @property
    def count(self):
        return self.packed[2 * self.eng.ne]

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
