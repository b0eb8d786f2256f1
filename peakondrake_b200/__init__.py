"""
peakondrake_b200 -- B200-native ensemble integrator behind peakondrake's Python surface.

    from peakondrake_b200 import experiment, simulation      # drop-in for `from peakondrake import *`
    from peakondrake_b200 import EnsembleIntegrator           # the batched operator API

The numerical hot path is libpeakon_b200.so (hand-written sm_100a CUDA, C ABI in
include/peakon_b200.h); importing this package does not need a GPU, using it does.
"""
from peakondrake_b200.engine import EnsembleIntegrator, validate_scheme          # noqa: F401
from peakondrake_b200.stochastic_functions import build_tables, draw_dW           # noqa: F401
from peakondrake_b200.initial_conditions import build_initial_u                   # noqa: F401
from peakondrake_b200.peakon_equations import PeakonEquations, pure_xi_tables     # noqa: F401
from peakondrake_b200.simulation import (simulation, run_batch, ConvergenceError,  # noqa: F401
                                         PeriodicIntervalMesh, num_steps)
from peakondrake_b200.experiment import experiment                                # noqa: F401
from peakondrake_b200.ensemble import EnsembleStatistics, partition               # noqa: F401
