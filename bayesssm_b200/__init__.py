"""bayesssm_b200: B200-native particle-filter / PMMH engine behind the bayesSSM API.

Host-side mirror of the reference's exported interface (NAMESPACE:3-11): bootstrap_filter,
auxiliary_filter, resample_move_filter, pmmh, default_tune_control, ess, rhat -- over the C ABI of
libbayesssm_b200.so (include/bayesssm_b200.h).  CUDA only; nothing here computes on the CPU.
"""
from . import _native, distributed, models, sharding
from .diagnostics import ess, rhat
from .filters import (auxiliary_filter, batched_bootstrap_filter, bootstrap_filter, particle_filter,
                      resample_move_filter)
from .pmmh import default_tune_control, pmmh, priors
from .resampling import (resample_multinomial, resample_multinomial_cpp, resample_stratified,
                         resample_stratified_cpp, resample_systematic, resample_systematic_cpp)

__all__ = ["bootstrap_filter", "auxiliary_filter", "resample_move_filter", "particle_filter", "pmmh",
           "default_tune_control", "priors", "ess", "rhat", "models", "batched_bootstrap_filter",
           "resample_multinomial_cpp", "resample_stratified_cpp", "resample_systematic_cpp",
           "resample_multinomial", "resample_stratified", "resample_systematic", "sharding", "distributed"]
__version__ = "0.1.0"
