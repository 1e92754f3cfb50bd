"""theboss_b200 -- B200 (sm_100a) implementation of the permanent hot path of Tomev-CTP/theboss.

The package mirrors the reference's module layout for the path it replaces
(``boson_sampling_utilities.permanent_calculators``, ``simulation_strategies``) so that
``from theboss...`` imports can be switched to ``from theboss_b200...`` unchanged.  All arithmetic
runs in hand-written CUDA kernels behind the C ABI of ``include/bossperm.h``; there is no CPU
fallback.
"""
__version__ = "0.1.0"
