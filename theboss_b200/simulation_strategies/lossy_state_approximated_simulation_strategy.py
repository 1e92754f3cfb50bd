"""Uniform-loss Brod-Oszmaniec approximate sampler (lossy-state approximation) on the B200.

Drop-in for ``LossyStateApproximationSimulationStrategy``
(theboss/simulation_strategies/lossy_state_approximated_simulation_strategy.py:36-362).  Per sample the
reference draws a lossy input state -- the first ``hierarchy_level`` modes keep each particle with probability
eta (:96-138, :312-327), the remaining (approximated) modes are replaced by l ~ Binomial(n_approx, eta)
particles in the first approximated mode (:170-203) -- and a matrix ``U[:, random permutation] @ random_phases @
QFT`` acting on the approximated modes (:329-362), then takes ONE lossless GCC-B sample in a spawn process pool
(:287-310).  Here the per-sample inputs, phases and column permutations are drawn vectorised on the host; the matrices are built on the
device (``bp_gccb_simulate_bobs``) and sampled by the batched device loop, one request per slice of samples.

Deviation: the not-approximated part is thinned particle by particle (Binomial(s_i, eta) per mode), which equals
the reference's weights for collision-free inputs and stays normalised for bunched ones (the reference's weights
do not sum to one there and numpy.random.choice raises).
"""
from typing import List, Sequence, Tuple

import numpy as np

from .. import _native
from ..boson_sampling_utilities.boson_sampling_utilities import generate_qft_matrix_for_first_m_modes
from .simulation_strategy_interface import SimulationStrategyInterface


class LossyStateApproximationSimulationStrategy(SimulationStrategyInterface):
    def __init__(self, bs_permanent_calculator, uniform_transmissivity: float, hierarchy_level: int,
                 threads_number: int = -1) -> None:
        self._permanent_calculator = bs_permanent_calculator      # must hold a UNITARY (losses live in the state)
        self._uniform_transmissivity = float(uniform_transmissivity)
        self._hierarchy_level = int(hierarchy_level)
        self._threads_number = threads_number                     # signature parity; the GPU batches instead
        self._device = getattr(bs_permanent_calculator, "device", 0)

    #: samples per device request (host arrays of one slice: states, phases, permutations)
    _SLICE_SAMPLES = 65536

    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        if samples_number < 1:
            return []
        U = _native.as_matrix(self._permanent_calculator.matrix)
        m, total, eta = U.shape[0], int(samples_number), self._uniform_transmissivity
        hl = min(max(self._hierarchy_level, 0), m)
        state = np.array(input_state, dtype=np.int64)
        # NOTE: like the reference (:345-362, :39-58) the random phases and the QFT act on the FIRST a = m - hl modes
        a = m - hl
        qft_a = generate_qft_matrix_for_first_m_modes(a, m)[:a, :a]
        seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        handle = _native.default_handle(self._device)
        step = max(1, min(total, self._SLICE_SAMPLES))
        out = np.zeros((total, m), dtype=np.int32)
        for lo in range(0, total, step):
            S = min(step, total - lo)
            states = np.zeros((S, m), dtype=np.int32)
            states[:, :hl] = np.random.binomial(np.repeat(state[None, :hl], S, axis=0), eta)
            if hl < m:
                states[:, hl] = np.random.binomial(int(state[hl:].sum()), eta, S)
            phases = np.exp(2j * np.pi * np.random.rand(S, a))
            perms = np.argsort(np.random.rand(S, m), axis=1).astype(np.int32)     # one column permutation per sample
            # Us[s] = U[:, perms[s]] @ diag(phases_s, 1 ...) @ QFT_a, built on the device
            out[lo:lo + S] = handle.gccb_simulate_bobs(U, qft_a, phases, perms, states, seed=seed, first_sample=lo)
        return [tuple(row) for row in out.tolist()]
