"""Generalized Clifford & Clifford sampler (version A, Oszmaniec & Brod) on the B200.

Drop-in for ``GeneralizedCliffordsSimulationStrategy``
(theboss/simulation_strategies/generalized_cliffords_simulation_strategy.py:24-266).  The chain-rule
bookkeeping stays on the host exactly like the reference (memo of pmfs per partial output :66/:132-133,
inverse-CDF draw on the un-normalised pmf with one ``numpy.random.random()`` per particle :249-266, so
a seeded numpy generator reproduces the reference's decisions).  What moves to the GPU is the hot part:
each new layer needs m x C(n, k) single permanents (:136-171, :224-247) and gets them from ONE batched
launch of kernel K2 instead of m x C(n, k) Python calculator calls.
"""
from itertools import product
from math import factorial
from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.special import binom

from .. import _native
from ..boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_interface import (
    BSPermanentCalculatorInterface,
)
from .simulation_strategy_interface import SimulationStrategyInterface


class GeneralizedCliffordsSimulationStrategy(SimulationStrategyInterface):
    def __init__(self, bs_permanent_calculator: BSPermanentCalculatorInterface) -> None:
        self._bs_permanent_calculator = bs_permanent_calculator
        self._device = getattr(bs_permanent_calculator, "device", 0)
        self.input_state: Sequence[int] = []
        self.number_of_input_photons = 0
        self.r_sample: List[int] = []
        self.pmfs: Dict[Tuple[int, ...], np.ndarray] = {}
        self._substates: Dict[int, np.ndarray] = {}
        self._weights: Dict[int, np.ndarray] = {}

    def set_new_matrix(self, new_matrix) -> None:
        self._bs_permanent_calculator.matrix = new_matrix

    # -- per-simulate preparation ------------------------------------------------------------------
    def _prepare_substates(self) -> None:
        """All sub-states of the input grouped by particle number, first mode varying slowest (the
        enumeration order of :75-121), and their layer weights l!(n-l)!/n! * prod C(s_v, kappa_v)
        normalised per layer (:136-155, :185-205)."""
        s = [int(v) for v in self.input_state]
        n = sum(s)
        groups: Dict[int, List[Tuple[int, ...]]] = {}
        for sub in product(*[range(v + 1) for v in s]):
            groups.setdefault(sum(sub), []).append(sub)
        self._substates, self._weights = {}, {}
        for k, subs in groups.items():
            raw = []
            for sub in subs:
                l = n - k
                w = factorial(l) * factorial(n - l) / factorial(n)
                for v in range(len(s)):
                    w *= binom(s[v], s[v] - sub[v])
                raw.append(w)
            self._substates[k] = np.array(subs, dtype=np.uint8).reshape(len(subs), len(s))
            self._weights[k] = np.array(raw) / sum(raw)

    def _layer_pmf(self, r_sample: Sequence[int]) -> np.ndarray:
        """Un-normalised pmf over the output mode of the next particle (:136-171)."""
        U = _native.as_matrix(self._bs_permanent_calculator.matrix)
        m_modes = U.shape[0]
        m = len(self.input_state)
        k = int(sum(r_sample)) + 1
        subs, weights = self._substates[k], self._weights[k]
        n_sub = subs.shape[0]
        S = np.zeros((m * n_sub, m_modes), dtype=np.uint8)
        T = np.zeros((m * n_sub, m_modes), dtype=np.uint8)
        S[:, :m] = np.tile(subs, (m, 1))
        T[:, :m] = np.asarray(r_sample, dtype=np.uint8)
        T[np.arange(m * n_sub), np.repeat(np.arange(m), n_sub)] += 1
        perms = _native.default_handle(self._device).perm_batched(U, S, T)
        norm = np.array([np.prod([factorial(int(o)) for o in sub]) for sub in subs], dtype=np.float64) * factorial(k)
        pmf = np.zeros(m)
        for j in range(m):
            acc = 0
            for i in range(n_sub):
                acc += abs(perms[j * n_sub + i]) ** 2 / norm[i] * weights[i]
            pmf[j] = acc
        return pmf

    # -- sampling ----------------------------------------------------------------------------------
    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        self.input_state = input_state
        self.number_of_input_photons = int(sum(input_state))
        self._prepare_substates()
        self.pmfs = {}
        samples = []
        while len(samples) < samples_number:
            self._fill_r_sample()
            samples.append(tuple(self.r_sample))
        return samples

    def _fill_r_sample(self) -> None:
        self.r_sample = [0 for _ in self.input_state]
        while self.number_of_input_photons > sum(self.r_sample):
            key = tuple(self.r_sample)
            if key not in self.pmfs:
                self.pmfs[key] = self._layer_pmf(self.r_sample)
            pmf = self.pmfs[key]
            threshold = np.random.random() * sum(pmf)   # pmfs are not normalised (:253-255)
            running, index = 0, 0
            for p in pmf:
                running += p
                if running > threshold:
                    break
                index += 1
            self.r_sample[index] += 1
