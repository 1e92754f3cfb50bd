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
        self._norms: Dict[int, np.ndarray] = {}
        self._prefetched: Dict[Tuple[int, ...], np.ndarray] = {}
        self._draw_cache: Dict[Tuple[int, ...], Tuple[List[float], float]] = {}

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
        self._substates, self._weights, self._norms = {}, {}, {}
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
            self._norms[k] = np.array([np.prod([factorial(int(o)) for o in sub]) for sub in subs], dtype=np.float64) * factorial(k)
        self._prefetched = {}

    def _layer_pmf(self, r_sample: Sequence[int]) -> np.ndarray:
        """Un-normalised pmf over the output mode of the next particle (:136-171)."""
        return self._layer_pmfs([tuple(int(v) for v in r_sample)])[0]

    def _layer_pmfs(self, r_samples: Sequence[Tuple[int, ...]]) -> List[np.ndarray]:
        """The layers of several partial outputs from ONE batched K2 launch: layer (r_sample, k = sum + 1) needs the
        m x C(n, k) permanents |perm(U; substate -> r_sample + e_j)|, combined per candidate mode j in the reference's
        order (sequential sum over the substates, :146-155 / :224-247)."""
        U = _native.as_matrix(self._bs_permanent_calculator.matrix)
        m_modes = U.shape[0]
        m = len(self.input_state)
        blocks_S, blocks_T, shapes = [], [], []
        for r_sample in r_samples:
            k = int(sum(r_sample)) + 1
            subs = self._substates[k]
            n_sub = subs.shape[0]
            S = np.zeros((m * n_sub, m_modes), dtype=np.uint8)
            T = np.zeros((m * n_sub, m_modes), dtype=np.uint8)
            S[:, :m] = np.tile(subs, (m, 1))
            T[:, :m] = np.asarray(r_sample, dtype=np.uint8)
            T[np.arange(m * n_sub), np.repeat(np.arange(m), n_sub)] += 1
            blocks_S.append(S); blocks_T.append(T); shapes.append((k, n_sub))
        perms = _native.default_handle(self._device).perm_batched(U, np.concatenate(blocks_S), np.concatenate(blocks_T))
        out, pos = [], 0
        for k, n_sub in shapes:
            p = perms[pos: pos + m * n_sub].reshape(m, n_sub)
            pos += m * n_sub
            mod = np.hypot(p.real, p.imag)               # scalar abs() semantics (numpy.abs on arrays rounds differently)
            terms = mod * mod / self._norms[k] * self._weights[k]
            # numpy.cumsum adds left to right like the reference's Python loop (numpy.sum would add pairwise)
            out.append(np.cumsum(terms, axis=1)[:, -1].copy())
        return out

    # Layers are requested one at a time by the chain rule, but a K2 launch costs the same ~0.1 ms for one layer
    # as for thousands of small permanents, so a miss also computes the descendants of the missing partial output,
    # level by level, while the request stays below this many permanents.  Speculated layers wait in a private
    # cache and enter ``self.pmfs`` (the reference's memo) only when the sampler actually visits them.
    _PREFETCH_BUDGET = 200_000

    def _pmf_for(self, key: Tuple[int, ...]) -> Tuple[np.ndarray, bool]:
        """pmf of the layer of partial output ``key`` and whether it is new to ``self.pmfs``."""
        if key in self.pmfs:
            return self.pmfs[key], False
        if key not in self._prefetched:
            n, m = self.number_of_input_photons, len(self.input_state)
            wanted, frontier, cost = [key], [key], m * self._substates[sum(key) + 1].shape[0]
            while True:
                depth = sum(frontier[0]) + 1                 # particles in the children
                if depth >= n:
                    break
                children = set()
                for y in frontier:
                    for j in range(m):
                        c = y[:j] + (y[j] + 1,) + y[j + 1:]
                        if c not in self.pmfs and c not in self._prefetched:
                            children.add(c)
                level_cost = len(children) * m * self._substates[depth + 1].shape[0]
                if not children or cost + level_cost > self._PREFETCH_BUDGET:
                    break
                frontier = sorted(children)
                wanted += frontier
                cost += level_cost
            for y, pmf in zip(wanted, self._layer_pmfs(wanted)):
                self._prefetched[y] = pmf
        self.pmfs[key] = self._prefetched.pop(key)
        return self.pmfs[key], True

    # -- sampling ----------------------------------------------------------------------------------
    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        self.input_state = input_state
        self.number_of_input_photons = int(sum(input_state))
        self._prepare_substates()
        self.pmfs = {}
        self._draw_cache = {}
        samples = []
        while len(samples) < samples_number:
            self._fill_r_sample()
            samples.append(tuple(self.r_sample))
        return samples

    def _fill_r_sample(self) -> None:
        self.r_sample = [0 for _ in self.input_state]
        n_left = self.number_of_input_photons
        while n_left > 0:
            key = tuple(self.r_sample)
            cached = self._draw_cache.get(key)
            if cached is None:
                # the memoised layer as Python floats plus its left-to-right sum: the same numbers and the same
                # summation order as sum(pmf) / the running sum over the ndarray (:253-262), without re-boxing
                # every element at every visit
                pmf_list = self._pmf_for(key)[0].tolist()
                cached = self._draw_cache[key] = (pmf_list, sum(pmf_list))
            pmf, total = cached
            n_left -= 1
            threshold = np.random.random() * total      # pmfs are not normalised (:253-255)
            running, index = 0, 0
            for p in pmf:
                running += p
                if running > threshold:
                    break
                index += 1
            self.r_sample[index] += 1
