"""Brod-Oszmaniec approximate sampler for non-uniformly lossy networks (BOBS) on the B200.

Drop-in for ``NonuniformLossesApproximationStrategy``
(theboss/simulation_strategies/nonuniform_losses_approximation_strategy.py:65-347).  For every sample the
reference (i) moves all particles of the approximated modes into one random approximated mode (:283-287),
(ii) thins the input by the uniform loss extracted from the matrix (:298-329), (iii) builds a fresh matrix
``M0 @ random_phases @ QFT`` (:331-347) and (iv) takes ONE sample of the lossy-network GCC-B sampler on its
2m x 2m dilation, all inside a spawn process pool (:254-257).  Here (i)-(ii) and the random phases of (iii) are vectorised NumPy on
the host; the per-sample matrices are built ON THE DEVICE from the dilation template, the phases and the QFT block
(``bp_gccb_simulate_bobs``: 16 k bytes per sample cross the bus instead of 16 (2m)^2), and (iv) is the batched device loop with a
matrix and an input state per sample.  Requests are issued in slices so that host memory stays bounded for any ``samples_number``.

The dilation needs no per-sample SVD: with ``M0 = u diag(sv) v`` the per-sample matrix is
``u diag(sv) (v W)`` for the unitary ``W = phases @ QFT``, so its dilation is
``[[M, u L], [L v W, diag(sv)]]`` with ``L = diag(sqrt(1 - sv^2))``; and because the phases and the QFT only act on
the k approximated modes, only the first k columns of ``M`` and ``L v W`` change from sample to sample.  (NumPy's SVD of ``M`` may differ from
this one by phases on the loss modes, which are traced out; output statistics are identical.)

Deviations: exactly ``samples_number`` samples are returned (the reference rounds up to a multiple of its thread
count, :245-252); random numbers are drawn vectorised from numpy's global generator, so a seeded run is
reproducible but does not replay the reference's per-process streams (those are unseeded spawned processes).
"""
from typing import List, Sequence, Tuple

import numpy as np

from .. import _native
from ..boson_sampling_utilities.boson_sampling_utilities import generate_qft_matrix_for_first_m_modes


class NonuniformLossesApproximationStrategy:
    def __init__(self, bs_permanent_calculator, approximated_modes_number: int, threads_number: int = -1) -> None:
        total_modes = len(bs_permanent_calculator.matrix)
        self._approximated_modes_number = int(min(max(approximated_modes_number, 0), total_modes))
        self._permanent_calculator = bs_permanent_calculator
        self._threads_number = threads_number          # kept for signature parity; the GPU batches instead
        self._device = getattr(bs_permanent_calculator, "device", 0)
        self._uniform_losses = 0.0
        self._initial_matrix = None
        self._extract_losses_from_the_interferometer(bs_permanent_calculator.matrix)

    def _extract_losses_from_the_interferometer(self, interferometer_matrix) -> None:
        """Largest uniform loss that can be pulled out of the matrix (:120-144)."""
        u, sv, v = np.linalg.svd(np.asarray(interferometer_matrix, dtype=np.complex128))
        eta = sv ** 2
        self._uniform_losses = float(np.min(1 - eta))
        sv = np.sqrt(eta / (1 - self._uniform_losses))
        self._svd = (u, np.clip(sv, 0.0, 1.0), v)
        self._initial_matrix = u @ np.diag(sv) @ v

    #: samples per device request (host arrays of one slice: states and phases)
    _SLICE_SAMPLES = 65536

    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        if samples_number < 1:
            return []
        m, k, total = len(input_state), self._approximated_modes_number, int(samples_number)
        base = np.array(input_state, dtype=np.int64)
        approx_particles = int(base[:k].sum())
        base[:k] = 0
        # Only the first k columns of M_s = M0 @ diag(phases_s) @ QFT and of its loss block L v W_s depend on the sample
        # (the QFT and the phases act on the approximated modes): stack B = [[M0], [L v]] (2m x m) once, then per sample
        # B[:, :k] diag(phases_s) QFT_k -- 2 m k^2 multiplications instead of two m^3 products.
        u, sv, v = self._svd
        loss = np.sqrt(np.clip(1.0 - sv ** 2, 0.0, None))
        stacked = np.concatenate([(u * sv[None, :]) @ v, loss[:, None] * v], axis=0)
        template = np.zeros((2 * m, 2 * m), dtype=np.complex128)
        template[:, k:m] = stacked[:, k:]
        template[:m, m:] = u * loss[None, :]
        template[m:, m:] = np.diag(sv)
        qft_k = generate_qft_matrix_for_first_m_modes(k, m)[:k, :k]
        seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        handle = _native.default_handle(self._device)
        template[:, :k] = stacked[:, :k]                 # the columns that meet the phases and the QFT on the device
        step = max(1, min(total, self._SLICE_SAMPLES))
        out = np.zeros((total, m), dtype=np.int32)
        for lo in range(0, total, step):
            S = min(step, total - lo)
            states = np.repeat(base[None, :], S, axis=0)
            if k > 0:
                states[np.arange(S), np.random.randint(0, k, S)] = approx_particles
            if not np.isclose(self._uniform_losses, 0):
                states = np.random.binomial(states, 1.0 - self._uniform_losses)     # each particle survives independently
            phases = np.exp(2j * np.pi * np.random.rand(S, k))
            big_states = np.zeros((S, 2 * m), dtype=np.int32)
            big_states[:, :m] = states
            # per-sample matrix = template @ diag(phases_s, 1 ...) @ QFT_k, built on the device; one Philox stream for the whole
            # request: slices continue the sample counter
            out[lo:lo + S] = handle.gccb_simulate_bobs(template, qft_k, phases, None, big_states, seed=seed, first_sample=lo)[:, :m]
        return [tuple(row) for row in out.tolist()]
