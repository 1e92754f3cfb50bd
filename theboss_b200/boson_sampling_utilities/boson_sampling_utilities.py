"""Host-side helpers of the permanent path (NumPy), mirroring the functions of the reference's
``theboss/boson_sampling_utilities/boson_sampling_utilities.py`` that the hot path and its callers touch:

* ``mode_occupation_to_mode_assignment``                 (:61-78)
* ``prepare_interferometer_matrix_in_expanded_space``    (:287-342, helper :263-284)
* ``EffectiveScatteringMatrixCalculator``                (:545-626)
* ``generate_possible_states`` / ``generate_lossy_n_particle_input_states`` (:81-203), needed by the exact
  distribution calculators that sit on top of the batched permanent kernel
* the state-space bookkeeping helpers its callers and tests import from the same module (``bosonic_space_dimension``,
  ``get_modes_transmissivity_values_from_matrix``, state types and their counts, :206-261, :345-502)

The expansion of rows/columns by occupation is done on the device inside the kernels
(theboss_b200/csrc/util_kernels.cu, guan_kernel.cu); the class below exists for API compatibility and
for callers that want the explicit matrix.
"""
import functools
import itertools
from collections import Counter
from math import comb, factorial
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np


def mode_occupation_to_mode_assignment(mode_occupation: Sequence[int]) -> Tuple[int, ...]:
    """2nd-quantisation occupation -> 1st-quantisation list of modes, e.g. [2,0,1] -> (0,0,2)."""
    occ = np.asarray(mode_occupation).astype(np.int64)
    return tuple(int(v) for v in np.repeat(np.arange(len(occ)), occ))


def mode_assignment_to_mode_occupation(modes_assignment: Sequence[int], observed_modes_number: int = 0) -> Tuple[int, ...]:
    """1st-quantisation list of modes -> occupation vector, e.g. (0,0,2), 4 -> (2,0,1,0)."""
    modes = [int(v) for v in modes_assignment]
    size = max(observed_modes_number, (max(modes) + 1) if modes else 0)
    occ = [0] * size
    for v in modes:
        occ[v] += 1
    return tuple(occ)


def _n_particle_states(n: int, modes_number: int) -> List[Tuple[int, ...]]:
    """All occupations of ``modes_number`` modes by exactly n particles, in descending lexicographic order
    (the order of the reference's generate_possible_states, :117-155)."""
    if modes_number == 1:
        return [(n,)]
    out = []
    for first in range(n, -1, -1):
        out.extend((first,) + rest for rest in _n_particle_states(n - first, modes_number - 1))
    return out


def generate_possible_states(particles_number: int, modes_number: int, losses: bool = False) -> List[Tuple[int, ...]]:
    """All m-mode states with exactly n particles, or (``losses=True``) with 0..n particles ordered by
    particle number; every block in descending lexicographic order."""
    if particles_number < 0 or modes_number < 1:
        return []
    if particles_number == 0:
        return [tuple([0] * modes_number)]
    first = 0 if losses else particles_number
    states: List[Tuple[int, ...]] = []
    for n in range(first, particles_number + 1):
        states.extend(_n_particle_states(n, modes_number))
    return states


def generate_lossy_n_particle_input_states(initial_state: Sequence[int], number_of_particles_left: int) -> List[Tuple[int, ...]]:
    """Distinct occupations obtained by keeping ``number_of_particles_left`` of the input particles, in
    order of first appearance over itertools.combinations of the particle list (reference :158-203)."""
    if sum(initial_state) == 0:
        return [tuple(initial_state)]
    m = len(initial_state)
    particles = mode_occupation_to_mode_assignment(initial_state)
    seen, out = set(), []
    for combo in itertools.combinations(particles, number_of_particles_left):
        occ = mode_assignment_to_mode_occupation(combo, m)
        if occ not in seen:
            seen.add(occ)
            out.append(occ)
    return out


def bosonic_space_dimension(particles_number: int, modes_number: int, losses: bool = False) -> int:
    """Number of m-mode Fock states with exactly n particles (stars and bars), or with at most n when ``losses``
    (reference :206-237)."""
    fewest = 0 if losses else particles_number
    return sum(comb(n + modes_number - 1, n) for n in range(fewest, particles_number + 1))


def get_modes_transmissivity_values_from_matrix(lossy_interferometer_matrix) -> np.ndarray:
    """Squared singular values of a (lossy) interferometer, smallest first; like the reference (:240-261) the order
    is the SVD's, not the modes'."""
    sv = np.linalg.svd(np.asarray(lossy_interferometer_matrix, dtype=np.complex128), compute_uv=False)
    return np.square(sv[::-1])


def _ascending_partitions(n: int, smallest: int = 1) -> Iterator[Tuple[int, ...]]:
    """Integer partitions of n with non-decreasing parts >= ``smallest``: the one-part partition first, then by first part."""
    yield (n,)
    for first in range(smallest, n // 2 + 1):
        for rest in _ascending_partitions(n - first, first):
            yield (first,) + rest


def generate_state_types(modes_number: int, particles_number: int, losses: bool = False) -> List[Tuple[int, ...]]:
    """State types (occupations up to mode permutations, written in non-increasing order and padded with zeros) of n
    particles in m modes; with ``losses`` followed by those of 0, 1, ..., n - 1 particles (reference :345-400)."""
    counts = [particles_number] + (list(range(particles_number)) if losses else [])
    types = []
    for n in counts:
        for parts in _ascending_partitions(n):
            if len(parts) <= modes_number:
                types.append(tuple(sorted(parts, reverse=True)) + (0,) * (modes_number - len(parts)))
    return types


@functools.lru_cache(maxsize=None)
def compute_number_of_k_element_integer_partitions_of_n(k: int, n: int) -> int:
    """p_k(n) by the recurrence p_k(n) = p_k(n - k) + p_{k-1}(n - 1), with the reference's conventions at the edges
    (:478-502): one 1-element partition for every n (also n = 0), none when k > n, n = 0 or k < 1."""
    if k == 1:
        return 1
    if k > n or n == 0 or k < 1:
        return 0
    return (compute_number_of_k_element_integer_partitions_of_n(k, n - k)
            + compute_number_of_k_element_integer_partitions_of_n(k - 1, n - 1))


def compute_number_of_state_types(modes_number: int, particles_number: int, losses: bool = False) -> int:
    """Number of partitions of n into at most m parts; with ``losses`` summed over 0 .. n particles (reference :403-441)."""
    counts = range(particles_number + 1) if losses else [particles_number]
    return sum(compute_number_of_k_element_integer_partitions_of_n(k, n) for n in counts for k in range(1, modes_number + 1))


def compute_number_of_states_of_given_type(state_type: Sequence[int]) -> int:
    """Distinct mode permutations of a state type: m! / prod(multiplicity of each occupation value)! (reference :444-475)."""
    number = factorial(len(state_type))
    for multiplicity in Counter(state_type).values():
        number //= factorial(multiplicity)
    return number


def prepare_interferometer_matrix_in_expanded_space(interferometer_matrix) -> np.ndarray:
    """m x m (possibly lossy) matrix -> 2m x 2m dilation (an isometry on the m physical input modes, which
    is all the samplers use): with the SVD ``V diag(sv) W`` the result is ``blockdiag(V, I) @ [[diag(sv), L], [L, diag(sv)]] @ blockdiag(W, I)`` where
    ``L = diag(sqrt(max(0, 1 - sv^2)))`` moves lost particles into the m extra modes."""
    A = np.asarray(interferometer_matrix, dtype=np.complex128)
    m = A.shape[0]
    V, sv, W = np.linalg.svd(A)
    eye, zero = np.eye(m), np.zeros((m, m), dtype=V.dtype)
    loss = np.sqrt(np.clip(1.0 - np.array([x ** 2 for x in sv]), 0.0, None))
    core = np.block([[np.diag(sv), np.diag(loss)], [np.diag(loss), np.diag(sv)]])
    return np.block([[V, zero], [zero, eye]]) @ core @ np.block([[W, zero], [zero, eye]])


class EffectiveScatteringMatrixCalculator:
    """Rows of ``matrix`` repeated by the output occupation, columns by the input occupation;
    ``[]`` when either side is empty (reference :606-607)."""

    def __init__(self, matrix, input_state: Optional[Sequence[int]] = None,
                 output_state: Optional[Sequence[int]] = None) -> None:
        self.matrix = matrix
        self.input_state = [] if input_state is None else input_state
        self.output_state = [] if output_state is None else output_state

    def calculate(self) -> List[np.ndarray]:
        if sum(self.input_state) == 0 or sum(self.output_state) == 0:
            return []
        U = np.asarray(self.matrix, dtype=np.complex128)
        cols = list(mode_occupation_to_mode_assignment(self.input_state))
        rows = list(mode_occupation_to_mode_assignment(self.output_state))
        return list(U[np.ix_(rows, cols)])


def compute_qft_matrix(n: int) -> np.ndarray:
    """n x n quantum Fourier transform, omega^(jk) / sqrt(n) (reference: quantum_computations_utilities.py:159-178)."""
    if n == 0:
        return np.asarray([])
    k = np.arange(n)
    return np.power(np.exp(2j * np.pi / n), np.outer(k, k)) / np.sqrt(n)


def generate_qft_matrix_for_first_m_modes(m: int, all_modes_number: int) -> np.ndarray:
    """QFT on the first m modes, identity on the rest (reference :505-522)."""
    out = np.eye(all_modes_number, dtype=np.complex128)
    if m > 0:
        out[:m, :m] = compute_qft_matrix(m)
    return out


def generate_random_phases_matrix_for_first_m_modes(m: int, all_modes_number: int) -> np.ndarray:
    """diag(e^{2 pi i u_1}, ..., e^{2 pi i u_m}, 1, ..., 1) with u ~ numpy.random.rand(m) (reference :525-543)."""
    phases = np.ones(all_modes_number, dtype=np.complex128)
    phases[:m] = np.exp(1j * 2 * np.pi * np.random.rand(m))
    return np.diag(phases)
