"""Generalized Clifford & Clifford sampler, version B, on the B200.

Drop-in for ``GeneralizedCliffordsBSimulationStrategy``
(theboss/simulation_strategies/generalized_cliffords_b_simulation_strategy.py:33-110).  The whole
chain-rule loop (:94-110) runs on the device for all requested samples at once: per step one launch of
the minors kernel K3 over (chunks x samples) blocks and one finish kernel that forms the pmf of
:82-92, draws the output mode and admits the next input particle (theboss_b200/csrc/sampler_kernel.cu).

Randomness.  The reference draws from numpy's global generator.  ``rng_mode``:

* ``"philox"`` (default): a 64-bit seed is drawn from numpy's global generator (so
  ``numpy.random.seed`` still makes runs reproducible) and the device fills the decision tape from a
  counter-based Philox stream keyed by (seed, sample, slot);
* ``"numpy"``: the host consumes numpy's global generator in exactly the reference's call order
  (decision_tape.numpy_compatible_tape), giving the reference's decisions for the same seed;
* an explicit tape can be passed to ``simulate(..., decision_tape=tape)``.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import _native
from ..boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_interface import (
    BSPermanentCalculatorInterface,
)
from .decision_tape import numpy_compatible_tape
from .simulation_strategy_interface import SimulationStrategyInterface


class GeneralizedCliffordsBSimulationStrategy(SimulationStrategyInterface):
    def __init__(self, bs_permanent_calculator: BSPermanentCalculatorInterface, rng_mode: str = "philox",
                 device: Optional[int] = None) -> None:
        if rng_mode not in ("philox", "numpy"):
            raise ValueError("rng_mode must be 'philox' or 'numpy'")
        self._bs_permanent_calculator = bs_permanent_calculator
        self._rng_mode = rng_mode
        self._device = getattr(bs_permanent_calculator, "device", 0) if device is None else int(device)
        self.input_state: Sequence[int] = []
        self.number_of_input_photons: int = 0

    def set_new_matrix(self, new_matrix) -> None:
        """generalized_cliffords_simulation_strategy.py:40-47 of the reference."""
        self._bs_permanent_calculator.matrix = new_matrix

    # -- pieces shared with the subclasses -------------------------------------------------------
    def _operands(self, input_state):
        U = _native.as_matrix(self._bs_permanent_calculator.matrix)   # re-read: may have been mutated in place
        if U.shape[0] != U.shape[1] or len(input_state) > U.shape[0]:
            raise AttributeError("matrix / input state shapes do not match")
        return U, _native.as_state(input_state, U.shape[0])

    def _uniform_losses_weights(self):
        return None

    def _eta(self) -> float:
        return -1.0

    def _run(self, input_state, samples_number: int, decision_tape) -> np.ndarray:
        self.input_state = input_state
        self.number_of_input_photons = int(np.sum(input_state))
        U, s = self._operands(input_state)
        n = int(s.sum())
        seed = 0
        if decision_tape is None:
            if self._rng_mode == "numpy":
                decision_tape = numpy_compatible_tape(n, samples_number, self._uniform_losses_weights())
            else:
                seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        handle = _native.default_handle(self._device)
        out = handle.gccb_simulate(U, s, samples_number, eta=self._eta(), seed=seed, tape=decision_tape)
        return out[:, : len(input_state)]

    def simulate(self, input_state: Sequence[int], samples_number: int = 1,
                 decision_tape: Optional[np.ndarray] = None) -> List[Tuple[int, ...]]:
        out = self._run(input_state, samples_number, decision_tape)
        return [tuple(row) for row in out.tolist()]   # Python ints, converted at C speed

    def compute_pmf(self, current_input: Sequence[int], r_sample: Sequence[int]) -> np.ndarray:
        """The pmf of one step (_compute_pmf, :69-92): probabilities of the next particle's output mode
        given the input particles admitted so far and the outputs sampled so far."""
        U, s = self._operands(current_input)
        t = _native.as_state(r_sample, U.shape[0])
        return _native.default_handle(self._device).gccb_pmf(U, s, t)
