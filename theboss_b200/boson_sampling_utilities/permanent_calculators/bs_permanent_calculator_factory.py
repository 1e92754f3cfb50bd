"""Factory of the GPU-backed single-permanent calculators.

Same names, enum values, defaults and fallback as the reference
(theboss/boson_sampling_utilities/permanent_calculators/bs_permanent_calculator_factory.py:29-115):
default type RYSER (:42), unknown types fall back to Chin-Huh (:84-86).
"""
import enum

from .bs_permanent_calculator_interface import BSPermanentCalculatorInterface
from .chin_huh_permanent_calculator import ChinHuhPermanentCalculator
from .classic_permanent_calculator import ClassicPermanentCalculator
from .glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
from .ryser_permanent_calculator import RyserPermanentCalculator


class PermanentCalculatorType(enum.IntEnum):
    CLASSIC = enum.auto()
    GLYNN = enum.auto()
    CHIN_HUH = enum.auto()
    RYSER = enum.auto()


_CLASSES = {
    PermanentCalculatorType.CLASSIC: ClassicPermanentCalculator,
    PermanentCalculatorType.GLYNN: GlynnGrayPermanentCalculator,
    PermanentCalculatorType.CHIN_HUH: ChinHuhPermanentCalculator,
    PermanentCalculatorType.RYSER: RyserPermanentCalculator,
}


class BSPermanentCalculatorFactory:
    def __init__(self, matrix, input_state, output_state,
                 calculator_type: PermanentCalculatorType = PermanentCalculatorType.RYSER, device: int = 0):
        self.matrix = matrix
        self.input_state = input_state
        self.output_state = output_state
        self._calculator_type = calculator_type
        self._device = device

    def generate_calculator(self) -> BSPermanentCalculatorInterface:
        cls = _CLASSES.get(self._calculator_type, ChinHuhPermanentCalculator)
        return cls(matrix=self.matrix, input_state=self.input_state, output_state=self.output_state,
                   device=self._device)
