"""Shared by the CPU and GPU statistics tests of the BOBS strategies: the cases of
tests/golden/bobs_frequencies.json (outcome frequencies of the UNMODIFIED reference, tests/golden/make_bobs_golden.py),
the drop-in strategy built for a case, and the two-sample total-variation acceptance bound."""
import json
import os
from collections import Counter
from math import comb

import numpy as np

from tests import workloads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bobs_frequencies.json")


def load_cases():
    with open(GOLDEN) as f:
        g = json.load(f)
    return g["samples"], g["cases"]


def case_matrix(case) -> np.ndarray:
    """Same recipe as make_bobs_golden.case_matrix."""
    U = workloads.haar(len(case["state"]), case["haar_seed"])
    if case["kind"] == "nla":
        U = U @ np.diag(np.sqrt(np.array(case["etas"])))
    return np.ascontiguousarray(U)


class PlainCalculator:
    """The strategies only read ``matrix`` / ``input_state`` (and ``device``) of the calculator they are given."""

    def __init__(self, matrix, state, device=0):
        self.matrix, self.input_state, self.output_state, self.device = matrix, list(state), list(state), device


def build_strategy(case, **overrides):
    from theboss_b200.simulation_strategies.lossy_state_approximated_simulation_strategy import (
        LossyStateApproximationSimulationStrategy)
    from theboss_b200.simulation_strategies.nonuniform_losses_approximation_strategy import (
        NonuniformLossesApproximationStrategy)
    case = dict(case, **overrides)
    calc = PlainCalculator(case_matrix(case), case["state"])
    if case["kind"] == "nla":
        return NonuniformLossesApproximationStrategy(calc, case["approximated_modes"])
    return LossyStateApproximationSimulationStrategy(calc, case["eta"], case["hierarchy_level"])


def outcomes_count(case) -> int:
    """Number of possible outcomes: at most n particles in m modes (losses allowed)."""
    m, n = len(case["state"]), sum(case["state"])
    return sum(comb(left + m - 1, m - 1) for left in range(n + 1))


def tvd_to_reference(samples, case) -> float:
    counts = Counter(",".join(str(int(v)) for v in x) for x in samples)
    ref = case["frequencies"]
    keys = set(counts) | set(ref)
    return 0.5 * sum(abs(counts.get(k, 0) / len(samples) - ref.get(k, 0.0)) for k in keys)


def tvd_bound(n_outcomes: int, samples: int, delta: float = 1e-3) -> float:
    """TVD between an empirical distribution of `samples` draws and its source stays below this with probability
    1 - delta (the bound the reference's own sampling tests use, tests/simulation_strategies_tests_common.py:192-208)."""
    return float(np.sqrt((-np.log(delta) + n_outcomes * np.log(2)) / (2 * samples)))


def chi2_pvalue(samples, case, ref_samples: int, min_count: int = 25) -> float:
    """Two-sample chi-square test (sharper than the TVD bound): p-value of "both samples come from one distribution".
    Outcomes whose pooled count is below `min_count` are merged into one bin."""
    from scipy.stats import chi2_contingency
    counts = Counter(",".join(str(int(v)) for v in x) for x in samples)
    ref = {k: int(round(v * ref_samples)) for k, v in case["frequencies"].items()}
    rows, rare = [], [0, 0]
    for k in sorted(set(counts) | set(ref)):
        a, b = counts.get(k, 0), ref.get(k, 0)
        if a + b < min_count:
            rare[0] += a
            rare[1] += b
        else:
            rows.append((a, b))
    if rare[0] + rare[1] > 0:
        rows.append(tuple(rare))
    return float(chi2_contingency(np.array(rows).T)[1])
