"""Generate the golden fixtures in this directory from the UNMODIFIED Python reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports theboss from /root/reference with the ``guancodes`` stand-in of oracle/refshim on the
path (the only un-vendored dependency of the permanent path, see oracle/refshim/guancodes), runs
the reference's own calculators / strategies on seeded inputs and stores inputs + outputs as
small ``.npz`` / ``.json`` files.  Random decisions of the samplers are injected through a
*decision tape* by monkey-patching the names the strategy modules imported from numpy.random
(SURVEY.md Appendix B); the patched ``choice`` is first verified to be identical to the real
``numpy.random.choice`` on 2000 seeds.
"""
import json
import os
import sys

import numpy as np
from scipy.stats import unitary_group

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REPO, "oracle", "refshim")]

from theboss.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import (  # noqa: E402
    BSPermanentCalculatorFactory,
    PermanentCalculatorType,
)
from theboss.boson_sampling_utilities.permanent_calculators.bs_cc_ch_submatrices_permanent_calculator import (  # noqa: E402
    BSCCCHSubmatricesPermanentCalculator,
)
from theboss.boson_sampling_utilities.permanent_calculators.bs_cc_ryser_submatrices_permanent_calculator import (  # noqa: E402
    BSCCRyserSubmatricesPermanentCalculator,
)
from theboss.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import (  # noqa: E402
    GlynnGrayPermanentCalculator,
)
from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import (  # noqa: E402
    RyserPermanentCalculator,
)
import theboss.simulation_strategies.generalized_cliffords_b_simulation_strategy as gccb_mod  # noqa: E402
import theboss.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy as gccbu_mod  # noqa: E402
import theboss.simulation_strategies.generalized_cliffords_simulation_strategy as gcc_mod  # noqa: E402
from theboss.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (  # noqa: E402
    LossyNetworksGeneralizedCliffordsSimulationStrategy,
)
from theboss.boson_sampling_utilities.boson_sampling_utilities import (  # noqa: E402
    generate_lossy_n_particle_input_states,
    generate_possible_states,
    prepare_interferometer_matrix_in_expanded_space,
)
from theboss.distribution_calculators.bs_exact_distribution_with_uniform_losses import (  # noqa: E402
    BosonSamplingExperimentConfiguration,
    BSDistributionCalculatorWithUniformLosses,
)

CALCS = {
    "classic": PermanentCalculatorType.CLASSIC,
    "glynn": PermanentCalculatorType.GLYNN,
    "chin_huh": PermanentCalculatorType.CHIN_HUH,
    "ryser": PermanentCalculatorType.RYSER,
}


def haar(m, seed):
    return unitary_group.rvs(m, random_state=seed) if m > 1 else np.array([[np.exp(1j * seed)]])


def random_occupation(rng, m, n):
    out = np.zeros(m, dtype=np.int64)
    for j in rng.randint(0, m, n):
        out[j] += 1
    return out


# ------------------------------------------------------------------------------------------------
def single_permanents():
    cases = []
    # the 9 (input, output) cases of the reference's tests/test_bs_permanent_calculators.py:88-122
    ref_cases = [
        ([1, 1, 1, 1], [1, 1, 1, 1]),
        ([1, 1, 0, 1], [1, 0, 1, 1]),     # gaps
        ([2, 1, 0, 1], [1, 1, 1, 1]),     # bunched input
        ([1, 1, 1, 1], [2, 0, 1, 1]),     # bunched output
        ([2, 0, 1, 1], [1, 3, 0, 0]),     # both
        ([0, 1, 0, 1], [0, 1, 1, 0]),     # skipping modes
        ([0, 2, 0, 1], [0, 1, 0, 2]),
        ([0, 0, 0, 3], [3, 0, 0, 0]),
        ([0, 0, 0, 0], [0, 0, 0, 0]),     # zero particles
    ]
    U4 = haar(4, 4)
    for s, t in ref_cases:
        cases.append((U4, np.array(s), np.array(t)))
    rng = np.random.RandomState(12345)
    for m, n in [(1, 1), (2, 1), (2, 2), (3, 3), (5, 3), (6, 4), (6, 6), (8, 5), (8, 7), (10, 5), (10, 8),
                 (12, 9), (12, 10)]:
        U = haar(m, 100 + m)
        for rep in range(3):
            if rep == 0 and n <= m:
                s = np.array([1] * n + [0] * (m - n))
                t = np.zeros(m, dtype=np.int64)
                t[rng.choice(m, n, replace=False)] = 1
            else:
                s, t = random_occupation(rng, m, n), random_occupation(rng, m, n)
            cases.append((U, s, t))
    # non-unitary complex Gaussian matrices (general complex128 input, not only unitaries)
    for m, n in [(5, 5), (7, 6)]:
        G = rng.randn(m, m) + 1j * rng.randn(m, m)
        cases.append((G, random_occupation(rng, m, n), random_occupation(rng, m, n)))

    out = {"n_cases": len(cases)}
    for idx, (U, s, t) in enumerate(cases):
        out[f"U_{idx}"] = np.asarray(U, dtype=np.complex128)
        out[f"s_{idx}"] = s
        out[f"t_{idx}"] = t
        n = int(s.sum())
        for name, typ in CALCS.items():
            if name == "classic" and n > 8:
                continue
            calc = BSPermanentCalculatorFactory(U, list(s), list(t), typ).generate_calculator()
            out[f"{name}_{idx}"] = np.complex128(calc.compute_permanent())
    np.savez_compressed(os.path.join(HERE, "single_permanents.npz"), **out)
    print("single_permanents:", len(cases), "cases")


# ------------------------------------------------------------------------------------------------
def submatrices_permanents():
    cases = []
    U4 = haar(4, 44)
    # tests/test_bs_submatrices_permanent_calculators.py:41-73
    cases.append((U4, np.array([1, 1, 1, 1]), np.array([0, 1, 1, 1])))
    cases.append((U4, np.array([1.0, 3.0, 0.0, 0.0]), np.array([0, 1, 1, 1])))   # float occupations (:45-51)
    cases.append((U4, np.array([0, 1, 0, 0]), np.array([0, 0, 0, 0])))           # k = 1 edge case (:53-57)
    rng = np.random.RandomState(777)
    for m, k in [(2, 2), (3, 2), (5, 3), (6, 4), (6, 6), (8, 5), (8, 8), (10, 6), (10, 9), (12, 10)]:
        U = haar(m, 200 + m)
        for rep in range(3):
            if rep == 0 and k <= m:
                s = np.array([1] * k + [0] * (m - k))
            else:
                s = random_occupation(rng, m, k)
            t = random_occupation(rng, m, k - 1)
            cases.append((U, s, t))
    # bunched output in one mode (SURVEY A.8 check case)
    U6 = haar(6, 66)
    cases.append((U6, np.array([1, 1, 1, 1, 1, 0]), np.array([0, 0, 0, 4, 0, 0])))
    out = {"n_cases": len(cases)}
    for idx, (U, s, t) in enumerate(cases):
        out[f"U_{idx}"] = np.asarray(U, dtype=np.complex128)
        out[f"s_{idx}"] = s
        out[f"t_{idx}"] = t
        out[f"ryser_{idx}"] = np.array(BSCCRyserSubmatricesPermanentCalculator(U, s, t).compute_permanents(), dtype=np.complex128)
        out[f"chin_huh_{idx}"] = np.array(BSCCCHSubmatricesPermanentCalculator(U, s, t).compute_permanents(), dtype=np.complex128)
    np.savez_compressed(os.path.join(HERE, "submatrices_permanents.npz"), **out)
    print("submatrices_permanents:", len(cases), "cases")


# ------------------------------------------------------------------------------------------------
def exact_distribution():
    """tests/test_exact_distribution_calculator.py:19-142: the reference's only literal known-answer
    vector that flows through a permanent calculator (Chin-Huh)."""
    P = np.array([[0, 0, 1, 0, 0], [1, 0, 0, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1], [0, 1, 0, 0, 0]], dtype=np.complex128)
    s0 = [1, 1, 1, 0, 0]
    cfg = BosonSamplingExperimentConfiguration(
        interferometer_matrix=P, initial_state=s0, number_of_modes=5, initial_number_of_particles=3,
        number_of_particles_lost=2, number_of_particles_left=1, uniform_transmissivity=0.8)
    calc = BSPermanentCalculatorFactory(None, None, None, PermanentCalculatorType.CHIN_HUH).generate_calculator()
    dc = BSDistributionCalculatorWithUniformLosses(cfg, calc)
    dist = [float(x) for x in dc.calculate_distribution()]
    literal = [0.008, 0.032, 0.032, 0.0, 0.0, 0.032, 0.0, 0.128, 0.0, 0.0, 0.128, 0.0, 0.0, 0.0, 0.128] + [0.0] * 14 + [0.512] + [0.0] * 26
    assert len(literal) == 56 and np.allclose(dist, literal), "reference no longer reproduces its own golden vector"
    outcomes = [list(map(int, o)) for o in generate_possible_states(3, 5, losses=True)]
    lossy_inputs = {}
    for l in range(1, 4):
        lossy_inputs[str(l)] = [list(map(int, x)) for x in generate_lossy_n_particle_input_states(s0, l)]
    with open(os.path.join(HERE, "exact_distribution.json"), "w") as f:
        json.dump({"matrix_real": P.real.tolist(), "initial_state": s0, "eta": 0.8, "outcomes": outcomes,
                   "lossy_inputs": lossy_inputs, "reference_literal": literal, "reference_computed": dist}, f)
    print("exact_distribution: 56 outcomes")


# ------------------------------------------------------------------------------------------------
class Tape:
    """Feeds the reference strategies' random decisions from an explicit array."""

    def __init__(self, tape):
        self.tape, self.sample, self.k = tape, 0, 0
        self.pmfs = []

    def start(self, sample):
        self.sample, self.k = sample, 0

    def randint(self, low, high):
        assert low == 0
        return int(self.tape[self.sample, 1 + 2 * self.k] * high)

    def choice(self, a, p):
        p = np.array(p, dtype=np.float64)
        self.pmfs.append(p.copy())
        u = self.tape[self.sample, 2 + 2 * self.k]
        self.k += 1
        cdf = p.cumsum()
        cdf /= cdf[-1]
        return int(cdf.searchsorted(u, side="right"))

    def random(self):
        return float(self.tape[self.sample, 0])


def verify_choice_model():
    rng = np.random.RandomState(5)
    for seed in range(2000):
        m = 2 + seed % 11
        p = rng.rand(m)
        p /= p.sum()
        np.random.seed(seed)
        real = np.random.choice(range(m), p=p)
        u = np.random.RandomState(seed).random_sample()
        cdf = p.cumsum()
        cdf /= cdf[-1]
        assert real == cdf.searchsorted(u, side="right")
    print("numpy.random.choice model verified on 2000 seeds")


def run_with_tape(strategy, module, input_state, tape, uses_random=False):
    t = Tape(tape)
    saved = (gccb_mod.randint, gccb_mod.choice)
    gccb_mod.randint, gccb_mod.choice = t.randint, t.choice
    if uses_random:
        saved_r = gccbu_mod.random
        gccbu_mod.random = t.random
    try:
        samples = []
        for i in range(tape.shape[0]):
            t.start(i)
            samples.append(np.array(strategy.simulate(input_state, 1)[0], dtype=np.int64))
    finally:
        gccb_mod.randint, gccb_mod.choice = saved
        if uses_random:
            gccbu_mod.random = saved_r
    return np.array(samples), t.pmfs


def gccb_samples():
    verify_choice_model()
    out = {}
    rng = np.random.RandomState(2024)
    cases = [
        ("plain_m6_n4", haar(6, 6), [1, 1, 1, 1, 0, 0], 40),
        ("plain_m5_bunched", haar(5, 5), [2, 0, 1, 2, 0], 40),
        ("plain_m8_n6", haar(8, 8), [1, 1, 1, 1, 1, 1, 0, 0], 24),
        ("plain_m10_n8", haar(10, 10), [1] * 8 + [0] * 2, 8),
    ]
    names = []
    for name, U, s, S in cases:
        n = sum(s)
        tape = rng.random_sample((S, 1 + 2 * n))
        calc = RyserPermanentCalculator(U.copy(), None, None)
        strat = gccb_mod.GeneralizedCliffordsBSimulationStrategy(calc)
        samples, pmfs = run_with_tape(strat, gccb_mod, s, tape)
        out[f"{name}_U"], out[f"{name}_s"], out[f"{name}_tape"] = U, np.array(s), tape
        out[f"{name}_samples"], out[f"{name}_pmfs"] = samples, np.array(pmfs)
        names.append(name)
    # uniform losses (no reference test exists for this class; generalized_cliffords_b_uniform_losses...:32)
    U, s, S, eta = haar(6, 16), [1, 1, 1, 1, 1, 0], 60, 0.6
    tape = rng.random_sample((S, 1 + 2 * sum(s)))
    strat = gccbu_mod.GeneralizedCliffordsBUniformLossesSimulationStrategy(RyserPermanentCalculator(U.copy(), None, None), eta)
    samples, _ = run_with_tape(strat, gccbu_mod, np.array(s), tape, uses_random=True)
    out["uniform_U"], out["uniform_s"], out["uniform_tape"], out["uniform_eta"] = U, np.array(s), tape, eta
    out["uniform_samples"] = samples
    # lossy network (m -> 2m dilation)
    m = 5
    U = haar(m, 25) @ np.diag(np.sqrt(np.linspace(0.4, 0.9, m)))
    s, S = [1, 1, 1, 1, 0], 40
    tape = rng.random_sample((S, 1 + 2 * sum(s)))
    strat = LossyNetworksGeneralizedCliffordsSimulationStrategy(RyserPermanentCalculator(U.copy(), None, None))
    samples, _ = run_with_tape(strat, gccb_mod, np.array(s), tape)
    out["lossynet_U"], out["lossynet_s"], out["lossynet_tape"], out["lossynet_samples"] = U, np.array(s), tape, samples
    out["lossynet_expanded"] = prepare_interferometer_matrix_in_expanded_space(U)
    out["plain_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "gccb_samples.npz"), **out)
    print("gccb_samples: done")


def gcc_samples():
    """BASELINE config 1: GCC, n=5, m=10, Haar(10, seed 2024), 1000 samples, Glynn calculator."""
    out = {}
    for name, U, s, S, calc_cls in [
        ("c1_glynn", haar(10, 2024), [1] * 5 + [0] * 5, 1000, GlynnGrayPermanentCalculator),
        ("bunched_ryser", haar(5, 55), [2, 1, 0, 1, 0], 200, RyserPermanentCalculator),
    ]:
        n = sum(s)
        uni = np.random.RandomState(7).random_sample((S, n))
        it = iter(uni.reshape(-1))
        saved = gcc_mod.random
        gcc_mod.random = lambda: float(next(it))
        try:
            strat = gcc_mod.GeneralizedCliffordsSimulationStrategy(calc_cls(U.copy(), None, None))
            samples = np.array(strat.simulate(s, S), dtype=np.int64)
            pm_keys = np.array(list(strat.pmfs.keys()), dtype=np.int64)
            pm_vals = np.array([strat.pmfs[tuple(k)] for k in pm_keys], dtype=np.float64)
        finally:
            gcc_mod.random = saved
        out[f"{name}_U"], out[f"{name}_s"], out[f"{name}_uniforms"] = U, np.array(s), uni
        out[f"{name}_samples"] = samples.astype(np.int8)
        out[f"{name}_pmf_keys"], out[f"{name}_pmf_vals"] = pm_keys.astype(np.int8), pm_vals
    np.savez_compressed(os.path.join(HERE, "gcc_samples.npz"), **out)
    print("gcc_samples: done")


if __name__ == "__main__":
    single_permanents()
    submatrices_permanents()
    exact_distribution()
    gccb_samples()
    gcc_samples()
