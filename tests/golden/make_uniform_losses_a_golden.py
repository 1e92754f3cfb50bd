"""Samples and recorded distributions of the UNMODIFIED reference's version-A uniform-loss GCC sampler (SURVEY.md
section 8, row f3) under fixed seeds of BOTH generators it consumes.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_uniform_losses_a_golden.py

``GeneralizedCliffordsUniformLossesSimulationStrategy`` (generalized_cliffords_uniform_losses_simulation_strategy.py:126-138)
draws one stdlib ``random.random()`` per input particle (lost or kept) and one ``numpy.random.random()`` per kept particle
(generalized_cliffords_simulation_strategy.py:249-266), so a drop-in that keeps this call order reproduces the samples
bit for bit.  Stored per case: matrix, input state, transmissivity, both seeds, the samples, and the ``distribution`` /
``unweighted_distribution`` lists the strategy has filled in by the end of the run.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REPO, REF, os.path.join(REPO, "oracle", "refshim")]

from theboss.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator  # noqa: E402
from theboss.simulation_strategies.generalized_cliffords_uniform_losses_simulation_strategy import (  # noqa: E402
    GeneralizedCliffordsUniformLossesSimulationStrategy,
)
from tests import workloads  # noqa: E402

CASES = {
    "m5_n4_bunched": dict(haar_seed=33, state=[1, 1, 2, 0, 0], eta=0.6, samples=400, seeds=(11, 12)),
    "m6_n4": dict(haar_seed=34, state=[1, 1, 1, 1, 0, 0], eta=0.8, samples=300, seeds=(21, 22)),
    "m4_n3_heavy_loss": dict(haar_seed=35, state=[0, 3, 0, 0], eta=0.25, samples=300, seeds=(31, 32)),
}


def main():
    out = {"names": np.array(sorted(CASES))}
    for name, c in CASES.items():
        U = workloads.haar(len(c["state"]), c["haar_seed"])
        strat = GeneralizedCliffordsUniformLossesSimulationStrategy(ChinHuhPermanentCalculator(U.copy()), c["eta"])
        random.seed(c["seeds"][0])
        np.random.seed(c["seeds"][1])
        samples = strat.simulate(list(c["state"]), c["samples"])
        out[f"{name}_U"] = U
        out[f"{name}_s"] = np.array(c["state"], dtype=np.int64)
        out[f"{name}_eta"] = np.float64(c["eta"])
        out[f"{name}_seeds"] = np.array(c["seeds"], dtype=np.int64)
        out[f"{name}_samples"] = np.array(samples, dtype=np.int64)
        out[f"{name}_distribution"] = np.array(strat.distribution, dtype=np.float64)
        out[f"{name}_unweighted"] = np.array(strat.unweighted_distribution, dtype=np.float64)
        print(name, np.array(samples).shape, "sum(distribution) =", float(np.sum(strat.distribution)))
    np.savez_compressed(os.path.join(HERE, "gcc_uniform_losses_samples.npz"), **out)


if __name__ == "__main__":
    main()
