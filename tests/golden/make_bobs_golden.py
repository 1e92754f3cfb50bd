"""Outcome frequencies of the UNMODIFIED reference's BOBS samplers on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_bobs_golden.py

The two strategies draw a fresh random matrix and a fresh lossy input per sample, so there is no decision tape to
share with them: parity of this row is statistical.  The script calls the per-worker body of each strategy
(`_simulate_in_parallel`, nonuniform_losses_approximation_strategy.py:263-296 and
lossy_state_approximated_simulation_strategy.py:287-310) in THIS process after doing what `simulate` does before it
fans out to its spawn pool (:215-257 resp. :96-118) -- same code, same distributions, but seedable and without
16 interpreter start-ups -- and stores the observed frequency of every outcome in `bobs_frequencies.json`.
tests/test_host_logic.py (oracle permanents underneath, CPU) and tests/test_gpu_zz_reference_runs.py
(CUDA kernels underneath) compare the drop-in strategies against these frequencies.
"""
import json
import os
import sys
import time
from collections import Counter

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REPO, REF, os.path.join(REPO, "oracle", "refshim")]

from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator  # noqa: E402
from theboss.simulation_strategies.lossy_state_approximated_simulation_strategy import (  # noqa: E402
    LossyStateApproximationSimulationStrategy,
)
from theboss.simulation_strategies.nonuniform_losses_approximation_strategy import (  # noqa: E402
    NonuniformLossesApproximationStrategy,
)
from tests import workloads  # noqa: E402

SAMPLES = int(os.environ.get("BOBS_SAMPLES", "40000"))

# name -> (kind, matrix recipe, input state, strategy parameters); tests rebuild the matrices from the same recipe
CASES = {
    # non-uniformly lossy 5-mode network, the two approximated modes hold one particle each
    "nla_m5_k2": dict(kind="nla", haar_seed=71, etas=[0.5, 0.6, 0.7, 0.8, 0.9], state=[1, 1, 1, 0, 0], approximated_modes=2),
    # same network, bunched input inside and outside the approximated modes, three modes approximated
    "nla_m5_k3_bunched": dict(kind="nla", haar_seed=72, etas=[0.9, 0.5, 0.8, 0.6, 0.7], state=[2, 0, 1, 1, 0], approximated_modes=3),
    # uniform losses applied to the state, first two modes kept exact
    "lsa_m5_hl2": dict(kind="lsa", haar_seed=81, eta=0.7, state=[1, 1, 1, 1, 0], hierarchy_level=2),
    # only the last mode approximated, bunched input (hierarchy_level = m makes the reference itself raise a TypeError:
    # it hstacks an empty list onto the state, :312-327, and the float array then fails in range())
    "lsa_m4_hl3_bunched": dict(kind="lsa", haar_seed=82, eta=0.6, state=[1, 0, 1, 2], hierarchy_level=3),
}


def case_matrix(case) -> np.ndarray:
    U = workloads.haar(len(case["state"]), case["haar_seed"])
    if case["kind"] == "nla":
        U = U @ np.diag(np.sqrt(np.array(case["etas"])))
    return np.ascontiguousarray(U)


def run_reference(case, samples: int):
    U, s = case_matrix(case), list(case["state"])
    calc = RyserPermanentCalculator(U, list(s), list(s))
    if case["kind"] == "nla":
        strat = NonuniformLossesApproximationStrategy(calc, case["approximated_modes"], threads_number=1)
        # what simulate() does before the pool (:215-243)
        strat._state_without_approximated_modes = list(s)
        for i in range(strat._approximated_modes_number):
            strat._approximated_modes_particles_number += s[i]
            strat._state_without_approximated_modes[i] = 0
        strat._binomial_weights = strat._compute_binomial_weights(max(max(s), strat._approximated_modes_particles_number))
    else:
        strat = LossyStateApproximationSimulationStrategy(calc, case["eta"], case["hierarchy_level"], threads_number=1)
        strat._prepare_not_approximated_lossy_mixed_state(s[: case["hierarchy_level"]])      # :96-103
        strat._prepare_approximated_input_state(s[case["hierarchy_level"]:])
    return strat._simulate_in_parallel(samples)


def main():
    out = {"samples": SAMPLES, "generator": "tests/golden/make_bobs_golden.py", "cases": {}}
    for seed, (name, case) in enumerate(CASES.items()):
        np.random.seed(1000 + seed)
        import random
        random.seed(1000 + seed)
        t0 = time.time()
        samples = run_reference(case, SAMPLES)
        counts = Counter(tuple(int(x) for x in smp) for smp in samples)
        out["cases"][name] = dict(case, frequencies={",".join(map(str, k)): v / SAMPLES for k, v in sorted(counts.items())})
        print(f"{name}: {len(counts)} distinct outcomes, {time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(HERE, "bobs_frequencies.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
