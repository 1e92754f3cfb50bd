"""Samples of the UNMODIFIED reference's GCC-B family at the largest sizes its Python loops finish in about a minute,
drawn after `numpy.random.seed(seed)`.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_gccb_seeded_golden.py

The drop-in strategies with ``rng_mode="numpy"`` consume NumPy's global generator in the reference's call order
(theboss_b200/simulation_strategies/decision_tape.py), so the same seed must give these samples bit for bit:
tests/test_oracle_golden.py checks that with the oracle's loop, tests/test_gpu_zz_reference_runs.py with kernels K3 / K4.
(tests/golden/gccb_samples.npz pins the same strategies through injected decision tapes at n <= 8.)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REPO, REF, os.path.join(REPO, "oracle", "refshim")]

from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator  # noqa: E402
from theboss.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy  # noqa: E402
from theboss.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (  # noqa: E402
    GeneralizedCliffordsBUniformLossesSimulationStrategy,
)
from theboss.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (  # noqa: E402
    LossyNetworksGeneralizedCliffordsSimulationStrategy,
)
from tests import workloads  # noqa: E402

# name -> kind, modes, input state, samples, numpy seed (+ transmissivity / per-mode transmissivities)
CASES = {
    "plain_m24_n12": dict(kind="plain", m=24, state=[1] * 12 + [0] * 12, samples=10, seed=101),
    "plain_m20_n12_bunched": dict(kind="plain", m=20, state=[3, 2, 0, 1, 1, 2, 0, 0, 1, 1, 1] + [0] * 9, samples=10, seed=102),
    "plain_m28_n14": dict(kind="plain", m=28, state=[1] * 14 + [0] * 14, samples=4, seed=103),
    "plain_m32_n16": dict(kind="plain", m=32, state=[1] * 16 + [0] * 16, samples=3, seed=106),
    "uniform_m26_n13": dict(kind="uniform", m=26, state=[1] * 13 + [0] * 13, samples=40, seed=104, eta=0.7),
    "lossynet_m8_n6": dict(kind="lossynet", m=8, state=[1, 1, 2, 1, 0, 1, 0, 0], samples=12, seed=105),
}


def case_matrix(name, case) -> np.ndarray:
    U = workloads.haar(case["m"], 500 + case["seed"])
    if case["kind"] == "lossynet":
        U = U @ np.diag(np.sqrt(np.linspace(0.45, 0.95, case["m"])))
    return np.ascontiguousarray(U)


def main():
    out = {"names": np.array(sorted(CASES))}
    for name, c in CASES.items():
        U = case_matrix(name, c)
        calc = RyserPermanentCalculator(U.copy(), None, None)
        if c["kind"] == "plain":
            strategy = GeneralizedCliffordsBSimulationStrategy(calc)
        elif c["kind"] == "uniform":
            strategy = GeneralizedCliffordsBUniformLossesSimulationStrategy(calc, c["eta"])
        else:
            strategy = LossyNetworksGeneralizedCliffordsSimulationStrategy(calc)
        t0 = time.time()
        np.random.seed(c["seed"])
        samples = strategy.simulate(np.array(c["state"]) if c["kind"] == "uniform" else list(c["state"]), c["samples"])
        out[f"{name}_kind"] = np.array(c["kind"])
        out[f"{name}_U"] = U
        out[f"{name}_s"] = np.array(c["state"], dtype=np.int64)
        out[f"{name}_seed"] = np.int64(c["seed"])
        out[f"{name}_eta"] = np.float64(c.get("eta", -1.0))
        out[f"{name}_samples"] = np.array([[int(v) for v in x] for x in samples], dtype=np.int64)
        print(name, out[f"{name}_samples"].shape, f"{time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(os.path.join(HERE, "gccb_seeded_samples.npz"), **out)


if __name__ == "__main__":
    main()
