"""Time the UNMODIFIED Python reference on bounded slices of BASELINE.json's configs (SURVEY.md section 8d,
"CPU baseline (1) and (2)").  Build container only: the GPU box has no /root/reference, so the numbers
are committed as a fixture (profiles/r01_reference_python_cpu.json) that bench.py quotes verbatim, labelled
with the host they were measured on.

    python scripts/time_reference_python.py [--quick] > profiles/r01_reference_python_cpu.json

Single process = 1 core (the reference is single-threaded), then the same calls fanned out over all host
cores with multiprocessing over independent items -- the reference's own parallel pattern
(theboss/simulation_strategies/nonuniform_losses_approximation_strategy.py:254-257).  Sizes beyond what
Python finishes in seconds are extrapolated by the term-count formula and SAY SO in the output.
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REPO, "oracle", "refshim"), os.path.join(REPO, "tests")]   # the reference has a `tests` package too

import workloads  # noqa: E402  (tests/workloads.py: the seeded inputs of BASELINE.json's configs)


def _calcs():
    from theboss.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
    from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    return {"glynn": GlynnGrayPermanentCalculator, "chin_huh": ChinHuhPermanentCalculator, "ryser": RyserPermanentCalculator}


def c4_one(n):
    """One Gray-code Glynn permanent of the C4 matrix family at size n (glynn_gray_permanent_calculator.py:41-84)."""
    A = workloads.c4_matrix(n)
    ones = [1] * n
    calc = _calcs()["glynn"](A, ones, ones)
    t0 = time.perf_counter()
    p = calc.compute_permanent()
    return time.perf_counter() - t0, complex(p)


def c2_one(args):
    """One C2 item (n = 20, m = 40, repeated rows and columns) through the named single-permanent calculator."""
    name, i, n, m = args
    U, S, T = workloads.c2_batch(n, m, i + 1)
    calc = _calcs()[name](U, [int(x) for x in S[i]], [int(x) for x in T[i]])
    t0 = time.perf_counter()
    p = calc.compute_permanent()
    return time.perf_counter() - t0, complex(p)


def c3_one(k):
    """One GCC-B step at k particles: BSCCRyserSubmatricesPermanentCalculator.compute_permanents
    (bs_submatrices_permanent_calculator_base.py:150-175), bunched outputs as in workloads.c3_step."""
    from theboss.boson_sampling_utilities.permanent_calculators.bs_cc_ryser_submatrices_permanent_calculator import (
        BSCCRyserSubmatricesPermanentCalculator,
    )
    U, s, t = workloads.c3_step(k, 2 * k)
    calc = BSCCRyserSubmatricesPermanentCalculator(U, np.array(s, dtype=int), np.array(t, dtype=int))
    t0 = time.perf_counter()
    calc.compute_permanents()
    return time.perf_counter() - t0


def gccb_run(args):
    """GCC-B sampling, lossless, n photons in 2n modes (generalized_cliffords_b_simulation_strategy.py:41-67)."""
    n, samples, seed = args
    from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
    U = workloads.haar(2 * n, n)
    np.random.seed(seed)
    strat = GeneralizedCliffordsBSimulationStrategy(RyserPermanentCalculator(U))
    t0 = time.perf_counter()
    strat.simulate([1] * n + [0] * n, samples)
    return time.perf_counter() - t0


def c1_run(samples=1000):
    """BASELINE config 1 as is: GCC, n=5, m=10, Haar(10, seed 2024), Gray-code Glynn calculator."""
    from theboss.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
    from theboss.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
    U = workloads.haar(10, 2024)
    np.random.seed(7)
    strat = GeneralizedCliffordsSimulationStrategy(GlynnGrayPermanentCalculator(U))
    t0 = time.perf_counter()
    out = strat.simulate([1] * 5 + [0] * 5, samples)
    return time.perf_counter() - t0, len(out)


def c5_run(args):
    """C5 at reduced n: (i) uniform losses eta = 0.5, (ii) non-uniform losses through the 2m dilation."""
    kind, n, samples = args
    from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    U, U_lossy, s = workloads.c5_lossy(n, 2 * n)
    np.random.seed(5)
    if kind == "uniform":
        from theboss.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
            GeneralizedCliffordsBUniformLossesSimulationStrategy,
        )
        strat = GeneralizedCliffordsBUniformLossesSimulationStrategy(RyserPermanentCalculator(U), 0.5)
    else:
        from theboss.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
            LossyNetworksGeneralizedCliffordsSimulationStrategy,
        )
        strat = LossyNetworksGeneralizedCliffordsSimulationStrategy(RyserPermanentCalculator(U_lossy))
    t0 = time.perf_counter()
    strat.simulate([int(x) for x in s], samples)
    return time.perf_counter() - t0


def bounded():
    """--bounded: the ~25 s slice bench.py runs LIVE on the bench host (reference from baseline/_ref, THEBOSS_REFERENCE):
    Glynn at N = 14 / 16 / 18 on one core, N = 16 over all cores, BASELINE config 1 as is, one small GCC-B run."""
    cores = os.cpu_count()
    flops = lambda n: (8 * n - 4) * 2.0 ** (n - 1)   # noqa: E731
    out = {"reference": REF, "host": {"cpu": platform.processor() or platform.machine(), "cores": cores,
                                      "python": platform.python_version(), "numpy": np.__version__}}
    c4 = {}
    for n in (14, 16, 18):
        dt, _ = c4_one(n)
        c4[str(n)] = {"seconds": dt, "permanents_per_s": 1.0 / dt, "ns_per_term": dt / 2 ** (n - 1) * 1e9}
    sec30 = c4["18"]["seconds"] * flops(30) / flops(18)
    out["glynn_1core"] = c4
    out["glynn_n30_permanents_per_s_1core_extrapolated"] = 1.0 / sec30
    dt, cnt = c1_run(1000)
    out["c1_gcc_n5_m10_1core"] = {"samples": cnt, "seconds": dt, "samples_per_s": cnt / dt}
    dt = gccb_run((10, 4, 11))
    out["gccb_n10_m20_1core"] = {"samples": 4, "seconds": dt, "samples_per_s": 4 / dt,
                                 "n24_samples_per_s_extrapolated": 4 / dt * (2.0 ** 10 * 10) / (2.0 ** 24 * 24)}
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(c3_one, [4] * cores)   # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(c4_one, [16] * (2 * cores))
        dt = time.perf_counter() - t0
    rate = 2 * cores / dt
    out["glynn_all_cores"] = {"n": 16, "permanents": 2 * cores, "cores": cores, "seconds": dt, "permanents_per_s": rate,
                              "n30_permanents_per_s_extrapolated": rate * flops(16) / flops(30)}
    out["note"] = ("unmodified Python reference (theboss v3.0.1 from baseline/_ref + the guancodes stand-in of oracle/refshim) timed on THIS "
                   "host; n = 30 / n = 24 figures are EXTRAPOLATED by (8N-4) 2^(N-1) resp. 2^n n and say so in their key")
    json.dump(out, sys.stdout)
    print()


def main():
    if "--bounded" in sys.argv:
        return bounded()
    quick = "--quick" in sys.argv
    cores = os.cpu_count()
    out = {
        "what": "unmodified Python reference (Tomev-CTP/theboss v3.0.1 + guancodes stand-in of oracle/refshim), timed in the build "
                "container; the GPU box has no /root/reference, so bench.py quotes this file",
        "host": {"cpu": platform.processor() or platform.machine(), "cores": cores, "python": platform.python_version(),
                 "numpy": np.__version__},
        "single_core": {}, "all_cores": {},
    }
    sc, ac = out["single_core"], out["all_cores"]

    # ---- C4: Glynn at N = 14 .. 20, extrapolated to 30 by (8N - 4) 2^(N-1)
    sizes = [12, 14, 16] if quick else [14, 16, 18, 20]
    c4 = {}
    for n in sizes:
        dt, p = c4_one(n)
        c4[str(n)] = {"seconds": dt, "permanents_per_s": 1.0 / dt, "ns_per_term": dt / 2 ** (n - 1) * 1e9, "re": p.real, "im": p.imag}
    nl = sizes[-1]
    flops = lambda n: (8 * n - 4) * 2.0 ** (n - 1)   # noqa: E731
    sec30 = c4[str(nl)]["seconds"] * flops(30) / flops(nl)
    c4["30_extrapolated"] = {"seconds": sec30, "permanents_per_s": 1.0 / sec30,
                             "note": f"EXTRAPOLATED from N={nl} by (8N-4)*2^(N-1); the reference would also need a 2^29-element Python list"}
    sc["c4_glynn_single_permanent"] = c4

    # ---- C2: 8 items at n = 20 per calculator
    n2, m2, items = (12, 24, 4) if quick else (20, 40, 8)
    c2 = {}
    for name in ("chin_huh", "ryser", "glynn"):
        ts = [c2_one((name, i, n2, m2))[0] for i in range(items)]
        c2[name] = {"items": items, "n": n2, "m": m2, "seconds_mean": float(np.mean(ts)), "permanents_per_s": items / float(np.sum(ts))}
    sc["c2_batched_permanents"] = c2

    # ---- C3: one step at k <= 16, extrapolated to k = 24 by the reference's own term count (2^k Guan terms x k)
    ks = [8, 10] if quick else [12, 14, 16]
    c3 = {}
    for k in ks:
        dt = c3_one(k)
        c3[str(k)] = {"seconds": dt, "steps_per_s": 1.0 / dt}
    kl = ks[-1]
    sec24 = c3[str(kl)]["seconds"] * (2.0 ** 24 * 24) / (2.0 ** kl * kl)
    c3["24_extrapolated"] = {"seconds": sec24, "steps_per_s": 1.0 / sec24,
                             "note": f"EXTRAPOLATED from k={kl} by 2^k * k (collision-free input: 2^k Guan terms, O(k) Python work per term)"}
    sc["c3_submatrices_step"] = c3

    # ---- GCC-B sampling at n <= 12, extrapolated to n = 24 (cost doubles per photon)
    ns = [(6, 8)] if quick else [(8, 16), (10, 8), (12, 4)]
    g = {}
    for n, samples in ns:
        dt = gccb_run((n, samples, 11))
        g[str(n)] = {"samples": samples, "seconds": dt, "samples_per_s": samples / dt}
    nl, sl = ns[-1]
    per = g[str(nl)]["seconds"] / sl
    sec24 = per * (2.0 ** 24 * 24) / (2.0 ** nl * nl)
    g["24_extrapolated"] = {"seconds_per_sample": sec24, "samples_per_s": 1.0 / sec24,
                            "note": f"EXTRAPOLATED from n={nl} by 2^n * n per sample"}
    sc["gccb_sampling_m_2n"] = g

    # ---- C1 as is
    dt, cnt = c1_run(200 if quick else 1000)
    sc["c1_gcc_n5_m10"] = {"samples": cnt, "seconds": dt, "samples_per_s": cnt / dt}

    # ---- C5 at n <= 10
    c5 = {}
    for kind, n, samples in ([("uniform", 6, 8), ("nonuniform", 6, 4)] if quick else [("uniform", 10, 16), ("nonuniform", 8, 8), ("nonuniform", 10, 4)]):
        dt = c5_run((kind, n, samples))
        c5[f"{kind}_n{n}"] = {"samples": samples, "seconds": dt, "samples_per_s": samples / dt}
    sc["c5_lossy_reduced_n"] = c5

    # ---- all cores: independent items over a process pool
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(c3_one, [4] * cores)   # warm the workers (imports)
        n4 = 14 if quick else 18
        t0 = time.perf_counter()
        pool.map(c4_one, [n4] * (2 * cores))
        dt = time.perf_counter() - t0
        rate = 2 * cores / dt
        ac["c4_glynn_single_permanent"] = {
            "n": n4, "permanents": 2 * cores, "seconds": dt, "permanents_per_s": rate,
            "n30_extrapolated_permanents_per_s": rate * flops(n4) / flops(30),
            "note": "independent permanents over a spawn pool (the reference has no way to split ONE permanent); n=30 figure EXTRAPOLATED by (8N-4)*2^(N-1)"}
        work = [("chin_huh", i, n2, m2) for i in range(2 * cores)]
        t0 = time.perf_counter()
        pool.map(c2_one, work)
        dt = time.perf_counter() - t0
        ac["c2_batched_permanents_chin_huh"] = {"items": len(work), "n": n2, "m": m2, "seconds": dt, "permanents_per_s": len(work) / dt}
        ng, sg = (6, 4) if quick else (12, 2)
        t0 = time.perf_counter()
        pool.map(gccb_run, [(ng, sg, 100 + i) for i in range(cores)])
        dt = time.perf_counter() - t0
        rate = cores * sg / dt
        ac["gccb_sampling_m_2n"] = {"n": ng, "samples": cores * sg, "seconds": dt, "samples_per_s": rate,
                                    "n24_extrapolated_samples_per_s": rate * (2.0 ** ng * ng) / (2.0 ** 24 * 24),
                                    "note": "independent samples over a spawn pool; n=24 figure EXTRAPOLATED by 2^n * n"}
    ac["cores"] = cores
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
