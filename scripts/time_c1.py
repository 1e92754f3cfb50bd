"""BASELINE config 1 through the strategy class: GCC (version A), n=5, m=10, 1000 samples, Glynn calculator."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
from theboss_b200.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
U = workloads.haar(10, 2024)
strat = GeneralizedCliffordsSimulationStrategy(GlynnGrayPermanentCalculator(U, None, None, device=0))
for rep in range(3):
    np.random.seed(7)
    t0 = time.perf_counter()
    out = strat.simulate([1] * 5 + [0] * 5, 1000)
    dt = time.perf_counter() - t0
    print(f"c1: 1000 samples in {dt*1e3:.1f} ms = {1000/dt:.0f} samples/s, pmf layers {len(strat.pmfs)}, first {out[0]}")
