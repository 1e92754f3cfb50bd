"""Decision tapes: the explicit form of the random draws of the GCC-B family.

Layout (include/bossperm.h): ``tape[sample, 0]`` particle-number uniform (uniform-loss variant only),
``tape[sample, 1 + 2k]`` u_pick of step k (index ``floor(u_pick * #remaining)`` of the remaining input
particles), ``tape[sample, 2 + 2k]`` the uniform ``numpy.random.choice`` consumes at step k.
"""
import numpy as np


def numpy_compatible_tape(n: int, samples_number: int, uniform_losses_weights=None) -> np.ndarray:
    """Consumes numpy's GLOBAL generator in exactly the order the reference strategies do, so that a run
    after ``numpy.random.seed(x)`` makes the same decisions as the reference after the same seed:
    per sample [one ``random()`` for the particle number, uniform-loss variant only
    (generalized_cliffords_b_uniform_losses_simulation_strategy.py:76)], then per particle one
    ``randint(0, #remaining)`` (generalized_cliffords_b_simulation_strategy.py:102-105) and the one
    ``random_sample()`` that ``numpy.random.choice`` draws (:107-110)."""
    tape = np.zeros((samples_number, 1 + 2 * n), dtype=np.float64)
    for i in range(samples_number):
        steps = n
        if uniform_losses_weights is not None:
            u = np.random.random()
            tape[i, 0] = u
            steps, run = 0, 0
            for w in uniform_losses_weights:
                run += w
                if run > u:
                    break
                steps += 1
            steps = min(steps, n)
        for k in range(steps):
            remaining = n - k
            pick = np.random.randint(0, remaining)
            tape[i, 1 + 2 * k] = (pick + 0.5) / remaining   # floor(u * remaining) == pick
            tape[i, 2 + 2 * k] = np.random.random_sample()
    return tape
