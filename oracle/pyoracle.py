"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/bossperm_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package ``theboss_b200`` never does (its ops fail loudly
when the CUDA library is missing instead of falling back to anything here).

The sampling loops below restate the reference's GCC-family strategies on top of the C routines,
with every random decision taken from an explicit per-sample *decision tape* (SURVEY.md
Appendix B) instead of NumPy's global generator, so that the same tape can be fed to the CUDA
path.  All citations are relative to /root/reference/theboss/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import factorial
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbossperm_oracle.so")
_lib = None

RYSER, CHIN_HUH = 0, 1


def build(force: bool = False) -> str:
    """Compile the C restatement in place (gcc, oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
        max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("bossperm_oracle.c", "oracle_impl.h"))
        > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        for sfx in ("d", "ld"):
            getattr(L, f"orc_glynn_gray_{sfx}").argtypes = [dp, C.c_int, dp]
            getattr(L, f"orc_glynn_gray_par_{sfx}").argtypes = [dp, C.c_int, C.c_int, C.c_int, dp]
            getattr(L, f"orc_glynn_gray_range_{sfx}").argtypes = [dp, C.c_int, C.c_uint64, C.c_uint64, dp]
            getattr(L, f"orc_guan_permanent_{sfx}").argtypes = [dp, C.c_int, ip, ip, C.c_int, dp]
            getattr(L, f"orc_submatrices_{sfx}").argtypes = [dp, C.c_int, ip, ip, C.c_int, dp]
            getattr(L, f"orc_gccb_pmf_{sfx}").argtypes = [dp, C.c_int, ip, ip, dp, dp]
        L.orc_effective_matrix.argtypes = [dp, C.c_int, ip, ip, dp]
        L.orc_classic.argtypes = [dp, C.c_int, dp]
        L.orc_numpy_choice.argtypes = [dp, C.c_int, C.c_double]
        L.orc_gcc_draw.argtypes = [dp, C.c_int, C.c_double]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _mat(U) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(U, dtype=np.complex128))


def _state(s, m: int) -> np.ndarray:
    out = np.zeros(m, dtype=np.int32)
    s = np.asarray(s)
    out[: len(s)] = s.astype(np.int64)
    return out


def effective_matrix(U, s, t) -> np.ndarray:
    """boson_sampling_utilities/boson_sampling_utilities.py:595-626."""
    U = _mat(U)
    m = U.shape[0]
    s, t = _state(s, m), _state(t, m)
    n = int(s.sum())
    A = np.zeros((max(n, 1), max(n, 1)), dtype=np.complex128)
    N = lib().orc_effective_matrix(_dp(U.view(np.float64)), m, _ip(s), _ip(t), _dp(A.view(np.float64)))
    if N < 0:
        raise ValueError("input and output particle numbers differ")
    return A[:N, :N] if N else np.zeros((0, 0), dtype=np.complex128)


def glynn_matrix(A, precision: str = "ld", nthreads: int = 1, nchunks: int = 0) -> complex:
    """permanent_calculators/glynn_gray_permanent_calculator.py:41-84 on an explicit matrix."""
    A = _mat(A)
    N = A.shape[0]
    out = np.zeros(2)
    if nthreads > 1 or nchunks > 0:
        fn = getattr(lib(), f"orc_glynn_gray_par_{precision}")
        rc = fn(_dp(A.view(np.float64)), N, nchunks if nchunks > 0 else 8 * nthreads, nthreads, _dp(out))
    else:
        rc = getattr(lib(), f"orc_glynn_gray_{precision}")(_dp(A.view(np.float64)), N, _dp(out))
    if rc:
        raise ValueError(f"oracle glynn rc={rc}")
    return complex(out[0], out[1])


def glynn_range(A, lo: int, hi: int, precision: str = "ld") -> complex:
    """Un-normalised partial sum over Gray steps [lo, hi)."""
    A = _mat(A)
    out = np.zeros(2)
    rc = getattr(lib(), f"orc_glynn_gray_range_{precision}")(_dp(A.view(np.float64)), A.shape[0], lo, hi, _dp(out))
    if rc:
        raise ValueError(f"oracle glynn_range rc={rc}")
    return complex(out[0], out[1])


def glynn(U, s, t, precision: str = "ld", nthreads: int = 1) -> complex:
    """GlynnGrayPermanentCalculator(U, s, t).compute_permanent()."""
    A = effective_matrix(U, s, t)
    if A.shape[0] == 0:
        return complex(1)  # glynn_gray_permanent_calculator.py:52-53
    return glynn_matrix(A, precision, nthreads)


def guan_permanent(U, s, t, formula: int, precision: str = "ld") -> complex:
    """Ryser (formula 0) / Chin-Huh (formula 1): bs_permanent_calculator_base.py:166-209."""
    U = _mat(U)
    m = U.shape[0]
    s, t = _state(s, m), _state(t, m)
    out = np.zeros(2)
    rc = getattr(lib(), f"orc_guan_permanent_{precision}")(_dp(U.view(np.float64)), m, _ip(s), _ip(t), formula, _dp(out))
    if rc:
        raise ValueError(f"oracle guan rc={rc}")
    return complex(out[0], out[1])


def classic(U, s, t) -> complex:
    """ClassicPermanentCalculator: classic_permanent_calculator.py:28-67."""
    if int(np.sum(s)) == 0:
        return complex(1) if int(np.sum(t)) == 0 else complex(0)
    A = effective_matrix(U, s, t)
    out = np.zeros(2)
    rc = lib().orc_classic(_dp(np.ascontiguousarray(A).view(np.float64)), A.shape[0], _dp(out))
    if rc:
        raise ValueError(f"oracle classic rc={rc}")
    return complex(out[0], out[1])


def submatrices(U, s, t, formula: int = RYSER, precision: str = "ld") -> np.ndarray:
    """compute_permanents(): bs_submatrices_permanent_calculator_base.py:150-175."""
    U = _mat(U)
    m = U.shape[0]
    s, t = _state(s, m), _state(t, m)
    out = np.zeros(m, dtype=np.complex128)
    rc = getattr(lib(), f"orc_submatrices_{precision}")(_dp(U.view(np.float64)), m, _ip(s), _ip(t), formula, _dp(out.view(np.float64)))
    if rc:
        raise ValueError(f"oracle submatrices rc={rc}")
    return out


def gccb_pmf(U, s_cur, r_sample, precision: str = "d", raw: bool = False):
    """_compute_pmf: simulation_strategies/generalized_cliffords_b_simulation_strategy.py:69-92."""
    U = _mat(U)
    m = U.shape[0]
    s, t = _state(s_cur, m), _state(r_sample, m)
    pmf, w = np.zeros(m), np.zeros(m)
    rc = getattr(lib(), f"orc_gccb_pmf_{precision}")(_dp(U.view(np.float64)), m, _ip(s), _ip(t), _dp(pmf), _dp(w))
    if rc:
        raise ValueError(f"oracle gccb_pmf rc={rc}")
    return (pmf, w) if raw else pmf


def numpy_choice(pmf: np.ndarray, u: float) -> int:
    pmf = np.ascontiguousarray(pmf, dtype=np.float64)
    return int(lib().orc_numpy_choice(_dp(pmf), len(pmf), float(u)))


def gcc_draw(pmf: np.ndarray, u: float) -> int:
    pmf = np.ascontiguousarray(pmf, dtype=np.float64)
    return int(lib().orc_gcc_draw(_dp(pmf), len(pmf), float(u)))


# ---------------------------------------------------------------------------------------------
# Sampling loops with an explicit decision tape.
#
# Tape layout (shared with the CUDA path, include/bossperm.h "decision tape"):
#   tape[sample, 0]          uniform for the particle-number draw (uniform-loss variant only)
#   tape[sample, 1 + 2*k]    uniform u_pick of step k: index floor(u_pick * len(remaining)) of the
#                            remaining input particles (replaces numpy.random.randint(0, len))
#   tape[sample, 2 + 2*k]    uniform u_choice of step k (the one draw numpy.random.choice makes)
# ---------------------------------------------------------------------------------------------


def mode_assignment(s: Sequence[int]) -> List[int]:
    """boson_sampling_utilities.py:61-78."""
    out: List[int] = []
    for i, c in enumerate(s):
        out += [i] * int(c)
    return out


def binomial_weights(n: int, eta: float) -> List[float]:
    """generalized_cliffords_b_uniform_losses_simulation_strategy.py:50-65."""
    from scipy.special import binom

    return [binom(n, l) * pow(eta, l) * pow(1 - eta, n - l) for l in range(n + 1)]


def particles_left(weights: Sequence[float], u: float) -> int:
    """generalized_cliffords_b_uniform_losses_simulation_strategy.py:67-85."""
    left, run = 0, 0
    for w in weights:
        run += w
        if run > u:
            return left
        left += 1
    return left


def gccb_simulate(U, input_state, tape: np.ndarray, n_particles: Optional[Sequence[int]] = None,
                  precision: str = "d", return_pmfs: bool = False):
    """GeneralizedCliffordsBSimulationStrategy.simulate
    (generalized_cliffords_b_simulation_strategy.py:41-67, :94-110) driven by ``tape``.
    ``n_particles[sample]`` limits the number of steps (uniform-loss variant, ...b_uniform_losses
    ...:110-121)."""
    U = _mat(U)
    m = U.shape[0]
    s = _state(input_state, m)
    n = int(s.sum())
    samples, pmfs = [], []
    for i in range(tape.shape[0]):
        cur = np.zeros(m, dtype=np.int32)
        remaining = mode_assignment(s)
        r = np.zeros(m, dtype=np.int32)
        steps = n if n_particles is None else int(n_particles[i])
        for k in range(steps):
            pick = int(tape[i, 1 + 2 * k] * len(remaining))
            cur[remaining.pop(pick)] += 1
            pmf = gccb_pmf(U, cur, r, precision)
            if return_pmfs:
                pmfs.append(pmf)
            r[numpy_choice(pmf, tape[i, 2 + 2 * k])] += 1
        samples.append(tuple(int(x) for x in r))
    return (samples, pmfs) if return_pmfs else samples


def gccb_uniform_losses_simulate(U, input_state, eta: float, tape: np.ndarray, precision: str = "d"):
    """GeneralizedCliffordsBUniformLossesSimulationStrategy.simulate (:87-121)."""
    n = int(np.sum(input_state))
    w = binomial_weights(n, eta)
    left = [particles_left(w, tape[i, 0]) for i in range(tape.shape[0])]
    return gccb_simulate(U, input_state, tape, n_particles=left, precision=precision)


def expanded_matrix(U) -> np.ndarray:
    """prepare_interferometer_matrix_in_expanded_space: boson_sampling_utilities.py:287-342
    (loss-transfer block sqrt(1 - sigma^2), helper at :268-284)."""
    U = _mat(U)
    m = U.shape[0]
    v, sv, u = np.linalg.svd(U)
    z, e = np.zeros_like(v), np.eye(m)
    ev = np.block([[v, z], [z, e]])
    eu = np.block([[u, z], [z, e]])
    tr = np.array([x ** 2 for x in sv])
    losses = 1.0 - tr
    losses[losses < 0] = 0  # boson_sampling_utilities.py:279-282
    lt = np.diag(np.sqrt(losses))
    es = np.block([[np.diag(sv), lt], [lt, np.diag(sv)]])
    return ev @ es @ eu


def lossy_net_simulate(U_lossy, input_state, tape: np.ndarray, precision: str = "d"):
    """LossyNetworksGeneralizedCliffordsSimulationStrategy.simulate
    (lossy_networks_generalized_cliffords_simulation_strategy.py:41-83)."""
    m = len(input_state)
    big = expanded_matrix(U_lossy)
    s2 = list(input_state) + [0] * m
    return [x[:m] for x in gccb_simulate(big, s2, tape, precision=precision)]


def gcc_weights(input_state: Sequence[int], substates: Sequence[Sequence[int]]) -> np.ndarray:
    """generalized_cliffords_simulation_strategy.py:173-205 (weights, normalised at :155)."""
    from scipy.special import binom

    n = int(sum(input_state))
    out = []
    for st in substates:
        kv = [int(input_state[i]) - int(st[i]) for i in range(len(st))]
        l = sum(kv)
        w = factorial(l) * factorial(n - l) / factorial(n)
        for mm in range(len(input_state)):
            w *= binom(input_state[mm], kv[mm])
        out.append(w)
    out = np.array(out)
    return out / sum(out)


def substates_by_size(input_state: Sequence[int]):
    """generalized_cliffords_simulation_strategy.py:75-121, same enumeration order."""
    def rec(part):
        if len(part) < 1:
            return [[]]
        smaller = rec(part[1:])
        return [[i] + sub for i in range(int(part[0]) + 1) for sub in smaller]

    lab = {}
    for st in rec(list(input_state)):
        lab.setdefault(sum(st), []).append(tuple(st))
    return lab


def gcc_layer_pmf(U, input_state, r_sample, calculator: str = "glynn", precision: str = "d") -> np.ndarray:
    """_calculate_new_layer_of_pmfs (generalized_cliffords_simulation_strategy.py:136-171) with
    _calculate_outputs_probability (:224-247)."""
    m = len(input_state)
    k = int(sum(r_sample)) + 1
    subs = substates_by_size(input_state)[k]
    w = gcc_weights(input_state, subs)
    pmf = []
    for j in range(m):
        out = list(r_sample)
        out[j] += 1
        acc = 0
        for i, st in enumerate(subs):
            if calculator == "glynn":
                p = glynn(U, st, out, precision)
            else:
                p = guan_permanent(U, st, out, RYSER if calculator == "ryser" else CHIN_HUH, precision)
            prob = abs(p) ** 2
            for occ in st:
                prob /= factorial(occ)
            prob /= factorial(sum(st))
            acc += prob * w[i]
        pmf.append(acc)
    return np.array(pmf)


def gcc_simulate(U, input_state, uniforms: np.ndarray, calculator: str = "glynn", precision: str = "d"):
    """GeneralizedCliffordsSimulationStrategy.simulate (:49-73, :123-134, :249-266); one uniform
    per particle: uniforms[sample, k]; pmfs memoised per partial output like the reference."""
    n = int(sum(input_state))
    m = len(input_state)
    memo = {}
    samples = []
    for i in range(uniforms.shape[0]):
        r = [0] * m
        for k in range(n):
            key = tuple(r)
            if key not in memo:
                memo[key] = gcc_layer_pmf(U, input_state, r, calculator, precision)
            r[gcc_draw(memo[key], uniforms[i, k])] += 1
        samples.append(tuple(r))
    return samples
