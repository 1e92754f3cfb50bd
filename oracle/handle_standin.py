"""Stand-in for ``theboss_b200._native.Handle`` in CPU tests of the HOST-side logic (test infrastructure only).

The product package has no CPU path: without a CUDA device every compute call raises.  To exercise the Python layer above
the C ABI (argument handling, memoisation, RNG call order, batching, bookkeeping) in the GPU-less build container, tests
monkey-patch ``_native.default_handle`` with this object, which answers the same method calls from the CPU oracle
(oracle/pyoracle.py).  Error behaviour mirrors the C ABI: shape mismatches raise ``AttributeError`` like ``BP_ERR_SHAPE``.
"""
import numpy as np

from oracle import pyoracle as orc


def _matrix(U) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(U, dtype=np.complex128))
    if a.ndim != 2:
        raise AttributeError
    return a


def _state(x, m: int) -> np.ndarray:
    a = np.asarray(x).reshape(-1)
    if len(a) > m:
        raise AttributeError
    out = np.zeros(m, dtype=np.int32)
    out[: len(a)] = a.astype(np.int64)
    return out


def _keyed_tape(seed, first_sample, n_samples: int, width: int) -> np.ndarray:
    """Uniforms keyed by (seed, global sample index) like the device's counter-based generator, so that results do not
    depend on how a job is cut into calls (different numbers than Philox, same distribution)."""
    tape = np.zeros((n_samples, width))
    for i in range(n_samples):
        key = (int(seed) * 0x9E3779B1 + int(first_sample) + i) % 2 ** 32
        tape[i] = np.random.RandomState(key).random_sample(width)
    return tape


class OracleHandle:
    device = 0

    def __init__(self, precision: str = "d"):
        self._precision = precision
        self.calls = 0

    # -- K1 ------------------------------------------------------------------------------------------
    def glynn_matrix(self, A):
        self.calls += 1
        A = _matrix(A)
        return complex(1) if A.shape[0] == 0 else orc.glynn_matrix(A, self._precision)

    def glynn_matrix_range(self, A, lo, hi):
        """Un-normalised partial over Gray steps [lo, hi) as (re_hi, re_lo, im_hi, im_lo)."""
        self.calls += 1
        part = orc.glynn_range(_matrix(A), int(lo), int(hi), "ld") if hi > lo else 0j
        return (part.real, 0.0, part.imag, 0.0)

    def glynn_single(self, U, s, t):
        self.calls += 1
        U = _matrix(U)
        s, t = _state(s, U.shape[0]), _state(t, U.shape[0])
        if s.sum() == 0 or t.sum() == 0:
            return complex(1)                       # bp_glynn_single: empty effective matrix
        if s.sum() != t.sum():
            raise AttributeError("input and output particle numbers differ")
        return orc.glynn(U, s, t, self._precision)

    # -- K2 ------------------------------------------------------------------------------------------
    def perm_batched(self, U, S, T, formula=None):
        self.calls += 1
        U = _matrix(U)
        S, T = np.asarray(S), np.asarray(T)
        if S.shape != T.shape or S.ndim != 2 or S.shape[1] != U.shape[0]:
            raise AttributeError
        if np.any(S.sum(axis=1) != T.sum(axis=1)):
            raise AttributeError("input and output particle numbers differ")
        return np.array([orc.guan_permanent(U, S[b].astype(np.int32), T[b].astype(np.int32), orc.CHIN_HUH, self._precision)
                         for b in range(S.shape[0])], dtype=np.complex128)

    # -- K3 ------------------------------------------------------------------------------------------
    def minors(self, U, s, t, formula=None):
        self.calls += 1
        U = _matrix(U)
        s, t = _state(s, U.shape[0]), _state(t, U.shape[0])
        if s.sum() < 1 or t.sum() != s.sum() - 1:
            raise AttributeError("sum(t) must equal sum(s) - 1")
        return orc.submatrices(U, s, t, orc.CHIN_HUH, self._precision)

    def gccb_pmf(self, U, s, t, want_minors=False):
        self.calls += 1
        pmf = orc.gccb_pmf(U, s, t, self._precision)
        return (pmf, self.minors(U, s, t)) if want_minors else pmf

    # -- K3 + K4 -------------------------------------------------------------------------------------
    def gccb_simulate(self, U, s, n_samples, eta=-1.0, seed=0, first_sample=0, tape=None):
        """Decisions from the given tape, else from uniforms keyed by (seed, sample index)."""
        self.calls += 1
        U = _matrix(U)
        s = _state(s, U.shape[0])
        n = int(s.sum())
        if tape is None:
            tape = _keyed_tape(seed, first_sample, int(n_samples), 1 + 2 * n)
        if n == 0 or n_samples == 0:
            return np.zeros((int(n_samples), U.shape[0]), dtype=np.int32)
        if eta >= 0:
            out = orc.gccb_uniform_losses_simulate(U, s, eta, tape, self._precision)
        else:
            out = orc.gccb_simulate(U, s, tape, precision=self._precision)
        return np.array(out, dtype=np.int32).reshape(int(n_samples), U.shape[0])

    def gccb_simulate_bobs(self, B, qft, phases, perms, states, seed=0, first_sample=0, tape=None):
        return self.gccb_simulate_batch(bobs_matrices(B, qft, phases, perms), states, seed=seed, first_sample=first_sample, tape=tape)

    def bobs_build(self, B, qft, phases, perms=None):
        return bobs_matrices(B, qft, phases, perms)

    def gccb_simulate_batch(self, Us, states, seed=0, first_sample=0, tape=None):
        """One GCC-B sample per (matrix, input state) pair through the oracle's sampling loop."""
        self.calls += 1
        states = np.asarray(states)
        out = np.zeros(states.shape, dtype=np.int32)
        if tape is None:
            tape = _keyed_tape(seed, first_sample, states.shape[0], 1 + 2 * int(states.sum(axis=1).max(initial=0)))
        for i in range(states.shape[0]):
            n = int(states[i].sum())
            if n:
                row = tape[i:i + 1, : 1 + 2 * n]
                out[i] = orc.gccb_simulate(Us[i], states[i], row, precision=self._precision)[0]
        return out


def bobs_matrices(B, qft, phases, perms=None):
    """NumPy restatement of the device-side BOBS matrix build (include/bossperm.h, bp_bobs_build):
    Us[i] = (B[:, perms[i]]) @ diag(phases[i], 1 ...) @ (QFT on the first a modes)
    -- nonuniform_losses_approximation_strategy.py:331-347, lossy_state_approximated_simulation_strategy.py:329-362."""
    B = np.asarray(B, dtype=np.complex128)
    phases = np.asarray(phases, dtype=np.complex128)
    S, a = phases.shape
    m = B.shape[0]
    if perms is None:
        Us = np.repeat(B[None, :, :], S, axis=0)
    else:
        Us = np.ascontiguousarray(np.transpose(B.T[np.asarray(perms)], (0, 2, 1)))
    if a > 0:
        Us[:, :, :a] = (Us[:, :, :a] * phases[:, None, :]) @ np.asarray(qft, dtype=np.complex128)
    return Us


def install(monkeypatch, precision: str = "d") -> OracleHandle:
    """Route ``_native.default_handle`` to one OracleHandle for the duration of a test."""
    from theboss_b200 import _native
    handle = OracleHandle(precision)
    monkeypatch.setattr(_native, "default_handle", lambda device=0: handle)
    return handle
