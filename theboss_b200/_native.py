"""ctypes binding of ``theboss_b200/lib/libbossperm.so`` (C ABI: include/bossperm.h).

There is no CPU fallback and no alternative backend: if the library is missing, or no sm_100 CUDA
device is visible, every compute entry point raises ``BossPermError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BOSSPERM_LIB: another build of the same library (A/B measurements of compile-time tuning knobs); never a different backend
LIB_PATH = os.environ.get("BOSSPERM_LIB") or os.path.join(_HERE, "lib", "libbossperm.so")

BP_MAX_N = 40
BP_MAX_MODES = 256
FORMULA_RYSER, FORMULA_CHIN_HUH, FORMULA_GLYNN = 0, 1, 2

BP_OK, BP_ERR_INVALID, BP_ERR_SHAPE, BP_ERR_UNSUPPORTED, BP_ERR_CUDA, BP_ERR_NOMEM, BP_ERR_DOMAIN = 0, -1, -2, -3, -4, -5, -6


class BossPermError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbossperm error {code}: {message}")
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); mirrors include/bossperm.h one to one.
SIGNATURES = {
    "bp_abi_version": (C.c_int, []),
    "bp_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bp_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bp_destroy": (C.c_int, [C.c_void_p]),
    "bp_last_error": (C.c_char_p, [C.c_void_p]),
    "bp_synchronize": (C.c_int, [C.c_void_p]),
    "bp_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bp_launch_count": (C.c_int64, [C.c_void_p]),
    "bp_timer_start": (C.c_int, [C.c_void_p]),
    "bp_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "bp_fp64_peak": (C.c_int, [C.c_void_p, C.c_double, _dp]),
    "bp_glynn_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _dp]),
    "bp_glynn_matrix_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, _dp]),
    "bp_glynn_matrix_range_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bp_glynn_set_resident": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bp_exchange_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "bp_exchange_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bp_exchange_destroy": (C.c_int, [C.c_void_p]),
    "bp_glynn_matrix_range_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bp_glynn_matrix_range_exchange_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bp_glynn_single": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _dp]),
    "bp_perm_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "bp_perm_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "bp_minors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bp_gccb_pmf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bp_gccb_simulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]),
    "bp_gccb_simulate_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "bp_gccb_simulate_bobs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_uint64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "bp_bobs_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load libbossperm.so and attach the prototypes.  Raises if the library was not built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise BossPermError(
                    BP_ERR_CUDA,
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(make -C theboss_b200/csrc).  There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def as_matrix(matrix) -> np.ndarray:
    """ndarray or list of lists -> C-contiguous complex128 copy-if-needed (re-read on every call: the
    reference lets callers mutate ``calculator.matrix`` in place,
    tests/gcc_based_strategies_tests_base.py:89-92 in the reference)."""
    a = np.ascontiguousarray(np.asarray(matrix, dtype=np.complex128))
    if a.ndim != 2:
        raise AttributeError
    return a


def as_state(state, m: Optional[int] = None) -> np.ndarray:
    """list / tuple / ndarray of ints or integer-valued floats -> int32, zero-padded to m."""
    a = np.asarray(state)
    if a.ndim != 1:
        a = a.reshape(-1)
    out = np.zeros(len(a) if m is None else m, dtype=np.int32)
    out[: len(a)] = a.astype(np.int64)
    return out


def occupations_u8(x) -> np.ndarray:
    """Occupation table -> C-contiguous uint8; counts outside 0 .. 255 raise instead of wrapping (the int32 entry points
    check the same on the C side)."""
    a = np.asarray(x)
    if a.dtype != np.uint8:
        if a.size and (a.min() < 0 or a.max() > 255):
            raise ValueError("occupation numbers must lie in 0 .. 255")
        a = a.astype(np.uint8)
    return np.ascontiguousarray(a)


class Handle:
    """One device + one stream + scratch (bp_handle).  Created lazily, never pickled."""

    def __init__(self, device: int = 0, stream_ptr: Optional[int] = None):
        self._lib = load_library()
        self._h = C.c_void_p()
        if stream_ptr is None:
            rc = self._lib.bp_create(int(device), C.byref(self._h))
        else:
            rc = self._lib.bp_create_on_stream(int(device), C.c_void_p(stream_ptr), C.byref(self._h))
        if rc != BP_OK:
            raise BossPermError(rc, self._lib.bp_last_error(None).decode())
        self.device = int(device)
        # The handle is NOT thread-safe on the C side (one pinned staging buffer, one set of scratch slots, one K1 arrival
        # counter).  ctypes releases the GIL during a call, so two Python threads that share the process-wide default handle --
        # independent calculator objects, safe in the reference -- would otherwise run bp_* concurrently on it: every call goes
        # through _call, which serialises the callers of one handle.
        self._lock = threading.RLock()

    # -- plumbing ----------------------------------------------------------------------------
    def _call(self, name: str, *args):
        """One C-ABI call on this handle under the handle's lock; the error text is read before the lock is released."""
        with self._lock:
            rc = getattr(self._lib, name)(self._h, *args)
            if rc == BP_OK:
                return
            msg = self._lib.bp_last_error(self._h).decode()
        if rc == BP_ERR_SHAPE:
            raise AttributeError(msg)   # bs_permanent_calculator_base.py:179-180
        if rc == BP_ERR_DOMAIN:
            raise ValueError(msg)       # numpy.random.choice: "probabilities do not sum to 1" (generalized_cliffords_b_simulation_strategy.py:107-110)
        raise BossPermError(rc, msg)

    @staticmethod
    def _normalised(U, s, t):
        """The C ABI reads raw buffers: C-contiguous complex128 (m, m) and int32[m] occupations, whatever the
        caller handed in (a no-op for arrays that already have that form)."""
        U = as_matrix(U)
        if U.shape[0] != U.shape[1]:
            raise AttributeError
        m = U.shape[0]

        def state(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray) and x.dtype == np.int32 and x.shape == (m,) and x.flags.c_contiguous:
                return x
            if np.size(x) > m:
                raise AttributeError
            return as_state(x, m)
        return U, state(s), state(t)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.bp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._call("bp_synchronize")

    def device_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._call("bp_device_info", C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"sm_count": a.value, "cc": (b.value, c.value), "clock_khz": d.value}

    def launch_count(self) -> int:
        return int(self._lib.bp_launch_count(self._h))

    def timer_start(self):
        self._call("bp_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._call("bp_timer_stop", C.byref(ms))
        return float(ms.value)

    def fp64_peak(self, target_ms: float = 200.0) -> float:
        tf = C.c_double()
        self._call("bp_fp64_peak", float(target_ms), C.byref(tf))
        return float(tf.value)

    # -- K1 ----------------------------------------------------------------------------------
    def glynn_matrix(self, A: np.ndarray) -> complex:
        A = as_matrix(A)
        if A.shape[0] != A.shape[1]:
            raise AttributeError
        out = (C.c_double * 2)()
        self._call("bp_glynn_matrix", A.ctypes.data, A.shape[0], out)
        return complex(out[0], out[1])

    def glynn_matrix_range(self, A: np.ndarray, lo: int, hi: int):
        """Un-normalised double-double partial (re_hi, re_lo, im_hi, im_lo) over Gray steps [lo, hi)."""
        A = as_matrix(A)
        if A.shape[0] != A.shape[1]:
            raise AttributeError("glynn_matrix_range needs a square matrix")
        if not 0 <= int(lo) <= int(hi):
            raise ValueError(f"Gray-step range [{lo}, {hi}) is not ordered")
        out = (C.c_double * 4)()
        self._call("bp_glynn_matrix_range", A.ctypes.data, A.shape[0], int(lo), int(hi), out)
        return tuple(out)

    def glynn_matrix_range_dev(self, dA_ptr: int, N: int, lo: int, hi: int, d_out_ptr: int):
        self._call("bp_glynn_matrix_range_dev", C.c_void_p(dA_ptr), int(N), int(lo), int(hi), C.c_void_p(d_out_ptr))

    def glynn_set_resident(self, dA_ptr: Optional[int]):
        """Declare a device matrix resident (None: no resident matrix); see include/bossperm.h."""
        self._call("bp_glynn_set_resident", C.c_void_p(dA_ptr) if dA_ptr else None)

    def exchange_create(self, world: int, rank: int) -> bytes:
        """Allocate this rank's slot buffer of the peer-memory partial exchange; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_ubyte * 64)()
        self._call("bp_exchange_create", int(world), int(rank), buf)
        return bytes(buf)

    def exchange_connect(self, ipc_handles) -> None:
        """Map the peers' slot buffers: `ipc_handles` = the handles of all ranks in rank order (64 bytes each)."""
        blob = b"".join(bytes(x) for x in ipc_handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._call("bp_exchange_connect", buf)

    def exchange_destroy(self) -> None:
        self._call("bp_exchange_destroy")

    def glynn_matrix_range_exchange(self, dA_ptr: int, N: int, lo: int, hi: int, d_out_all_ptr: int):
        self._call("bp_glynn_matrix_range_exchange", C.c_void_p(dA_ptr), int(N), int(lo), int(hi), C.c_void_p(d_out_all_ptr))

    def glynn_matrix_range_exchange_host(self, A: np.ndarray, lo: int, hi: int, world: int) -> np.ndarray:
        """Host-buffer form of the collective: this rank's slice [lo, hi) of the Gray range of A plus the partial exchange, one call;
        returns all ranks' un-normalised double-double partials, shape (world, 4)."""
        A = np.ascontiguousarray(A, dtype=np.complex128)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise AttributeError
        if not 0 <= int(lo) <= int(hi):
            raise ValueError(f"Gray step range [{lo}, {hi})")
        out = np.empty((int(world), 4), dtype=np.float64)
        self._call("bp_glynn_matrix_range_exchange_host", C.c_void_p(A.ctypes.data), int(A.shape[0]), int(lo), int(hi), C.c_void_p(out.ctypes.data))
        return out

    def glynn_single(self, U: np.ndarray, s: np.ndarray, t: np.ndarray) -> complex:
        U, s, t = self._normalised(U, s, t)
        out = (C.c_double * 2)()
        self._call("bp_glynn_single", U.ctypes.data, U.shape[0], s.ctypes.data, t.ctypes.data, out)
        return complex(out[0], out[1])

    # -- K2 ----------------------------------------------------------------------------------
    def perm_batched(self, U: np.ndarray, S: np.ndarray, T: np.ndarray, formula: int = FORMULA_GLYNN) -> np.ndarray:
        U = as_matrix(U)
        S, T = occupations_u8(S), occupations_u8(T)
        if S.shape != T.shape or S.ndim != 2 or S.shape[1] != U.shape[0]:
            raise AttributeError
        out = np.zeros(S.shape[0], dtype=np.complex128)
        self._call("bp_perm_batched", U.ctypes.data, U.shape[0], S.ctypes.data, T.ctypes.data,
                                              S.shape[0], int(formula), out.ctypes.data)
        return out

    def perm_batched_dev(self, dU: int, m: int, dS: int, dT: int, B: int, formula: int, d_out: int):
        self._call("bp_perm_batched_dev", C.c_void_p(dU), int(m), C.c_void_p(dS), C.c_void_p(dT),
                                                  int(B), int(formula), C.c_void_p(d_out))

    # -- K3 ----------------------------------------------------------------------------------
    def minors(self, U: np.ndarray, s: np.ndarray, t: np.ndarray, formula: int = FORMULA_CHIN_HUH) -> np.ndarray:
        U, s, t = self._normalised(U, s, t)
        out = np.zeros(U.shape[0], dtype=np.complex128)
        self._call("bp_minors", U.ctypes.data, U.shape[0], s.ctypes.data, t.ctypes.data, int(formula),
                                        out.ctypes.data)
        return out

    def gccb_pmf(self, U: np.ndarray, s: np.ndarray, t: np.ndarray, want_minors: bool = False):
        U, s, t = self._normalised(U, s, t)
        m = U.shape[0]
        pmf = np.zeros(m, dtype=np.float64)
        minors = np.zeros(m, dtype=np.complex128) if want_minors else None
        self._call("bp_gccb_pmf", U.ctypes.data, m, s.ctypes.data, t.ctypes.data, pmf.ctypes.data,
                                          minors.ctypes.data if want_minors else None)
        return (pmf, minors) if want_minors else pmf

    # -- K3 + K4 -----------------------------------------------------------------------------
    def gccb_simulate(self, U: np.ndarray, s: np.ndarray, n_samples: int, eta: float = -1.0, seed: int = 0,
                      first_sample: int = 0, tape: Optional[np.ndarray] = None) -> np.ndarray:
        U, s, _ = self._normalised(U, s, None)
        m = U.shape[0]
        out = np.zeros((int(n_samples), m), dtype=np.int32)
        tp = None
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.float64)
            n = int(s.sum())
            if tape.shape != (int(n_samples), 1 + 2 * n):
                raise ValueError(f"decision tape must have shape ({n_samples}, {1 + 2 * n})")
            tp = tape.ctypes.data
        self._call("bp_gccb_simulate", U.ctypes.data, m, s.ctypes.data, int(n_samples), float(eta),
                                               int(seed) & (2 ** 64 - 1), int(first_sample), tp, out.ctypes.data)
        return out

    def gccb_simulate_batch(self, Us: np.ndarray, states: np.ndarray, seed: int = 0, first_sample: int = 0,
                            tape: Optional[np.ndarray] = None) -> np.ndarray:
        """One GCC-B sample per (matrix, input state) pair: Us (S, m, m) complex128, states (S, m) ints."""
        Us = np.ascontiguousarray(Us, dtype=np.complex128)
        states = np.ascontiguousarray(states, dtype=np.int32)
        if Us.ndim != 3 or Us.shape[1] != Us.shape[2] or states.shape != Us.shape[:2]:
            raise AttributeError("Us must be (S, m, m) and states (S, m)")
        S, m = states.shape
        out = np.zeros((S, m), dtype=np.int32)
        tp, tn = None, 0
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.float64)
            tn = (tape.shape[1] - 1) // 2
            if tape.shape[0] != S or tape.shape[1] != 1 + 2 * tn or tn < int(states.sum(axis=1).max(initial=0)):
                raise ValueError("decision tape must have shape (S, 1 + 2 * n_max)")
            tp = tape.ctypes.data
        self._call("bp_gccb_simulate_batch", Us.ctypes.data, m, states.ctypes.data, S, int(seed) & (2 ** 64 - 1),
                                                     int(first_sample), tp, tn, out.ctypes.data)
        return out

    def gccb_simulate_bobs(self, B, qft, phases, perms, states, seed: int = 0, first_sample: int = 0,
                           tape: Optional[np.ndarray] = None) -> np.ndarray:
        """One GCC-B sample per row of ``states`` on the matrix (B[:, perms[i]]) @ diag(phases[i], 1...) @ QFT_a built on the
        device (what the BOBS strategies construct per sample; see include/bossperm.h)."""
        B, qft, a, phases, perms, S = _bobs_operands(B, qft, phases, perms)
        m = B.shape[0]
        states = np.ascontiguousarray(states, dtype=np.int32)
        if states.shape != (S, m):
            raise AttributeError("states must be (S, m)")
        out = np.zeros((S, m), dtype=np.int32)
        tp, tn = None, 0
        if tape is not None:
            tape = np.ascontiguousarray(tape, dtype=np.float64)
            tn = (tape.shape[1] - 1) // 2
            if tape.shape[0] != S or tape.shape[1] != 1 + 2 * tn or tn < int(states.sum(axis=1).max(initial=0)):
                raise ValueError("decision tape must have shape (S, 1 + 2 * n_max)")
            tp = tape.ctypes.data
        self._call("bp_gccb_simulate_bobs", B.ctypes.data, m, qft.ctypes.data if a else None, a, phases.ctypes.data if a else None,
                   perms.ctypes.data if perms is not None else None, states.ctypes.data, S, int(seed) & (2 ** 64 - 1),
                   int(first_sample), tp, tn, out.ctypes.data)
        return out

    def bobs_build(self, B, qft, phases, perms=None) -> np.ndarray:
        """The per-sample matrices of gccb_simulate_bobs themselves, (S, m, m) complex128, built on the device."""
        B, qft, a, phases, perms, S = _bobs_operands(B, qft, phases, perms)
        m = B.shape[0]
        out = np.zeros((S, m, m), dtype=np.complex128)
        self._call("bp_bobs_build", B.ctypes.data, m, qft.ctypes.data if a else None, a, phases.ctypes.data if a else None,
                   perms.ctypes.data if perms is not None else None, S, out.ctypes.data)
        return out


def _bobs_operands(B, qft, phases, perms):
    """Normalised operands of the device-side BOBS matrix build (include/bossperm.h): (B, qft, a, phases, perms, S)."""
    B = as_matrix(B)
    if B.shape[0] != B.shape[1]:
        raise AttributeError("B must be square")
    m = B.shape[0]
    phases = np.ascontiguousarray(phases, dtype=np.complex128)
    if phases.ndim != 2:
        raise AttributeError("phases must be (S, a)")
    S, a = phases.shape
    qft = np.ascontiguousarray(qft, dtype=np.complex128)
    if qft.shape != (a, a) or a > m:
        raise AttributeError("qft must be (a, a) with a <= m")
    if perms is not None:
        perms = np.ascontiguousarray(perms, dtype=np.int32)
        if perms.shape != (S, m) or (S and (perms.min() < 0 or perms.max() >= m)):
            raise AttributeError("perms must be (S, m) column indices")
    return B, qft, a, phases, perms, S


_default_handles = {}
_default_lock = threading.Lock()


def default_handle(device: int = 0) -> Handle:
    """Process-wide handle per device, created on first use (after fork/spawn/unpickle too)."""
    key = (os.getpid(), int(device))
    with _default_lock:
        h = _default_handles.get(key)
        if h is None:
            h = Handle(device)
            _default_handles[key] = h
        return h
