"""Seeded synthetic workloads of BASELINE.json's configs, shared by the tests,
bench.py and the golden generators.  Pure NumPy/SciPy; no reference or oracle imports."""
import numpy as np
from scipy.stats import unitary_group


def haar(m: int, seed: int) -> np.ndarray:
    return np.ascontiguousarray(unitary_group.rvs(m, random_state=seed).astype(np.complex128))


def c4_matrix(n: int = 30) -> np.ndarray:
    """C4: U = Haar(2n, seed n); A = U[rows, :n] with n distinct rows from RandomState(n)."""
    U = haar(2 * n, n)
    rows = np.sort(np.random.RandomState(n).choice(2 * n, n, replace=False))
    return np.ascontiguousarray(U[rows, :n])


def c2_batch(n: int = 20, m: int = 40, items: int = 10_000):
    """C2: U = Haar(m, seed n); item i: s, t ~ multinomial(n, uniform over m modes) from
    RandomState(1000 + i) (repeated rows and columns)."""
    U = haar(m, n)
    S = np.zeros((items, m), dtype=np.uint8)
    T = np.zeros((items, m), dtype=np.uint8)
    p = np.full(m, 1.0 / m)
    for i in range(items):
        rng = np.random.RandomState(1000 + i)
        S[i] = rng.multinomial(n, p)
        T[i] = rng.multinomial(n, p)
    return U, S, T


def c3_step(n: int = 24, m: int = 48, collision_free: bool = False):
    """C3: U = Haar(m, seed n); s = [1]*n + [0]*(m-n); t = n-1 particles placed by
    RandomState(n).randint(0, m, n-1) (realistic bunching) or n-1 distinct modes."""
    U = haar(m, n)
    s = np.array([1] * n + [0] * (m - n), dtype=np.int32)
    t = np.zeros(m, dtype=np.int32)
    rng = np.random.RandomState(n)
    if collision_free:
        t[rng.choice(m, n - 1, replace=False)] = 1
    else:
        for j in rng.randint(0, m, n - 1):
            t[j] += 1
    return U, s, t


def c5_lossy(n: int = 30, m: int = 60):
    """C5: U = Haar(m, seed m); (i) lossless U + eta = 0.5; (ii) U @ diag(sqrt(linspace(.3,.9,m)))."""
    U = haar(m, m)
    s = np.array([1] * n + [0] * (m - n), dtype=np.int32)
    eta = np.linspace(0.3, 0.9, m)
    return U, np.ascontiguousarray(U @ np.diag(np.sqrt(eta))), s
