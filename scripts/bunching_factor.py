"""How much do bunched outputs shrink the K3 walk?  f = prod(t_j + 1) / 2^(k-1) for occupations t drawn uniformly over
the Fock states of k - 1 particles in m modes (the Haar-averaged output law of boson sampling), and the share of the total
work carried by the samples above a threshold.  CPU only; backs the decision recorded in DESIGN.md section 9 not to add a
second (input-side) walk for nearly collision-free outputs.

    python scripts/bunching_factor.py
"""
import numpy as np


def uniform_fock_states(m, particles, size, rng):
    """Uniform multisets by stars and bars: `particles` bar positions among m + particles - 1 slots."""
    out = np.zeros((size, m), dtype=np.int64)
    for i in range(size):
        pos = np.sort(rng.choice(m + particles - 1, particles, replace=False))
        np.add.at(out[i], pos - np.arange(particles), 1)
    return out


def main():
    rng = np.random.default_rng(1)
    print("  m  particles   mean f  median f   share of samples / of work with f > 0.87, 0.71, 0.5")
    for m, particles in [(48, 23), (40, 19), (120, 29), (120, 24), (60, 29)]:
        t = uniform_fock_states(m, particles, 20000, rng)
        f = 2.0 ** (np.sum(np.log2(t + 1), axis=1) - particles)
        work = f / f.sum()
        cells = "   ".join(f"{(f > thr).mean():.4f} / {work[f > thr].sum():.4f}" for thr in (0.87, 0.71, 0.5))
        print(f"{m:4d} {particles:9d} {f.mean():9.3f} {np.median(f):9.3f}   {cells}")


if __name__ == "__main__":
    main()
