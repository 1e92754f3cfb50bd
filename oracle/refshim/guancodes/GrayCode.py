"""TEST INFRASTRUCTURE ONLY -- stand-in for ``guancodes.GrayCode`` (PyPI ``guancodes~=0.0.3``, not
vendored under /root/reference).

Sole reference call site: theboss/boson_sampling_utilities/permanent_calculators/
glynn_gray_permanent_calculator.py:57 (consumed at :60-67).  Contract inferred from that use: return
the 2^k - 1 positions in [0, k) whose successive sign flips visit every pattern of k signs exactly
once.  The canonical binary-reflected Gray code flips bit ctz(t) at step t = 1 .. 2^k - 1.
"""


def get_gray_code_update_indices(k: int):
    return [(t & -t).bit_length() - 1 for t in range(1, 1 << k)]
