"""Small invocation of every kernel and launch path for compute-sanitizer (scripts/sanitize.sh): K1 generic / block-4 / fused finish /
in-kernel exchange, K2 host-prep and device-prep paths, K3 in one-warp and 128-thread blocks with one, two and four lanes per term
stream, the sampling loop (lossless, uniform losses, per-sample matrices built on the device)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads
from theboss_b200 import _native
from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import generate_qft_matrix_for_first_m_modes

h = _native.default_handle(0)
rng = np.random.RandomState(1)


def occ(m, n):
    out = np.zeros(m, dtype=np.int32)
    for j in rng.randint(0, m, n):
        out[j] += 1
    return out


print("K1", h.glynn_matrix(workloads.c4_matrix(12)), h.glynn_matrix(workloads.c4_matrix(23)))
A = workloads.c4_matrix(24)
print("K1 range", h.glynn_matrix_range(A, 100, 70000)[:2])
print("K1 wide", h.glynn_matrix_range(workloads.c4_matrix(36), 0, 1 << 17)[:2], h.glynn_matrix_range(workloads.c4_matrix(39), 64, (1 << 16) + 64)[:2])
dA = torch.from_numpy(np.ascontiguousarray(A).view(np.float64).reshape(-1).copy()).cuda()
d_all = torch.zeros(4, dtype=torch.float64, device="cuda")
hx = _native.Handle(0)
hx.exchange_connect([hx.exchange_create(1, 0)])
hx.glynn_set_resident(dA.data_ptr())
for lo, hi in ((0, 1 << 17), (64, 1 << 17), (3, 900)):
    hx.glynn_matrix_range_exchange(dA.data_ptr(), 24, lo, hi, d_all.data_ptr())
hx.synchronize()
print("K1 exchange", d_all.cpu().numpy()[:2])
hx.close()
U = workloads.haar(14, 14)
for B in (5, 300):
    S = np.array([occ(14, b % 9) for b in range(B)], dtype=np.uint8)
    T = np.array([occ(14, b % 9) for b in range(B)], dtype=np.uint8)
    print("K2", B, np.abs(h.perm_batched(U, S, T)).sum())
for k, m, bunch in ((3, 8, False), (10, 20, False), (14, 28, False), (18, 36, True), (31, 12, True)):
    U = workloads.haar(m, k)
    s = occ(m, k) if bunch else np.array([1] * k + [0] * (m - k), dtype=np.int32)
    t = np.zeros(m, dtype=np.int32)
    if bunch:
        t[: 3] = [(k - 1) // 3, (k - 1) // 3, (k - 1) - 2 * ((k - 1) // 3)]
    else:
        t = occ(m, k - 1)
    pmf, minors = h.gccb_pmf(U, s, t, want_minors=True)
    print("K3", k, np.abs(minors).max(), pmf.sum())
U = workloads.haar(12, 5)
s = np.array([1] * 6 + [0] * 6, dtype=np.int32)
print("K4", h.gccb_simulate(U, s, 40, seed=3).sum(), h.gccb_simulate(U, s, 40, eta=0.6, seed=3).sum())
qft = generate_qft_matrix_for_first_m_modes(5, 12)[:5, :5]
phases = np.exp(2j * np.pi * rng.random_sample((9, 5)))
perms = np.argsort(rng.random_sample((9, 12)), axis=1).astype(np.int32)
states = np.array([occ(12, i % 5) for i in range(9)], dtype=np.int32)
print("BOBS", h.gccb_simulate_bobs(U, qft, phases, perms, states, seed=4).sum(), np.abs(h.bobs_build(U, qft, phases, None)).sum())
print("sanitize job done; kernels launched:", h.launch_count())
