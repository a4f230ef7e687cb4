"""tools/jacobi_t_ab.py — whole-step time of the randomized SVD (BASELINE configs[1] shape) with the one-sided Jacobi kernel
applied to Rhat (default) or to Rhat^T (option jacobi_transpose, Drmac-Veselic: the transposed triangular factor converges in
fewer sweeps); prints the sweep counts (verbose) and the device time per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

m, n, k, p = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (50000, 20000, 500, 20)
lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
st = D.stream()
g = torch.Generator(device="cuda").manual_seed(1234)
r = 640 if k + p <= 640 else 1280
X = torch.randn((m, r), dtype=torch.float64, device="cuda", generator=g) / m ** 0.5
W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
sig = torch.logspace(1, -3, r, dtype=torch.float64, device="cuda")
A = torch.empty((n, m), dtype=torch.float64, device="cuda")
for j0 in range(0, n, 4096):
    j1 = min(n, j0 + 4096)
    torch.matmul(W[j0:j1] * sig, X.t(), out=A[j0:j1])
    A[j0:j1] += 1e-6 * torch.randn((j1 - j0, m), dtype=torch.float64, device="cuda", generator=g)
del X, W
U = D.new_cm(m, k); V = D.new_cm(n, k); S = torch.empty(k, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
res = {}
for jt in (0, 1, 0, 1):
    lib.rsvd_b200_set_option(b"jacobi_transpose", jt)
    lib.rsvd_b200_set_option(b"verbose", 1)
    native.check(lib.rsvd_b200_svd_rand_dev(A.data_ptr(), m, n, m, k, p, 1, 2, 1, 777, None, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n))
    lib.rsvd_b200_set_option(b"verbose", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
    for _ in range(3):
        native.check(lib.rsvd_b200_svd_rand_dev(A.data_ptr(), m, n, m, k, p, 1, 2, 1, 777, None, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n))
    with torch.cuda.stream(st):
        e1.record()
    lib.rsvd_b200_sync()
    pe = lib.rsvd_b200_svd_percent_error_dev(A.data_ptr(), m, n, m, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n, k)
    orth = (U @ U.t() - torch.eye(k, dtype=torch.float64, device="cuda")).abs().max().item()
    res[jt] = S.clone()
    print("jacobi_transpose=%d: %.3f ms per step, percent error %.6f, ||UtU-I||max %.1e" % (jt, e0.elapsed_time(e1) / 3, pe, orth), flush=True)
print("max rel sigma difference between the two: %.2e" % ((res[0] - res[1]).abs() / res[0]).max().item())
lib.rsvd_b200_set_option(b"jacobi_transpose", 0)
