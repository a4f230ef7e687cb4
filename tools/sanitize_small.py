"""tools/sanitize_small.py — small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck|racecheck|synccheck`:
dataflow Cholesky + inverse, Gram-based Jacobi with the live replay, the read-modify-write GEMM epilogue (blocked QB), shifted
CholeskyQR3."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
rng = np.random.default_rng(0)
for n in (40, 200):
    Y = rng.standard_normal((2 * n, n))
    Gd, Xd = D.from_numpy_cm(Y.T @ Y), D.new_cm(n, n)
    assert lib.rsvd_b200_chol_inv(D.ptr(Gd), n, n, D.ptr(Xd), n, None) == 0
    lib.rsvd_b200_sync()
for n in (20, 150):
    A = rng.standard_normal((n, n))
    Ad = D.from_numpy_cm(A)
    U = torch.empty((n, n), dtype=torch.float64, device="cuda"); Vt = torch.empty_like(U); s = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
    lib.rsvd_b200_sync()
if os.environ.get("SANITIZE_SMALL_KERNELS_ONLY"):
    print("SANITIZE_SMALL DONE (small kernels only)")
    sys.exit(0)
# blocked QB on a small matrix: sketch, TMA GEMMs with the -C preload epilogue, orthonormalisations
m, n, kstep = 1024, 768, 64
A = (rng.standard_normal((m, 100)) * np.logspace(0, -6, 100)) @ rng.standard_normal((100, n))
Ad = D.from_numpy_cm(A)
Q = torch.zeros((256, m), dtype=torch.float64, device="cuda"); B = torch.zeros((n, 256), dtype=torch.float64, device="cuda")
fr = C.c_longlong(0)
native.check(lib.rsvd_b200_randqb_dev(Ad.data_ptr(), m, n, m, kstep, 3, 0.0, 2, 1, 777, Q.data_ptr(), m, B.data_ptr(), 256, C.byref(fr)))
lib.rsvd_b200_sync()
# ill-conditioned panel -> shifted CholeskyQR3
Q0, _ = np.linalg.qr(rng.standard_normal((2000, 80)))
Yd = D.from_numpy_cm(Q0 * np.logspace(0, -11, 80))
native.check(lib.rsvd_b200_orthonormalize(Yd.data_ptr(), 2000, 2000, 80, None, 0))
lib.rsvd_b200_sync()
print("qr path", lib.rsvd_b200_get_option(b"last_qr_path"), "SANITIZE_SMALL DONE")
