"""tools/profile_gemm.py — runs the three streaming-pass kernels once each at the BASELINE configs[1] shape
(for `ncu --set full -k regex:gemm_tma_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
m, n, l = 50000, 20000, 520
A = torch.randn((n, m), dtype=torch.float64, device="cuda")
B = torch.randn((l, n), dtype=torch.float64, device="cuda")
Y = torch.empty((l, m), dtype=torch.float64, device="cuda")
Z = torch.empty((l, n), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    D.gemm("N", "N", m, l, n, A, m, B, n, Y, m)
    D.gemm("T", "N", n, l, m, A, m, Y, m, Z, n)
    native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y.data_ptr(), m))
lib.rsvd_b200_sync()
print("done")
