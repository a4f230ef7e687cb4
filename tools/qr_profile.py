"""tools/qr_profile.py l n [reps] — one pivoted QR (rsvd_b200_geqp3) of a sketch-like l x n matrix, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qr.csv python tools/qr_profile.py 1020 50000"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lowrankmatrixdecompositioncodes_b200 import native  # noqa: E402

l, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
g = torch.Generator(device="cuda").manual_seed(7 + n)
R = torch.randn((l, l), dtype=torch.float64, device="cuda", generator=g)
sig = torch.logspace(1, -2, l, dtype=torch.float64, device="cuda")
Y0 = torch.empty((n, l), dtype=torch.float64, device="cuda")
for j0 in range(0, n, 50000):
    j1 = min(n, j0 + 50000)
    W = torch.randn((j1 - j0, l), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    torch.matmul(W * sig, R.t(), out=Y0[j0:j1])
jp = torch.empty(n, dtype=torch.float64, device="cuda")
for it in range(reps):
    Y = Y0.clone()
    torch.cuda.synchronize()
    t0 = time.time()
    native.check(lib.rsvd_b200_geqp3(Y.data_ptr(), l, l, n, jp.data_ptr()))
    lib.rsvd_b200_sync()
    print("geqp3 %d x %d: %.2f ms" % (l, n, (time.time() - t0) * 1e3), flush=True)
