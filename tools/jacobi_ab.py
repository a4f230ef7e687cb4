"""Timing and accuracy of the one-sided Jacobi kernel on a B200 (run with RSVD_B200_JACOBI_BW=1|2|4 to compare the number of
columns a CTA rotates between two device-wide barriers)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowrankmatrixdecompositioncodes_b200 import native, device as D
lib = native.dev()
lib.rsvd_b200_init(0)
lib.rsvd_b200_set_option(b"verbose", 1)
def sync(): native.check(lib.rsvd_b200_sync())
n = 520
rng = np.random.default_rng(n)
U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
s0 = np.logspace(0, -5, n)
S = (V0 * s0 ** 2) @ V0.T
S = (S + S.T) / 2
R = np.triu(np.linalg.qr((U0 * np.logspace(1, -2.5, n)) @ V0.T)[1])
for mode in ("bw=" + os.environ.get("RSVD_B200_JACOBI_BW", "default"),):
    Sd = D.from_numpy_cm(S)
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_eig_small(Sd.data_ptr(), n, n, w.data_ptr())); sync()
    wn = w.cpu().numpy()
    err = np.abs(wn[::-1] - s0 ** 2)
    print(mode, "eig: max abs err %.3e at idx %d (lambda %.3e)" % (err.max(), err.argmax(), s0[err.argmax()] ** 2), flush=True)
    for rep in range(2):
        Ad = D.from_numpy_cm(R)
        U = torch.empty((n, n), dtype=torch.float64, device="cuda"); Vt = torch.empty_like(U); s = torch.empty(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize(); t0 = time.time()
        native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n)); sync()
        dt = time.time() - t0
    Un, Vtn, sn = U.t().cpu().numpy(), Vt.t().cpu().numpy(), s.cpu().numpy()
    print(mode, "svd(R): %.2f ms  recon %.2e" % (dt * 1e3, np.linalg.norm((Un * sn) @ Vtn - R) / np.linalg.norm(R)), flush=True)
