"""tools/jacobi_live_check.py — one-sided Jacobi SVD with V rebuilt next to the Jacobi kernel (live replay on a side stream)
against the after-the-fact replay: identical factors, time of the whole rsvd_b200_svd_small call."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowrankmatrixdecompositioncodes_b200 import native, device as D
lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
ok_all = True
for n in ((130, 520) if os.environ.get("RSVD_B200_JACOBI_GRAM") else (2, 7, 64, 130, 520, 592, 593, 777, 1050)):
    rng = np.random.default_rng(n)
    U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    R = np.triu(np.linalg.qr((U0 * np.logspace(1, -2.5, n)) @ V0.T)[1])
    out = {}
    for live in (1, 0):
        lib.rsvd_b200_set_option(b"no_live_replay", 1 - live)
        ts = []
        for rep in range(3):
            Ad = D.from_numpy_cm(R)
            U = torch.empty((n, n), dtype=torch.float64, device="cuda"); Vt = torch.empty_like(U); s = torch.empty(n, dtype=torch.float64, device="cuda")
            Vt.fill_(float("nan"))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(D.stream()):
                e0.record()
                native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
                e1.record()
            native.check(lib.rsvd_b200_sync()); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[live] = (U.t().cpu().numpy(), s.cpu().numpy(), Vt.t().cpu().numpy(), min(ts))
    (U1, s1, V1, t1), (U0_, s0, V0_, t0) = out[1], out[0]
    recon = np.linalg.norm((U1 * s1) @ V1 - R) / np.linalg.norm(R)
    orth = np.abs(V1 @ V1.T - np.eye(n)).max()
    same = bool(np.array_equal(V1, V0_) and np.array_equal(s1, s0) and np.array_equal(U1, U0_))
    sv = np.linalg.svd(R, compute_uv=False)
    ok = recon < 1e-13 and orth < 1e-13 and same and np.max(np.abs(s1 - sv) / sv[0]) < 1e-13
    ok_all &= ok
    print("n=%5d  %s  live %.3f ms   after-the-fact %.3f ms   recon %.2e  ||VVt-I|| %.2e  bitwise equal to the after-the-fact replay: %s"
          % (n, "OK  " if ok else "FAIL", t1, t0, recon, orth, same), flush=True)
print("JACOBI_LIVE", "PASS" if ok_all else "FAIL")
