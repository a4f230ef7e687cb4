"""tools/chol_check.py — the fused Cholesky + inverse dataflow kernel (cholinv.cu) against numpy, and its time against the
per-block launch sequence it replaces (option no_chol_dataflow).  `gpurun -- python tools/chol_check.py`."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402

lib = native.dev()
assert lib.rsvd_b200_init(0) == 0
ok_all = True


def run(n, cond=1e3, seed=0, reps=5):
    global ok_all
    rng = np.random.default_rng(seed)
    m = max(2 * n, 64)
    Y = rng.standard_normal((m, n)) * np.logspace(0, -np.log10(cond), n)
    G = Y.T @ Y
    Rref = np.linalg.cholesky(G).T
    res = {}
    for mode in (0, 1):
        lib.rsvd_b200_set_option(b"no_chol_dataflow", mode)
        ts = []
        for it in range(reps):
            Gd = D.from_numpy_cm(np.triu(G) + np.tril(np.full((n, n), np.nan), -1) if mode == 0 else G)   # the lower triangle must never be read
            Xd = D.new_cm(n, n)
            Xd.fill_(float("nan") if mode == 0 else 0.0)
            mm = (C.c_double * 2)()
            torch.cuda.synchronize()
            st = D.stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record()
                info = lib.rsvd_b200_chol_inv(D.ptr(Gd), n, n, D.ptr(Xd), n, mm)
                e1.record()
            lib.rsvd_b200_sync()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        R, X = D.to_numpy(Gd), D.to_numpy(Xd)
        res[mode] = (info, R, X, min(ts), (mm[0], mm[1]))
    info, R, X, t_new, mm = res[0]
    _, R_old, X_old, t_old, _ = res[1]
    eR = np.abs(R - Rref).max() / np.abs(Rref).max()
    eI = np.abs(R @ X - np.eye(n)).max()
    lower_zero = bool(np.all(np.tril(R, -1) == 0) and np.all(np.tril(X, -1) == 0))
    d = np.abs(np.diag(Rref))
    mm_ok = abs(mm[0] - d.min()) <= 1e-10 * d.max() and abs(mm[1] - d.max()) <= 1e-10 * d.max()
    eI_old = np.abs(np.triu(R_old) @ np.triu(X_old) - np.eye(n)).max()
    ok = info == 0 and eR < 1e-9 and eI < 1e-7 * max(1.0, cond / 1e3) and lower_zero and mm_ok
    ok_all &= ok
    print("n=%5d cond=%.0e  %s  |R-Rref|/|R| %.2e  |R X - I| %.2e (old path %.2e)  zeros below %s  minmax ok %s   dataflow %.3f ms   per-block sequence %.3f ms"
          % (n, cond, "OK  " if ok else "FAIL", eR, eI, eI_old, lower_zero, mm_ok, t_new, t_old), flush=True)


for n in ((520,) if os.environ.get("CHOL_CHECK_QUICK") else (1, 5, 32, 33, 64, 100, 257, 520, 544, 545, 1020, 1050, 2048)):
    run(n)
run(520, cond=1e6)
# not positive definite: the failing column must be reported
n = 200
A = np.eye(n); A[150, 150] = -1.0
Gd = D.from_numpy_cm(A); Xd = D.new_cm(n, n)
lib.rsvd_b200_set_option(b"no_chol_dataflow", 0)
info = lib.rsvd_b200_chol_inv(D.ptr(Gd), n, n, D.ptr(Xd), n, None)
print("indefinite matrix: info =", info, "(expected 151)")
ok_all &= (info == 151)

# orthonormalize timing at the bench shapes
for (m, l) in ((50000, 520), (20000, 520), (125000, 1050)):
    for mode in (1, 0):
        lib.rsvd_b200_set_option(b"no_chol_dataflow", mode)
        gen = torch.Generator(device="cuda").manual_seed(1)
        Y0 = torch.randn((l, m), dtype=torch.float64, device="cuda", generator=gen)
        ts = []
        for it in range(4):
            Y = Y0.clone()
            torch.cuda.synchronize()
            st = D.stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record()
                rc = lib.rsvd_b200_orthonormalize(D.ptr(Y), m, m, l, None, 0)
                e1.record()
            lib.rsvd_b200_sync(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        Q = Y.t()
        orth = (Q.t() @ Q - torch.eye(l, dtype=torch.float64, device="cuda")).abs().max().item()
        print("orthonormalize %d x %d  %s: %.3f ms   ||QtQ - I||max %.2e  path %d" % (
            m, l, "per-block sequence" if mode else "dataflow kernel   ", min(ts), orth, lib.rsvd_b200_get_option(b"last_qr_path")), flush=True)
        ok_all &= orth < 1e-13
        del Y, Y0, Q
print("CHOL_CHECK", "PASS" if ok_all else "FAIL")
