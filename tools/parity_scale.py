"""tools/parity_scale.py — parity of the CUDA path against the compiled reference (oracle/_ref) at BENCHMARK scale, through
the reference's C API with host buffers, identical Omega (shared Philox generator, seed 777):

    python tools/parity_scale.py c2      BASELINE configs[1] in full: 50000 x 20000, k=500 p=20 q=2 (RRA:73-234)
    python tools/parity_scale.py c4s     configs[3] scaled in rows: 8000 x 50000, k=1000 p=20 q=2, two-sided ID + CUR
                                         (RRA:1863-1965, 2060-2082, 2191-2258): index sets bit-exact
    python tools/parity_scale.py geqp3   the two pivoted-QR shapes of configs[3] at FULL size, 1020 x 50000 and
                                         1000 x 400000, against LAPACK dgeqp3 (scipy): pivot vectors bit-exact
    python tools/parity_scale.py c3s     configs[2] scaled: blockrand QB in tolerance mode, TOL = 1e-6 REACHED
                                         (exact rank 600 + 1e-12 noise), 20000 x 5000, kstep = 200 (RRA:1576-1801)
    python tools/parity_scale.py all

Tolerances (BASELINE.json north_star, Omega imported): singular values 1e-10 relative, subspaces by principal angle,
reconstruction error within 1 %, ID/CUR index sets bit-exact.  Prints one line per check and PARITY_SCALE PASS|FAIL."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
os.environ["OPENBLAS_NUM_THREADS"] = os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
import torch  # noqa: E402
import lowrankmatrixdecompositioncodes_b200 as pkg  # noqa: E402
from lowrankmatrixdecompositioncodes_b200 import native  # noqa: E402
from oracle import ref_lib  # noqa: E402

lib = native.dev()
assert lib.rsvd_b200_init(0) == 0, lib.rsvd_b200_last_error().decode()
OK = True


def report(name, cond, detail):
    global OK
    OK = OK and bool(cond)
    print("%-58s %s  %s" % (name, "OK  " if cond else "FAIL", detail), flush=True)


def sin_theta(X, Y):
    """sin of the largest principal angle between span(X) and span(Y) (orthonormal columns) as ||(I - X X^T) Y||_2 —
    accurate for small angles, unlike sqrt(1 - cos^2) which bottoms out at sqrt(eps) ~ 1e-8."""
    R = Y - X @ (X.T @ Y)
    return float(np.linalg.norm(R, 2))


def device_matrix(m, n, r, lo, noise, seed):
    """column-major m x n in HBM: X diag(logspace(1, lo, r)) W^T + noise (torch only generates data)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn((m, r), dtype=torch.float64, device="cuda", generator=g) / m ** 0.5
    W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    sig = torch.logspace(1, lo, r, dtype=torch.float64, device="cuda")
    A = torch.empty((n, m), dtype=torch.float64, device="cuda")
    for j0 in range(0, n, 4096):
        j1 = min(n, j0 + 4096)
        torch.matmul(W[j0:j1] * sig, X.t(), out=A[j0:j1])
        if noise:
            A[j0:j1] += noise * torch.randn((j1 - j0, m), dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    return A


def host_mat_from_device(api, A_cm, m, n):
    M = api.lib.matrix_new(m, n)
    native.check(lib.rsvd_b200_d2h(C.cast(M.contents.d, C.c_void_p), A_cm.data_ptr(), m * n))
    return M


def quiet(fn):
    """run fn with fd 1 on /dev/null (the reference printf()s progress)"""
    sys.stdout.flush()
    dn, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(dn, 1)
    try:
        return fn()
    finally:
        os.dup2(saved, 1); os.close(dn); os.close(saved)


def c2():
    m, n, k, p, q, s = 50000, 20000, 500, 20, 2, 1
    api, L = pkg.Api(32), ref_lib.RefLib(32)
    A = device_matrix(m, n, 640, -3.0, 1e-6, 1234)
    M = host_mat_from_device(api, A, m, n)
    del A
    torch.cuda.empty_cache()
    api.set_seed(777)
    Um, Sm, Vm, fr = api.PM(), api.PM(), api.PM(), api.I(0)
    t0 = time.time()
    api.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, q, s, C.byref(fr), C.byref(Um), C.byref(Sm), C.byref(Vm))
    t_gpu = time.time() - t0
    api.check()
    api.lib.use_low_rank_svd_for_approximation(M, Um, Sm, Vm)
    pe = api.lib.rsvd_b200_api_last_percent_error()
    U, S, V = api.from_mat(Um), api.from_mat(Sm), api.from_mat(Vm)
    Mr = L.Mat(nrows=m, ncols=n, d=M.contents.d)
    PM = C.POINTER(L.Mat)
    Ur, Sr, Vr, frr = PM(), PM(), PM(), L.I(0)
    L.set_seed(777)
    t0 = time.time()
    quiet(lambda: L.lib.low_rank_svd_rand_decomp_fixed_rank(C.pointer(Mr), k, p, 1, q, s, C.byref(frr), C.byref(Ur), C.byref(Sr), C.byref(Vr)))
    t_cpu = time.time() - t0
    api.lib.use_low_rank_svd_for_approximation(M, C.cast(Ur, api.PM), C.cast(Sr, api.PM), C.cast(Vr, api.PM))
    per = api.lib.rsvd_b200_api_last_percent_error()
    Ur, Sr, Vr = L.from_mat(Ur), L.from_mat(Sr), L.from_mat(Vr)
    rel = float(np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr)))
    su, sv = sin_theta(U, Ur), sin_theta(V, Vr)
    align = float(np.max(1.0 - np.abs(np.sum(U * Ur, axis=0))))
    report("C2 svd_rand 50000x20000 k=500 p=20 q=2 (same Omega)", rel <= 1e-10 and su <= 1e-8 and sv <= 1e-8 and abs(pe - per) <= 0.01 * per,
           "max rel sigma err %.2e  sin(theta) U %.2e V %.2e  max(1-|u_i.u_i_ref|) %.2e  percent error %.6f (ref %.6f)  API call %.3f s (ref on %d host cores %.1f s)"
           % (rel, su, sv, align, pe, per, t_gpu, os.cpu_count(), t_cpu))
    api.lib.matrix_delete(M)


def c4s():
    m, n, k, p, q, s = 8000, 50000, 1000, 20, 2, 1
    api, L = pkg.Api(32), ref_lib.RefLib(32)
    A = device_matrix(m, n, 1536, -2.0, 1e-8, 99)
    An = A.t().cpu().numpy()
    del A
    torch.cuda.empty_cache()
    t0 = time.time()
    Ic, Ir, T, Sm = api.id_two_sided_rand(An, k, p, q, s, seed=777)
    t_gpu = time.time() - t0
    t0 = time.time()
    Icr, Irr, Tr, Smr = quiet(lambda: L.id_two_sided_rand(An, k, p, q, s, seed=777))
    t_cpu = time.time() - t0
    first = lambda a, b: int(np.argmax(a != b)) if not np.array_equal(a, b) else -1   # noqa: E731
    report("C4-scaled id_two_sided 8000x50000 k=1000 p=20 q=2", np.array_equal(Ic, Icr) and np.array_equal(Ir, Irr)
           and np.abs(T - Tr).max() <= 1e-9 * max(1.0, np.abs(Tr).max()) and np.abs(Sm - Smr).max() <= 1e-9 * max(1.0, np.abs(Smr).max()),
           "Icol bit-exact %s (first diff %d)  Irow bit-exact %s (first diff %d)  max|T-Tref| %.2e  max|S-Sref| %.2e  API %.2f s (ref %.1f s)"
           % (np.array_equal(Ic, Icr), first(Ic, Icr), np.array_equal(Ir, Irr), first(Ir, Irr), np.abs(T - Tr).max(), np.abs(Sm - Smr).max(), t_gpu, t_cpu))
    t0 = time.time()
    Cm, Um, Rm = api.cur_rand(An, k, p, q, s, seed=777)
    t_gpu = time.time() - t0
    t0 = time.time()
    Cr, Uc, Rr = quiet(lambda: L.cur_rand(An, k, p, q, s, seed=777))
    t_cpu = time.time() - t0
    # U solves an (R R^T) system (cond^2-sensitive): compare the APPROXIMATIONS, and U itself loosely
    nA = np.linalg.norm(An)
    e, er = np.linalg.norm(An - Cm @ (Um @ Rm)) / nA, np.linalg.norm(An - Cr @ (Uc @ Rr)) / nA
    report("C4-scaled cur_rand 8000x50000 k=1000", np.array_equal(Cm, Cr) and np.array_equal(Rm, Rr) and abs(e - er) <= 0.01 * er,
           "C bit-exact %s  R bit-exact %s  ||A-CUR||/||A|| %.6e (ref %.6e)  max|U-Uref|/max|Uref| %.2e  API %.2f s (ref %.1f s)"
           % (np.array_equal(Cm, Cr), np.array_equal(Rm, Rr), e, er, np.abs(Um - Uc).max() / np.abs(Uc).max(), t_gpu, t_cpu))


def geqp3():
    from scipy.linalg import lapack
    for (l, n, lo) in [(1020, 50000, -2.0), (1000, 400000, -2.0)]:
        g = torch.Generator(device="cuda").manual_seed(7 + n)
        R = torch.randn((l, l), dtype=torch.float64, device="cuda", generator=g)
        sig = torch.logspace(1, lo, l, dtype=torch.float64, device="cuda")
        Yt = torch.empty((n, l), dtype=torch.float64, device="cuda")           # column-major l x n
        for j0 in range(0, n, 50000):
            j1 = min(n, j0 + 50000)
            W = torch.randn((j1 - j0, l), dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
            torch.matmul(W * sig, R.t(), out=Yt[j0:j1])
        Yh = Yt.t().cpu().numpy()                                              # (l, n) Fortran-ordered view
        jp = torch.empty(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        lib.rsvd_b200_sync()
        t0 = time.time()
        native.check(lib.rsvd_b200_geqp3(Yt.data_ptr(), l, l, n, jp.data_ptr()))
        lib.rsvd_b200_sync()
        t_gpu = time.time() - t0
        ours = jp.cpu().numpy().astype(np.int64)
        Rd = np.abs(np.diagonal(Yt.t().cpu().numpy()[:l, :l]))
        del Yt
        t0 = time.time()
        qr, jpvt, tau, work, info = lapack.dgeqp3(np.asfortranarray(Yh))
        t_cpu = time.time() - t0
        ref = jpvt.astype(np.int64) - 1
        eq_lead = np.array_equal(ours[:l], ref[:l])
        nd = int(np.argmax(ours != ref)) if not np.array_equal(ours, ref) else -1
        rdiff = float(np.max(np.abs(Rd - np.abs(np.diagonal(qr[:l, :l]))) / np.abs(np.diagonal(qr[:l, :l]))))
        report("geqp3 %d x %d vs LAPACK dgeqp3" % (l, n), np.array_equal(ours, ref),
               "pivot vector bit-exact %s (leading %d: %s, first diff %d)  max rel |diag R| diff %.2e  device %.1f ms (LAPACK on host %.1f s)"
               % (np.array_equal(ours, ref), l, eq_lead, nd, rdiff, t_gpu * 1e3, t_cpu))
        del Yh, qr
        torch.cuda.empty_cache()


def c3s():
    m, n, kstep, q, s, r, tol = 20000, 5000, 200, 2, 1, 600, 1e-6
    api, L = pkg.Api(32), ref_lib.RefLib(32)
    g = np.random.default_rng(5)
    X, _ = np.linalg.qr(g.standard_normal((m, r)))
    W, _ = np.linalg.qr(g.standard_normal((n, r)))
    A = (X * np.logspace(1, -2, r)) @ W.T + 1e-12 * g.standard_normal((m, n))
    t0 = time.time()
    f, Qm, Bm = api.randQB_pb_new(A, kstep, 0, tol, q, s, seed=777)
    t_gpu = time.time() - t0
    t0 = time.time()
    fr, Qr, Br = quiet(lambda: L.randQB_pb_new(A, kstep, 0, tol, q, s, seed=777))
    t_cpu = time.time() - t0
    nA = np.linalg.norm(A)
    res, resr = np.linalg.norm(A - Qm @ Bm), np.linalg.norm(A - Qr @ Br)
    d = np.linalg.norm(Qm @ Bm - Qr @ Br) / nA
    orth = np.abs(Qm.T @ Qm - np.eye(Qm.shape[1])).max()
    report("C3-scaled randQB_pb_new tol mode 20000x5000 TOL=1e-6", f == fr and res < tol and d <= 1e-10 and sin_theta(Qm[:, :r], Qr[:, :r]) <= 1e-6,
           "frank %d (ref %d)  ||A-QB||_F %.3e (ref %.3e) < TOL %.0e: %s  ||QB-QrBr||/||A|| %.2e  sin(theta) first %d cols %.2e  ||QtQ-I||max %.1e  API %.2f s (ref %.1f s)"
           % (f, fr, res, resr, tol, res < tol, d, r, sin_theta(Qm[:, :r], Qr[:, :r]), orth, t_gpu, t_cpu))
    # the blockrand SVD entry point on the same input, rank mode with l = k + p = the exact rank (a fourth block would be
    # drawn from the 1e-12 noise floor, whose rounding-level residual makes sigma agree only to ~1e-9 between ANY two implementations)
    fo, U, S, V = api.svd_blockrand(A, 400, 200, tol, 1, kstep, q, s, seed=777)
    frr, Ur, Sr, Vr = quiet(lambda: L.svd_blockrand(A, 400, 200, tol, 1, kstep, q, s, seed=777))
    rel = float(np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr)))
    report("C3-scaled low_rank_svd_blockrand k=400 p=200 kstep=200", fo == frr and rel <= 1e-10 and sin_theta(U, Ur) <= 1e-6,
           "frank %d (ref %d)  max rel sigma err %.2e  sin(theta) U %.2e V %.2e" % (fo, frr, rel, sin_theta(U, Ur), sin_theta(V, Vr)))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if not ref_lib.available(32):
        print("oracle/_ref is not built"); sys.exit(2)
    for name, fn in (("c2", c2), ("c4s", c4s), ("geqp3", geqp3), ("c3s", c3s)):
        if what in (name, "all"):
            fn()
    print("status:", lib.rsvd_b200_status(), lib.rsvd_b200_last_error().decode())
    print("PARITY_SCALE", "PASS" if OK and lib.rsvd_b200_status() == 0 else "FAIL", flush=True)
    sys.exit(0 if OK else 1)
