"""tools/mg_check.py — single-process multi-GPU behind the UNCHANGED C API (csrc/device/multi.cu, hostapi.cu).

    RSVD_B200_DEVICES=0-7 python tools/mg_check.py [--big]

One Python process (standing in for the reference's single-threaded C driver, multi_core_mkl_code_64bit/driver1.c:40-50)
calls low_rank_svd_rand_decomp_fixed_rank / randQB_pb_new / low_rank_svd_blockrand_... / id_two_sided_rand_... /
cur_rand_... on host `mat` structs; the library row-partitions every call over the listed devices.  Checked against the
compiled reference (same Omega) and against the same call on ONE device; also the residency cache (rsvd_b200_pin_matrix).
Prints one line per check and MG_CHECK PASS|FAIL."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import lowrankmatrixdecompositioncodes_b200 as pkg  # noqa: E402
from lowrankmatrixdecompositioncodes_b200 import native  # noqa: E402
from oracle import ref_lib, rsvd_numpy as O  # noqa: E402

lib = native.dev()
OK = True


def report(name, cond, detail):
    global OK
    OK = OK and bool(cond)
    print("%-60s %s  %s" % (name, "OK  " if cond else "FAIL", detail), flush=True)


def sin_theta(X, Y):
    return float(np.linalg.norm(Y - X @ (X.T @ Y), 2))


def devices(n):
    ids = (C.c_int * max(n, 1))(*range(max(n, 1)))
    native.check(lib.rsvd_b200_set_devices(n, ids))
    return lib.rsvd_b200_active_devices()


def quiet(fn):
    sys.stdout.flush()
    dn, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(dn, 1)
    try:
        return fn()
    finally:
        os.dup2(saved, 1); os.close(dn); os.close(saved)


def main():
    big = "--big" in sys.argv
    ndev = lib.rsvd_b200_device_count()
    want = lib.rsvd_b200_active_devices()        # reads RSVD_B200_DEVICES
    if want < 2:
        want = devices(min(ndev, 8))
    print("devices on the box: %d, active workers: %d" % (ndev, want), flush=True)
    if want < 2:
        print("MG_CHECK SKIP (needs >= 2 GPUs)")
        return 0
    api = pkg.Api(32)
    L = ref_lib.RefLib(32) if ref_lib.available(32) else None

    if "--only-scale" not in sys.argv:
        # ---- against the compiled reference, same Omega -------------------------------------------------------------------
        for (m, n, k, p, q, s, spec) in [(6000, 1500, 100, 20, 2, 1, "gap"), (9001, 900, 40, 10, 2, 1, "logspace")]:
            A, _ = O.make_matrix(m, n, spec, seed=3, k=k, tail=1e-7)
            U, S, V = api.svd_rand(A, k, p, 1, q, s, seed=777)
            Ur, Sr, Vr = quiet(lambda: L.svd_rand(A, k, p, 1, q, s, seed=777)) if L else O.low_rank_svd_rand_decomp_fixed_rank(A, k, p, 1, q, s, 777)
            rel = float(np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr)))
            e, er = np.linalg.norm(A - U @ S @ V.T) / np.linalg.norm(A), np.linalg.norm(A - Ur @ Sr @ Vr.T) / np.linalg.norm(A)
            report("svd_rand %dx%d on %d GPUs (%s)" % (m, n, want, spec), rel < 1e-10 and abs(e - er) <= 0.01 * er and sin_theta(U, Ur) < 1e-6,
                   "max rel sigma err %.2e  sin(theta) U %.2e  recon %.6e (ref %.6e)  ||UtU-I|| %.1e" % (rel, sin_theta(U, Ur), e, er, np.abs(U.T @ U - np.eye(k)).max()))
            Ic, Ir, T, Sm = api.id_two_sided_rand(A, k, p, q, s, seed=777)
            Icr, Irr, Tr, Smr = quiet(lambda: L.id_two_sided_rand(A, k, p, q, s, seed=777)) if L else O.id_two_sided_rand_decomp_fixed_rank(A, k, p, q, s, 777)
            report("id_two_sided %dx%d on %d GPUs" % (m, n, want), np.array_equal(Ic, Icr) and np.array_equal(Ir, Irr) and np.abs(T - Tr).max() < 1e-9 and np.abs(Sm - Smr).max() < 1e-9,
                   "Icol bit-exact %s  Irow bit-exact %s  max|T-Tref| %.2e  max|S-Sref| %.2e" % (np.array_equal(Ic, Icr), np.array_equal(Ir, Irr), np.abs(T - Tr).max(), np.abs(Sm - Smr).max()))
            Cm, Um, Rm = api.cur_rand(A, k, p, q, s, seed=777)
            Cr, Uc, Rr = quiet(lambda: L.cur_rand(A, k, p, q, s, seed=777)) if L else O.cur_rand_decomp_fixed_rank(A, k, p, q, s, 777)
            report("cur_rand %dx%d on %d GPUs" % (m, n, want), np.array_equal(Cm, Cr) and np.array_equal(Rm, Rr) and np.abs(Um - Uc).max() <= 1e-6 * np.abs(Uc).max(),
                   "C bit-exact %s  R bit-exact %s  max|U-Uref|/max|U| %.2e" % (np.array_equal(Cm, Cr), np.array_equal(Rm, Rr), np.abs(Um - Uc).max() / np.abs(Uc).max()))
            for (kstep, nstep, tol) in [(20, 4, 0.0), (20, 0, float(np.linalg.norm(A)) * 0.3)]:
                f, Qm, Bm = api.randQB_pb_new(A, kstep, nstep, tol, q, s, seed=777)
                fr, Qr, Br = quiet(lambda: L.randQB_pb_new(A, kstep, nstep, tol, q, s, seed=777)) if L else O.randQB_pb_new(A, kstep, nstep, tol, q, s, 777)
                d = np.linalg.norm(Qm @ Bm - Qr @ Br) / np.linalg.norm(A)
                report("randQB_pb_new kstep=%d nstep=%d tol=%.3g on %d GPUs" % (kstep, nstep, tol, want), f == fr and d < 1e-11, "frank %d (ref %d)  ||QB-QrBr||/||A|| %.2e" % (f, fr, d))
            fo, U, S, V = api.svd_blockrand(A, 60, 20, 0.0, 1, 20, q, s, seed=777)
            frr, Ur, Sr, Vr = quiet(lambda: L.svd_blockrand(A, 60, 20, 0.0, 1, 20, q, s, seed=777)) if L else O.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(A, 60, 20, 0.0, 1, 20, q, s, 777)
            rel = float(np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr)))
            report("low_rank_svd_blockrand k=60 p=20 kstep=20 on %d GPUs" % want, fo == frr and rel < 1e-10 and sin_theta(U, Ur) < 1e-6,
                   "frank %d (ref %d)  max rel sigma err %.2e  sin(theta) U %.2e" % (fo, frr, rel, sin_theta(U, Ur)))

    # ---- N devices against ONE device at a size where the partition matters, and the residency cache --------------------
    m, n, k, p, q, s = (50000, 100000, 1000, 20, 3, 1) if big else (40000, 12000, 300, 20, 2, 1)    # --big: the shape of driver1.c (64-bit ABI)
    if "--c2" in sys.argv:      # BASELINE configs[1] from ONE host matrix: end-to-end strong scaling through the unchanged C API
        m, n, k, p, q, s = 50000, 20000, 500, 20, 2, 1
    apiL = pkg.Api(64) if big else api
    M = apiL.lib.matrix_new(m, n)
    t0 = time.time()
    apiL.set_seed(5)
    apiL.lib.initialize_random_matrix(M)          # Gaussian, as driver1.c:33 — generated on the device, no structure: a hard case for sigma agreement
    # give it a decaying column scaling so that the leading singular values are separated
    buf = np.ctypeslib.as_array(M.contents.d, shape=(n, m))
    buf *= np.logspace(0, -3, n)[:, None]
    print("   %d x %d host matrix ready in %.1f s" % (m, n, time.time() - t0), flush=True)
    apiL.set_seed(777)
    res = {}
    for nd in (want, 1):
        devices(nd)
        Um, Sm, Vm, fr = apiL.PM(), apiL.PM(), apiL.PM(), apiL.I(0)
        times = []
        for it in range(2):
            t0 = time.time()
            apiL.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, q, s, C.byref(fr), C.byref(Um), C.byref(Sm), C.byref(Vm))
            times.append(time.time() - t0)
            apiL.check()
            if it == 0:
                for x in (Um, Sm, Vm):
                    apiL.lib.matrix_delete(x)
        res[nd] = (apiL.from_mat(Um), np.diag(apiL.from_mat(Sm)).copy(), apiL.from_mat(Vm), min(times))
    devices(want)
    (Un, Sn, Vn, tn), (U1, S1, V1, t1) = res[want], res[1]
    rel = float(np.max(np.abs(Sn - S1) / S1))
    report("svd_rand %dx%d k=%d q=%d: %d GPUs vs 1 GPU" % (m, n, k, q, want), rel < 1e-10 and sin_theta(U1, Un) < 1e-6 and sin_theta(V1, Vn) < 1e-6,
           "max rel sigma diff %.2e  sin(theta) U %.2e V %.2e  API call %.3f s on %d GPUs, %.3f s on 1" % (rel, sin_theta(U1, Un), sin_theta(V1, Vn), tn, want, t1))
    del res, Un, Vn, U1, V1
    # residency: pin, first call uploads, second call must not
    native.check(lib.rsvd_b200_pin_matrix(C.cast(M.contents.d, C.c_void_p), m, n))
    tt = []
    for it in range(2):
        Um, Sm, Vm, fr = apiL.PM(), apiL.PM(), apiL.PM(), apiL.I(0)
        t0 = time.time()
        apiL.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, q, s, C.byref(fr), C.byref(Um), C.byref(Sm), C.byref(Vm))
        tt.append(time.time() - t0)
        apiL.check()
        S2 = np.diag(apiL.from_mat(Sm)).copy()
        for x in (Um, Vm):
            apiL.lib.matrix_delete(x)
    resident = lib.rsvd_b200_is_resident(C.cast(M.contents.d, C.c_void_p))
    # (the resident call sketches A in one launch, the uploading call chunk by chunk: same sums in a different order)
    drel = float(np.max(np.abs(S2 - Sn) / Sn))
    report("residency: pinned matrix, 2nd call skips the upload", resident == 1 and drel < 1e-12 and tt[1] < tt[0],
           "resident %d  max rel sigma diff to the unpinned call %.2e  1st call %.3f s, 2nd call %.3f s" % (resident, drel, tt[0], tt[1]))
    lib.rsvd_b200_unpin_matrix(C.cast(M.contents.d, C.c_void_p))
    apiL.lib.matrix_delete(M)
    print("status:", lib.rsvd_b200_status(), lib.rsvd_b200_last_error().decode())
    print("MG_CHECK", "PASS" if OK and lib.rsvd_b200_status() == 0 else "FAIL", flush=True)
    return 0 if OK else 1


if __name__ == "__main__":
    sys.exit(main())
