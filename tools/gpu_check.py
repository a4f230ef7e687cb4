"""tools/gpu_check.py — developer diagnostics: exercises every kernel family once on a B200 and prints
accuracy / timing figures.  Not a test (tests/ has the assertions); meant for `gpurun -- python tools/gpu_check.py`."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import lowrankmatrixdecompositioncodes_b200 as pkg  # noqa: E402
from lowrankmatrixdecompositioncodes_b200 import device as D, native  # noqa: E402
from oracle import ref_lib, rsvd_numpy as O  # noqa: E402

lib = native.dev()
SECTIONS = sys.argv[1:] or ["peak", "rng", "gemm", "tma", "perf", "qr", "jacobi", "geqp3", "svd", "id", "qb"]


def section(name):
    def deco(fn):
        if name in SECTIONS:
            print("\n===== %s =====" % name, flush=True)
            try:
                fn()
            except Exception:
                traceback.print_exc()
                lib.rsvd_b200_clear_error()
            sys.stdout.flush()
        return fn
    return deco


def sync():
    lib.rsvd_b200_sync()
    torch.cuda.synchronize()


def timeit(fn, reps=3, warm=1):
    st = D.stream()
    for _ in range(warm):
        fn()
    sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            fn()
            e1.record()
        sync()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


@section("peak")
def _peak():
    assert lib.rsvd_b200_init(0) == 0, lib.rsvd_b200_last_error()
    print("SMs:", lib.rsvd_b200_get_option(b"sms"))
    for it in (2000, 20000):
        print("DMMA peak TFLOP/s (iters=%d): %.2f" % (it, lib.rsvd_b200_fp64_peak_tflops(it, 0)))
        print("DFMA peak TFLOP/s (iters=%d): %.2f" % (it, lib.rsvd_b200_fp64_peak_tflops(it, 1)))


@section("rng")
def _rng():
    n = 1_000_003
    t = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_fill_normal(t.data_ptr(), n, 12345, 77))
    sync()
    ref = ref_lib.normal_stream(12345, 77, n)
    got = t.cpu().numpy()
    print("device Philox+BoxMuller vs CPU: mismatches = %d of %d, max abs diff = %g" % ((got != ref).sum(), n, np.abs(got - ref).max()))


def _gemm_case(ta, tb, m, n, k, force_generic, alpha=1.0, beta=0.0, odd_ld=False):
    rng = np.random.default_rng(m * 31 + n * 7 + k)
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + (1 if odd_ld else 0), br + (1 if odd_ld else 0), m + (1 if odd_ld else 0)
    A = torch.from_numpy(rng.standard_normal((ac, lda))).cuda()
    B = torch.from_numpy(rng.standard_normal((bc, ldb))).cuda()
    Cm = torch.from_numpy(rng.standard_normal((n, ldc))).cuda()
    C0 = Cm.clone()
    lib.rsvd_b200_set_option(b"force_generic_gemm", 1 if force_generic else 0)
    D.gemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha, beta)
    sync()
    path = lib.rsvd_b200_get_option(b"last_gemm_path")
    lib.rsvd_b200_set_option(b"force_generic_gemm", 0)
    Am = A[:, :ar].t() if True else None
    Bm = B[:, :br].t()
    opA = Am if ta == "N" else Am.t()
    opB = Bm if tb == "N" else Bm.t()
    ref = alpha * (opA @ opB) + beta * C0[:, :m].t()
    got = Cm[:, :m].t()
    err = (got - ref).abs().max().item() / max(1e-300, ref.abs().max().item())
    pad_ok = True
    if odd_ld:
        pad_ok = torch.equal(Cm[:, m:], C0[:, m:])
    return err, path, pad_ok


@section("gemm")
def _gemm():
    for (ta, tb, m, n, k) in [("N", "N", 100, 70, 50), ("T", "N", 129, 65, 1000), ("N", "T", 64, 64, 16), ("T", "T", 33, 17, 9),
                             ("T", "N", 120, 120, 5000), ("N", "N", 1, 1, 1), ("N", "N", 300, 200, 3)]:
        err, path, ok = _gemm_case(ta, tb, m, n, k, True, 1.3, 0.7, odd_ld=True)
        print("generic %s%s m=%d n=%d k=%d: rel err %.2e path=%d pad_intact=%s" % (ta, tb, m, n, k, err, path, ok))


@section("tma")
def _tma():
    for (ta, tb, m, n, k, al, be) in [("N", "N", 1024, 256, 512, 1.0, 0.0), ("T", "N", 1024, 256, 512, 1.0, 0.0),
                                      ("N", "N", 1000, 130, 530, 1.5, 0.5), ("T", "N", 778, 120, 2050, -1.0, 1.0),
                                      ("T", "N", 520, 520, 50000, 1.0, 0.0), ("N", "N", 5000, 200, 200, -1.0, 1.0),
                                      ("N", "N", 4112, 520, 3000, 1.0, 0.0)]:
        err, path, ok = _gemm_case(ta, tb, m, n, k, False, al, be)
        print("tma %s%s m=%d n=%d k=%d alpha=%g beta=%g: rel err %.2e path=%d" % (ta, tb, m, n, k, al, be, err, path))
    # fused Philox sketch vs explicit Omega, both index maps
    for (ta, m, n, k, sk, sc) in [("N", 2000, 120, 1500, 1, 1500), ("T", 1500, 120, 2000, 120, 1), ("N", 4096, 520, 2051, 1, 2051)]:
        rng = np.random.default_rng(1)
        ar, ac = (m, k) if ta == "N" else (k, m)
        A = torch.from_numpy(rng.standard_normal((ac, ar))).cuda()
        Cm = torch.empty((n, m), dtype=torch.float64, device="cuda")
        for force in (0, 1):
            lib.rsvd_b200_set_option(b"force_generic_gemm", force)
            native.check(lib.rsvd_b200_sketch(ta.encode(), m, n, k, A.data_ptr(), ar, 99, sk, sc, 5, Cm.data_ptr(), m))
            sync()
            path = lib.rsvd_b200_get_option(b"last_gemm_path")
            idx = 5 + np.arange(k)[:, None] * sk + np.arange(n)[None, :] * sc
            flat = ref_lib.normal_stream(99, 0, int(idx.max()) + 1)
            Om = torch.from_numpy(flat[idx]).cuda()
            opA = A.t() if ta == "N" else A
            ref = opA @ Om
            err = (Cm.t() - ref).abs().max().item() / ref.abs().max().item()
            print("sketch %s m=%d n=%d k=%d sk=%d sc=%d force_generic=%d: rel err %.2e path=%d" % (ta, m, n, k, sk, sc, force, err, path))
        lib.rsvd_b200_set_option(b"force_generic_gemm", 0)


@section("perf")
def _perf():
    for (ta, m, n, k) in [("N", 32768, 512, 8192), ("T", 8192, 512, 32768), ("N", 50000, 520, 20000), ("T", 20000, 520, 50000)]:
        ar, ac = (m, k) if ta == "N" else (k, m)
        A = torch.randn((ac, ar), dtype=torch.float64, device="cuda")
        B = torch.randn((n, k), dtype=torch.float64, device="cuda")
        Cm = torch.empty((n, m), dtype=torch.float64, device="cuda")
        sync()
        t = timeit(lambda: D.gemm(ta, "N", m, n, k, A, ar, B, k, Cm, m))
        print("tma gemm %sN m=%d n=%d k=%d: %.3f ms  %.2f TFLOP/s (path=%d)" % (ta, m, n, k, t * 1e3, 2.0 * m * n * k / t / 1e12, lib.rsvd_b200_get_option(b"last_gemm_path")))
        if ta == "N":
            t = timeit(lambda: native.check(lib.rsvd_b200_sketch(b"N", m, n, k, A.data_ptr(), ar, 7, 1, k, 0, Cm.data_ptr(), m)))
            print("  fused Philox sketch: %.3f ms  %.2f TFLOP/s" % (t * 1e3, 2.0 * m * n * k / t / 1e12))
        t = timeit(lambda: torch.matmul(A if ta == "T" else A.t(), B.t()))
        print("  torch(cuBLAS) fp64 matmul for scale: %.3f ms  %.2f TFLOP/s" % (t * 1e3, 2.0 * m * n * k / t / 1e12))
        del A, B, Cm
        torch.cuda.empty_cache()


@section("qr")
def _qr():
    for (m, l, cond, force) in [(5000, 120, 1e3, 0), (5000, 120, 1e12, 0), (3000, 200, 1e2, 1), (50000, 520, 1e5, 0)]:
        rng = np.random.default_rng(3)
        Q0, _ = np.linalg.qr(rng.standard_normal((m, l)))
        W, _ = np.linalg.qr(rng.standard_normal((l, l)))
        Y = (Q0 * np.logspace(0, -np.log10(cond), l)) @ W.T
        Yd = D.from_numpy_cm(Y)
        R = torch.zeros((l, l), dtype=torch.float64, device="cuda")
        lib.rsvd_b200_set_option(b"force_qr_fallback", force)
        t0 = time.time()
        native.check(lib.rsvd_b200_orthonormalize(Yd.data_ptr(), m, m, l, R.data_ptr(), l))
        sync()
        dt = time.time() - t0
        lib.rsvd_b200_set_option(b"force_qr_fallback", 0)
        Q = D.to_numpy(Yd)
        Rn = R.t().cpu().numpy()
        print("orthonormalize m=%d l=%d cond=%.0e force_fallback=%d: path=%d  ||QtQ-I||=%.2e  ||QR-Y||/||Y||=%.2e  lower(R)=%.1e  %.1f ms"
              % (m, l, cond, force, lib.rsvd_b200_get_option(b"last_qr_path"), np.abs(Q.T @ Q - np.eye(l)).max(),
                 np.linalg.norm(Q @ Rn - Y) / np.linalg.norm(Y), np.abs(np.tril(Rn, -1)).max(), dt * 1e3))


@section("jacobi")
def _jacobi():
    for n in (5, 120, 520, 1050):
        rng = np.random.default_rng(n)
        U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
        V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = np.triu(np.linalg.qr((U0 * np.logspace(1, -2.5, n)) @ V0.T)[1])   # triangular, cond ~3e3 (like Rhat of the pipeline)
        Ad = D.from_numpy_cm(A)
        U = torch.empty((n, n), dtype=torch.float64, device="cuda")
        Vt = torch.empty((n, n), dtype=torch.float64, device="cuda")
        s = torch.empty(n, dtype=torch.float64, device="cuda")
        A0d = Ad.clone()
        native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
        sync()
        Ad.copy_(A0d)
        torch.cuda.synchronize()
        t0 = time.time()
        native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
        sync()
        dt = time.time() - t0
        sn = s.cpu().numpy()
        Un, Vtn = U.t().cpu().numpy(), Vt.t().cpu().numpy()
        sref = np.linalg.svd(A, compute_uv=False)
        print("jacobi_svd n=%d: max rel sigma err %.2e  ||USVt-A||/||A||=%.2e  ||UtU-I||=%.2e  %.1f ms"
              % (n, np.max(np.abs(sn - sref) / sref), np.linalg.norm((Un * sn) @ Vtn - A) / np.linalg.norm(A),
                 np.abs(Un.T @ Un - np.eye(n)).max(), dt * 1e3))


@section("geqp3")
def _geqp3():
    from scipy.linalg import lapack
    for (m, n) in [(12, 40), (120, 1500), (100, 2000), (300, 5000), (1020, 20000)]:
        rng = np.random.default_rng(m + n)
        Y = rng.standard_normal((m, 60)) @ (np.logspace(0, -5, 60)[:, None] * rng.standard_normal((60, n))) if m > 60 else rng.standard_normal((m, n))
        Y = Y + 1e-9 * rng.standard_normal((m, n))
        Yd = D.from_numpy_cm(Y)
        jp = torch.empty(n, dtype=torch.float64, device="cuda")
        t0 = time.time()
        native.check(lib.rsvd_b200_geqp3(Yd.data_ptr(), m, m, n, jp.data_ptr()))
        sync()
        dt = time.time() - t0
        qr, jpvt, tau, _, info = lapack.dgeqp3(np.asfortranarray(Y))
        got = jp.cpu().numpy().astype(int)
        Rg = np.triu(D.to_numpy(Yd)[:min(m, n), :])
        Rr = np.triu(qr[:min(m, n), :])
        same = (got == jpvt - 1)
        first_bad = int(np.argmin(same)) if not same.all() else -1
        print("geqp3 %dx%d: pivots equal=%s (first mismatch at %d)  max|abs(R)-abs(Rref)|=%.2e  %.1f ms"
              % (m, n, same.all(), first_bad, np.abs(np.abs(Rg) - np.abs(Rr)).max() if same.all() else float("nan"), dt * 1e3))


@section("svd")
def _svd():
    api = pkg.Api(32)
    L = ref_lib.RefLib(32)
    for (m, n, k, p, vnum, q, s, spec) in [(2000, 1500, 100, 20, 1, 2, 1, "gap"), (2000, 1500, 100, 20, 1, 2, 1, "logspace"),
                                           (2000, 1500, 100, 20, 2, 3, 2, "gap"), (600, 900, 30, 10, 1, 1, 1, "exp")]:
        A, sig = O.make_matrix(m, n, spec, seed=0, k=k, tail=1e-7)
        t0 = time.time()
        U, S, V = api.svd_rand(A, k, p, vnum, q, s, seed=777)
        dt = time.time() - t0
        t0 = time.time()
        Ur, Sr, Vr = L.svd_rand(A, k, p, vnum, q, s, seed=777)
        dtr = time.time() - t0
        rel = np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr))
        e, er = np.linalg.norm(A - U @ S @ V.T) / np.linalg.norm(A), np.linalg.norm(A - Ur @ Sr @ Vr.T) / np.linalg.norm(A)
        sv = np.linalg.svd(U.T @ Ur, compute_uv=False)
        print("svd_rand %dx%d k=%d p=%d vnum=%d q=%d s=%d %s: max rel sigma err %.2e  recon %.6e (ref %.6e)  min cos(U,Uref)=%.12f  ours %.3fs ref %.3fs"
              % (m, n, k, p, vnum, q, s, spec, rel, e, er, sv.min(), dt, dtr))


@section("id")
def _id():
    api = pkg.Api(32)
    L = ref_lib.RefLib(32)
    for (m, n, k, p, q, s, spec) in [(2000, 1500, 100, 20, 2, 1, "logspace"), (800, 1200, 40, 10, 1, 2, "exp")]:
        A, sig = O.make_matrix(m, n, spec, seed=1)
        Ic, Ir, T, S = api.id_two_sided_rand(A, k, p, q, s, seed=777)
        Icr, Irr, Tr, Sr = L.id_two_sided_rand(A, k, p, q, s, seed=777)
        print("id_two_sided %dx%d k=%d: Icol equal=%s (first k equal=%s) Irow equal=%s  max|T-Tref|=%.2e max|S-Sref|=%.2e"
              % (m, n, k, np.array_equal(Ic, Icr), np.array_equal(Ic[:k], Icr[:k]), np.array_equal(Ir, Irr),
                 np.abs(T - Tr).max() if T.shape == Tr.shape else -1, np.abs(S - Sr).max() if S.shape == Sr.shape else -1))
        Cm, U, R = api.cur_rand(A, k, p, q, s, seed=777)
        Cr, Ur, Rr = L.cur_rand(A, k, p, q, s, seed=777)
        print("cur %dx%d k=%d: C equal=%s R equal=%s  max|U-Uref|/max|Uref|=%.2e  err %.4f%% (ref %.4f%%)"
              % (m, n, k, np.array_equal(Cm, Cr), np.array_equal(R, Rr), np.abs(U - Ur).max() / np.abs(Ur).max(),
                 O.get_percent_error_between_two_mats(A, Cm @ U @ R), O.get_percent_error_between_two_mats(A, Cr @ Ur @ Rr)))


@section("qb")
def _qb():
    api = pkg.Api(32)
    L = ref_lib.RefLib(32)
    A, sig = O.make_matrix(1200, 900, "logspace", seed=2)
    f, Q, B = api.randQB_pb_new(A, 40, 5, 0.0, 2, 1, seed=777)
    fr, Qr, Br = L.randQB_pb_new(A, 40, 5, 0.0, 2, 1, seed=777)
    print("randQB rank mode: frank %d (ref %d)  ||QB-QrBr||/||A||=%.2e  ||A-QB||/||A||=%.6e (ref %.6e) ||QtQ-I||=%.1e"
          % (f, fr, np.linalg.norm(Q @ B - Qr @ Br) / np.linalg.norm(A), np.linalg.norm(A - Q @ B) / np.linalg.norm(A),
             np.linalg.norm(A - Qr @ Br) / np.linalg.norm(A), np.abs(Q.T @ Q - np.eye(Q.shape[1])).max()))
    f, Q, B = api.randQB_pb_new(A, 40, 0, 20.0, 1, 1, seed=777)
    fr, Qr, Br = L.randQB_pb_new(A, 40, 0, 20.0, 1, 1, seed=777)
    print("randQB tol mode: frank %d (ref %d) shapes %s %s  ||A-QB||_F=%.4f (ref %.4f)" % (f, fr, Q.shape, B.shape, np.linalg.norm(A - Q @ B), np.linalg.norm(A - Qr @ Br)))
    f, U, S, V = api.svd_blockrand(A, 100, 20, 0.0, 1, 40, 2, 1, seed=777)
    fr, Ur, Sr, Vr = L.svd_blockrand(A, 100, 20, 0.0, 1, 40, 2, 1, seed=777)
    print("svd_blockrand: frank %d (ref %d)  max rel sigma err %.2e" % (f, fr, np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr))))
    f, U, S, V = api.svd_blockrand(A, 0, 20, 1.0, 1, 40, 2, 1, seed=777)
    fr, Ur, Sr, Vr = L.svd_blockrand(A, 0, 20, 1.0, 1, 40, 2, 1, seed=777)
    print("svd_blockrand k=0 (quirk Q1): frank %d (ref %d)  max rel sigma err %.2e" % (f, fr, np.max(np.abs(np.diag(S) - np.diag(Sr)) / np.diag(Sr))))


@section("phases")
def _phases():
    """phase breakdown of the BASELINE configs[1] step (device resident) and of the host API call"""
    import ctypes as C
    m, n, k, p = 50000, 20000, 500, 20
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((m, 640), dtype=torch.float64, device="cuda", generator=gen) / m ** 0.5
    W = torch.randn((n, 640), dtype=torch.float64, device="cuda", generator=gen) / n ** 0.5
    A_cm = torch.matmul(W * torch.logspace(1, -3, 640, dtype=torch.float64, device="cuda"), X.t())
    A_cm += 1e-6 * torch.randn((n, m), dtype=torch.float64, device="cuda", generator=gen)
    del X, W
    torch.cuda.synchronize()
    for jt in (0, 1):
        lib.rsvd_b200_set_option(b"jacobi_transpose", jt)
        D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777)
        sync()
        lib.rsvd_b200_set_option(b"verbose", 2)
        print("--- device-resident step, jacobi_transpose=%d" % jt, flush=True)
        U, S, V = D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777)
        sync()
        lib.rsvd_b200_set_option(b"verbose", 0)
        t = timeit(lambda: D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777), reps=3, warm=0)
        print("    whole step: %.2f ms  S[0]=%.6f S[-1]=%.6e" % (t * 1e3, S[0].item(), S[-1].item()), flush=True)
    lib.rsvd_b200_set_option(b"jacobi_transpose", 0)
    del A_cm
    torch.cuda.empty_cache()
    os.environ["RSVD_B200_VERBOSE"] = "1"
    api = pkg.Api(32)
    M = api.lib.matrix_new(m, n)
    hA = np.ctypeslib.as_array(M.contents.d, shape=(n, m))
    rng = np.random.default_rng(0)
    np.matmul(rng.standard_normal((n, 32)), rng.standard_normal((32, m)), out=hA)
    for it in range(3):
        Um, Sm, Vm = api.PM(), api.PM(), api.PM()
        fr = api.I(0)
        t0 = time.time()
        api.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, 1, 2, 1, C.byref(fr), C.byref(Um), C.byref(Sm), C.byref(Vm))
        print("    API call wall: %.3f s" % (time.time() - t0), flush=True)
        for x in (Um, Sm, Vm):
            api.lib.matrix_delete(x)
    api.lib.matrix_delete(M)


print("\nlaunches total:", lib.rsvd_b200_launch_count(), " status:", lib.rsvd_b200_status(), lib.rsvd_b200_last_error())
