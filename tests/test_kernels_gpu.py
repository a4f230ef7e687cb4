"""GPU tests of the individual kernel families through the device-layer C-ABI (device pointers), against torch FP64
(cuBLAS/cuSOLVER are used ONLY here, as the checker) and scipy LAPACK.  Also size-independent properties at a size
the oracle cannot run in seconds."""
import numpy as np
import pytest
import torch

from lowrankmatrixdecompositioncodes_b200 import device as D, native
from oracle import ref_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    l = native.dev()
    assert l.rsvd_b200_init(0) == 0, l.rsvd_b200_last_error().decode()
    return l


def sync(lib):
    lib.rsvd_b200_sync()
    torch.cuda.synchronize()


def run_gemm(lib, ta, tb, m, n, k, alpha, beta, pad, force_generic):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k)
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + pad, br + pad, m + pad
    A = torch.randn((ac, lda), dtype=torch.float64, generator=g).cuda()
    B = torch.randn((bc, ldb), dtype=torch.float64, generator=g).cuda()
    Cm = torch.randn((n, ldc), dtype=torch.float64, generator=g).cuda()
    C0 = Cm.clone()
    lib.rsvd_b200_set_option(b"force_generic_gemm", int(force_generic))
    D.gemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha, beta)
    sync(lib)
    path = lib.rsvd_b200_get_option(b"last_gemm_path")
    lib.rsvd_b200_set_option(b"force_generic_gemm", 0)
    opA = A[:, :ar].t() if ta == "N" else A[:, :ar]
    opB = B[:, :br].t() if tb == "N" else B[:, :br]
    ref = alpha * (opA @ opB) + beta * C0[:, :m].t()
    err = ((Cm[:, :m].t() - ref).abs().max() / ref.abs().max()).item()
    assert torch.equal(Cm[:, m:], C0[:, m:])   # padding rows untouched
    return err, path


@pytest.mark.parametrize("ta,tb,m,n,k", [("N", "N", 100, 70, 50), ("T", "N", 129, 65, 1000), ("N", "T", 64, 64, 16),
                                         ("T", "T", 33, 17, 9), ("T", "N", 120, 120, 5000), ("N", "N", 1, 1, 1)])
def test_generic_gemm(lib, ta, tb, m, n, k):
    err, path = run_gemm(lib, ta, tb, m, n, k, 1.3, 0.7, pad=1, force_generic=True)
    assert path == 0 and err < 1e-13


@pytest.mark.parametrize("ta,m,n,k,alpha,beta", [("N", 1024, 256, 512, 1.0, 0.0), ("T", 1024, 256, 512, 1.0, 0.0),
                                                 ("N", 1000, 130, 530, 1.5, 0.5), ("T", 778, 120, 2050, -1.0, 1.0),
                                                 ("T", 520, 520, 50000, 1.0, 0.0), ("N", 5000, 200, 200, -1.0, 1.0),
                                                 ("N", 4112, 520, 3000, 1.0, 0.0), ("N", 2050, 8, 700, 1.0, 0.0),
                                                 ("T", 300, 1050, 4000, 1.0, 0.0),
                                                 # every instantiated tile width: NB = 4, 8, 8, 14, 14 (two tiles), 15, 16
                                                 ("N", 4000, 24, 600, 1.0, 0.0), ("T", 3000, 40, 600, 1.0, 0.0), ("N", 2000, 64, 500, 2.0, 1.0),
                                                 ("N", 1200, 112, 500, 1.0, 0.0), ("T", 600, 216, 700, 1.0, 0.0), ("T", 1000, 120, 640, 1.0, 0.0),
                                                 ("N", 1000, 128, 512, 1.0, 0.0)])
def test_tma_gemm(lib, ta, m, n, k, alpha, beta):
    err, path = run_gemm(lib, ta, "N", m, n, k, alpha, beta, pad=0, force_generic=False)
    assert err < 2e-13
    if n >= 16:
        assert path == 1   # the TMA kernel served it (no silent fallback to the generic kernel)


@pytest.mark.parametrize("ta,m,n,k,sk,sc,off", [("N", 2000, 120, 1500, 1, 1500, 0), ("T", 1500, 120, 2000, 120, 1, 240),
                                                ("N", 4096, 520, 2051, 1, 2051, 5), ("N", 3000, 200, 1000, 1, 1000, 200 * 1000 * 3)])
def test_fused_philox_sketch_equals_explicit_omega(lib, ta, m, n, k, sk, sc, off):
    g = torch.Generator(device="cpu").manual_seed(1)
    ar, ac = (m, k) if ta == "N" else (k, m)
    A = torch.randn((ac, ar), dtype=torch.float64, generator=g).cuda()
    idx = off + np.arange(k)[:, None] * sk + np.arange(n)[None, :] * sc
    flat = ref_lib.normal_stream(99, 0, int(idx.max()) + 1)
    Om = torch.from_numpy(flat[idx]).cuda()
    ref = (A.t() if ta == "N" else A) @ Om
    for force in (0, 1):
        Cm = torch.empty((n, m), dtype=torch.float64, device="cuda")
        lib.rsvd_b200_set_option(b"force_generic_gemm", force)
        native.check(lib.rsvd_b200_sketch(ta.encode(), m, n, k, A.data_ptr(), ar, 99, sk, sc, off, Cm.data_ptr(), m))
        sync(lib)
        assert lib.rsvd_b200_get_option(b"last_gemm_path") == 1 - force
        lib.rsvd_b200_set_option(b"force_generic_gemm", 0)
        assert ((Cm.t() - ref).abs().max() / ref.abs().max()).item() < 1e-13


def test_gemm_linearity_at_scale(lib):
    """size-independent property at a size the CPU oracle cannot do in seconds: A(x+y) = Ax + Ay, (A^T)(Ax) symmetric."""
    m, n, l = 40000, 8192, 264
    A = torch.randn((n, m), dtype=torch.float64, device="cuda")
    X = torch.randn((l, n), dtype=torch.float64, device="cuda")
    Y = torch.randn((l, n), dtype=torch.float64, device="cuda")
    out = [torch.empty((l, m), dtype=torch.float64, device="cuda") for _ in range(3)]
    XY = X + Y
    torch.cuda.synchronize()    # inputs are produced on torch's stream; the library runs on its own stream
    for B, Cm in zip((X, Y, XY), out):
        D.gemm("N", "N", m, l, n, A, m, B, n, Cm, m)
    sync(lib)
    assert ((out[0] + out[1] - out[2]).abs().max() / out[2].abs().max()).item() < 1e-13
    G = torch.empty((l, l), dtype=torch.float64, device="cuda")
    D.gemm("T", "N", l, l, m, out[0], m, out[0], m, G, l)
    sync(lib)
    assert ((G - G.t()).abs().max() / G.abs().max()).item() < 1e-13


@pytest.mark.parametrize("m,l,cond,force", [(5000, 120, 1e3, 0), (5000, 120, 1e12, 0), (3000, 200, 1e2, 1), (20000, 520, 1e5, 0),
                                            (64, 64, 10, 0), (1000, 1, 1, 0), (5000, 120, 1e12, 1), (30000, 200, 1e9, 0), (4000, 300, 1e14, 0),
                                            (3000, 200, 1e2, 2)])
def test_orthonormalize(lib, m, l, cond, force):
    rng = np.random.default_rng(3)
    Q0, _ = np.linalg.qr(rng.standard_normal((m, l)))
    W, _ = np.linalg.qr(rng.standard_normal((l, l)))
    Y = (Q0 * np.logspace(0, -np.log10(cond), l)) @ W.T
    Yd = D.from_numpy_cm(Y)
    R = torch.zeros((l, l), dtype=torch.float64, device="cuda")
    lib.rsvd_b200_set_option(b"force_qr_fallback", force)
    native.check(lib.rsvd_b200_orthonormalize(Yd.data_ptr(), m, m, l, R.data_ptr(), l))
    sync(lib)
    lib.rsvd_b200_set_option(b"force_qr_fallback", 0)
    path = lib.rsvd_b200_get_option(b"last_qr_path")
    # 1 = CholeskyQR2, 4 = shifted CholeskyQR3 (cond beyond the plain pass's limit, or forced with 2), 2 = TSQR-preconditioned (forced with 1)
    assert path == (2 if force == 1 else 4 if (force == 2 or cond > 1e8) else 1)
    Q, Rn = D.to_numpy(Yd), R.t().cpu().numpy()
    assert np.abs(Q.T @ Q - np.eye(l)).max() < 1e-13
    assert np.linalg.norm(Q @ Rn - Y) / np.linalg.norm(Y) < 1e-13
    assert np.abs(np.tril(Rn, -1)).max() == 0.0
    # same range as Householder QR
    Qh, _ = np.linalg.qr(Y)
    if cond < 1e10:
        assert np.linalg.norm(Q - Qh @ (Qh.T @ Q)) < 1e-6


@pytest.mark.parametrize("m,l,rank", [(3000, 64, 10), (500, 40, 1), (2000, 130, 129), (64, 8, 0)])
def test_orthonormalize_exactly_rank_deficient_panel(lib, m, l, rank):
    """A sketch with more columns than the matrix has rank (the reference's dgeqrf + dorgqr, MVF:1251-1263, returns an
    orthonormal completion there): Q must be orthonormal, Q R = Y, R upper triangular — never an error."""
    rng = np.random.default_rng(m + l)
    Y = rng.standard_normal((m, rank)) @ rng.standard_normal((rank, l)) if rank else np.zeros((m, l))
    Yd = D.from_numpy_cm(Y)
    R = torch.zeros((l, l), dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_orthonormalize(Yd.data_ptr(), m, m, l, R.data_ptr(), l))
    sync(lib)
    Q, Rn = D.to_numpy(Yd), R.t().cpu().numpy()
    assert np.all(np.isfinite(Q)) and np.all(np.isfinite(Rn))
    assert np.abs(Q.T @ Q - np.eye(l)).max() < 1e-12
    assert np.linalg.norm(Q @ Rn - Y) <= 1e-12 * max(np.linalg.norm(Y), 1.0)
    assert np.abs(np.tril(Rn, -1)).max() == 0.0


@pytest.mark.parametrize("n", [1, 2, 5, 33, 120, 520, 1500, 2500])
def test_jacobi_svd_and_eig(lib, n):
    rng = np.random.default_rng(n)
    U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s0 = np.logspace(0, -5, n) if n > 1 else np.array([2.5])
    A = (U0 * s0) @ V0.T
    Ad = D.from_numpy_cm(A)
    U = torch.empty((n, n), dtype=torch.float64, device="cuda")
    Vt = torch.empty((n, n), dtype=torch.float64, device="cuda")
    s = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
    sync(lib)
    sn, Un, Vtn = s.cpu().numpy(), U.t().cpu().numpy(), Vt.t().cpu().numpy()
    assert np.max(np.abs(sn - s0) / s0) < 1e-10
    assert np.all(np.diff(sn) <= 0)
    assert np.linalg.norm((Un * sn) @ Vtn - A) / np.linalg.norm(A) < 1e-12
    assert np.abs(Un.T @ Un - np.eye(n)).max() < 1e-12 and np.abs(Vtn @ Vtn.T - np.eye(n)).max() < 1e-12
    # symmetric PSD eigenproblem (B B^T of the vnum=2 branch): ascending eigenvalues like dsyev
    S = (V0 * s0 ** 2) @ V0.T
    S = (S + S.T) / 2
    Sd = D.from_numpy_cm(S)
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_eig_small(Sd.data_ptr(), n, n, w.data_ptr()))
    sync(lib)
    wn, Vn = w.cpu().numpy(), D.to_numpy(Sd)
    assert np.all(np.diff(wn) >= 0)
    assert np.max(np.abs(wn[::-1] - s0 ** 2) / (s0 ** 2).max()) < 1e-12
    assert np.linalg.norm(S @ Vn - Vn * wn) / np.linalg.norm(S) < 1e-12


@pytest.mark.parametrize("n", [2, 64, 130, 520, 592, 700])
def test_jacobi_live_replay_is_bitwise_the_after_the_fact_replay(lib, n):
    """V rebuilt by the replay kernel that runs NEXT TO the Jacobi kernel (side stream, follows the rotation log as it is
    written; n <= 4 * SMs) equals the replay launched after the Jacobi kernel has finished, bit for bit."""
    rng = np.random.default_rng(100 + n)
    U0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    R = np.triu(np.linalg.qr((U0 * np.logspace(1, -2.5, n)) @ V0.T)[1])
    out = []
    for no_live in (0, 1):
        lib.rsvd_b200_set_option(b"no_live_replay", no_live)
        Ad = D.from_numpy_cm(R)
        U = torch.empty((n, n), dtype=torch.float64, device="cuda")
        Vt = torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda")
        s = torch.empty(n, dtype=torch.float64, device="cuda")
        native.check(lib.rsvd_b200_svd_small(Ad.data_ptr(), n, n, U.data_ptr(), n, s.data_ptr(), Vt.data_ptr(), n))
        sync(lib)
        out.append((U.t().cpu().numpy(), s.cpu().numpy(), Vt.t().cpu().numpy()))
    lib.rsvd_b200_set_option(b"no_live_replay", 0)
    (U1, s1, V1), (U2, s2, V2) = out
    assert np.array_equal(V1, V2) and np.array_equal(s1, s2) and np.array_equal(U1, U2)
    assert np.linalg.norm((U1 * s1) @ V1 - R) / np.linalg.norm(R) < 1e-13
    assert np.abs(V1 @ V1.T - np.eye(n)).max() < 1e-13


@pytest.mark.parametrize("n,cond", [(1, 1.0), (5, 10.0), (32, 1e3), (33, 1e3), (100, 1e3), (520, 1e3), (520, 1e6), (545, 1e2), (1050, 1e3),
                                    (2100, 1e2)])
def test_cholesky_dataflow_kernel(lib, n, cond):
    """cholinv.cu: G = R^T R and R^{-1} in one dataflow kernel (32 x 32 block tasks) against numpy, next to the per-block launch
    sequence it replaces.  Only the upper triangle of G may be read (the lower one is NaN here); both outputs carry exact
    zeros below the diagonal (the streaming GEMM's triangular hint relies on them)."""
    import ctypes as C
    rng = np.random.default_rng(n)
    Y = rng.standard_normal((max(2 * n, 64), n)) * np.logspace(0, -np.log10(cond), n)
    G = Y.T @ Y
    Rref = np.linalg.cholesky(G).T
    for no_dataflow in (0, 1):
        lib.rsvd_b200_set_option(b"no_chol_dataflow", no_dataflow)
        Gd = D.from_numpy_cm(np.triu(G) + np.tril(np.full((n, n), np.nan), -1))
        Xd = torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda")
        mm = (C.c_double * 2)()
        info = lib.rsvd_b200_chol_inv(D.ptr(Gd), n, n, D.ptr(Xd), n, mm)
        sync(lib)
        R, X = D.to_numpy(Gd), D.to_numpy(Xd)
        assert info == 0
        assert np.all(np.tril(R, -1) == 0) and np.all(np.tril(X, -1) == 0)
        assert np.abs(R - Rref).max() / np.abs(Rref).max() < 1e-12 * max(1.0, cond / 1e3)
        assert np.abs(R @ X - np.eye(n)).max() < 1e-12 * cond
        if not no_dataflow:
            d = np.diag(Rref)
            assert abs(mm[0] - d.min()) <= 1e-10 * d.max() and abs(mm[1] - d.max()) <= 1e-10 * d.max()
    lib.rsvd_b200_set_option(b"no_chol_dataflow", 0)


def test_cholesky_dataflow_kernel_reports_the_failing_column(lib):
    n = 200
    A = np.eye(n)
    A[150, 150] = -1.0
    for no_dataflow in (0, 1):
        lib.rsvd_b200_set_option(b"no_chol_dataflow", no_dataflow)
        Gd, Xd = D.from_numpy_cm(A), D.new_cm(n, n)
        assert lib.rsvd_b200_chol_inv(D.ptr(Gd), n, n, D.ptr(Xd), n, None) == 151
    lib.rsvd_b200_set_option(b"no_chol_dataflow", 0)
    native.check(lib.rsvd_b200_sync())


@pytest.mark.parametrize("m,n", [(12, 40), (120, 1500), (100, 2000), (300, 5000), (40, 40), (7, 3), (200, 9000), (1500, 400), (2500, 2500)])
def test_geqp3_pivots_bit_exact_vs_lapack(lib, m, n):
    from scipy.linalg import lapack
    rng = np.random.default_rng(m + n)
    r = min(m, n, 60)
    Y = rng.standard_normal((m, r)) @ (np.logspace(0, -5, r)[:, None] * rng.standard_normal((r, n)))
    Y = Y + 1e-9 * rng.standard_normal((m, n))
    Yd = D.from_numpy_cm(Y)
    jp = torch.empty(n, dtype=torch.float64, device="cuda")
    native.check(lib.rsvd_b200_geqp3(Yd.data_ptr(), m, m, n, jp.data_ptr()))
    sync(lib)
    qr, jpvt, tau, _, info = lapack.dgeqp3(np.asfortranarray(Y))
    assert np.array_equal(jp.cpu().numpy().astype(int), jpvt - 1)
    k = min(m, n)
    Rg, Rr = np.triu(D.to_numpy(Yd)[:k, :]), np.triu(qr[:k, :])
    assert np.abs(np.abs(Rg) - np.abs(Rr)).max() < 1e-11 * np.abs(Rr).max()
    d = np.abs(np.diag(Rg))
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-12))     # non-increasing |R_ii|: the pivoting property


@pytest.mark.parametrize("m,n", [(300, 5000), (5000, 300), (64, 64)])
def test_geqp3_unblocked_kernel_still_matches_lapack(lib, m, n):
    """the one-reflector-per-step kernel serves tall inputs (> 4096 rows) and stays selectable for the others"""
    from scipy.linalg import lapack
    rng = np.random.default_rng(3 * m + n)
    r = min(m, n, 50)
    Y = rng.standard_normal((m, r)) @ (np.logspace(0, -5, r)[:, None] * rng.standard_normal((r, n))) + 1e-9 * rng.standard_normal((m, n))
    Yd = D.from_numpy_cm(Y)
    jp = torch.empty(n, dtype=torch.float64, device="cuda")
    lib.rsvd_b200_set_option(b"force_unblocked_qr", 1)
    try:
        native.check(lib.rsvd_b200_geqp3(Yd.data_ptr(), m, m, n, jp.data_ptr()))
        sync(lib)
    finally:
        lib.rsvd_b200_set_option(b"force_unblocked_qr", 0)
    qr, jpvt, tau, _, info = lapack.dgeqp3(np.asfortranarray(Y))
    assert np.array_equal(jp.cpu().numpy().astype(int), jpvt - 1)


def test_geqp3_blocked_kernel_on_a_noise_dominated_square_matrix(lib):
    """2500 x 2500, numerical rank 60 over a 1e-9 noise floor: after 60 steps every remaining column norm is noise whose computed
    value carries ~1e-7 relative rounding error (cancellation against O(1) entries), so the ORDER of near-tied pivots is not
    defined to the last bit by the algorithm — blocked (dlaqps-style) and unblocked arithmetic legitimately differ there.  What is
    defined: the leading pivots, |diag R| (non-increasing, equal to LAPACK's to the tie tolerance) and a valid permutation."""
    from scipy.linalg import lapack
    m = n = 2500
    rng = np.random.default_rng(m + n)
    r = 60
    Y = rng.standard_normal((m, r)) @ (np.logspace(0, -5, r)[:, None] * rng.standard_normal((r, n))) + 1e-9 * rng.standard_normal((m, n))
    Yd = D.from_numpy_cm(Y)
    jp = torch.empty(n, dtype=torch.float64, device="cuda")
    lib.rsvd_b200_set_option(b"qr_blocked_rows", 4096)
    try:
        native.check(lib.rsvd_b200_geqp3(Yd.data_ptr(), m, m, n, jp.data_ptr()))
        sync(lib)
    finally:
        lib.rsvd_b200_set_option(b"qr_blocked_rows", 2048)
    qr, jpvt, _, _, _ = lapack.dgeqp3(np.asfortranarray(Y))
    ours, ref = jp.cpu().numpy().astype(int), jpvt - 1
    assert np.array_equal(ours[:r], ref[:r]) and sorted(ours) == list(range(n))
    d, dr = np.abs(np.diag(D.to_numpy(Yd))), np.abs(np.diag(qr))
    assert np.all(d[:-1] >= d[1:] * (1 - 1e-6))
    assert np.max(np.abs(d - dr)[:-1] / dr[:-1]) < 5e-4          # near-tied pivots may swap; their |R_ii| agree to the tie width
    nd = int(np.argmax(ours != ref)) if not np.array_equal(ours, ref) else n
    print("first pivot difference at position %d of %d" % (nd, n))


def test_geqp3_blocked_edge_cases(lib):
    """exact ties (duplicated columns: first-index rule), zero columns, exact rank deficiency, degenerate shapes, and a column
    count that leaves ragged CTAs; the safeguard recompute is exercised by the 1e-9 floor under a 1e5 dynamic range"""
    from scipy.linalg import lapack
    rng = np.random.default_rng(11)

    def run(Y):
        m, n = Y.shape
        Yd = D.from_numpy_cm(Y)
        jp = torch.empty(n, dtype=torch.float64, device="cuda")
        native.check(lib.rsvd_b200_geqp3(Yd.data_ptr(), m, m, n, jp.data_ptr()))
        sync(lib)
        return jp.cpu().numpy().astype(int), D.to_numpy(Yd)

    # duplicated and zero columns
    Y = rng.standard_normal((40, 300)) * np.logspace(0, -3, 300)
    Y[:, 17] = Y[:, 5]; Y[:, 250] = Y[:, 5]; Y[:, 100] = 0.0; Y[:, 0] = 0.0
    jp, R = run(Y)
    qr, jpvt, _, _, _ = lapack.dgeqp3(np.asfortranarray(Y))
    assert np.array_equal(jp, jpvt - 1)
    # exactly rank 6: the first 6 pivots are determined, the rest is rounding noise in ANY implementation
    Y = rng.standard_normal((50, 6)) @ rng.standard_normal((6, 777))
    jp, R = run(Y)
    qr, jpvt, _, _, _ = lapack.dgeqp3(np.asfortranarray(Y))
    assert np.array_equal(jp[:6], (jpvt - 1)[:6]) and sorted(jp) == list(range(777))
    assert np.abs(np.diag(R)[6:50]).max() < 1e-12 * abs(R[0, 0])
    # degenerate shapes
    for shape in [(1, 1), (1, 9), (9, 1), (33, 33), (2, 4097)]:
        Y = rng.standard_normal(shape)
        jp, R = run(Y)
        qr, jpvt, _, _, _ = lapack.dgeqp3(np.asfortranarray(Y))
        assert np.array_equal(jp, jpvt - 1), shape
        k = min(shape)
        assert np.abs(np.abs(np.triu(R[:k])) - np.abs(np.triu(qr[:k]))).max() < 1e-12 * np.abs(qr).max(), shape


def test_geqp3_explicit_q_is_orthonormal_for_any_input(lib):
    """dgeqp3 + dorgqr (pivotedQR_mkl, RRA:924-976): Q comes from the reflectors, so it is orthonormal for rank-deficient and
    ill-conditioned inputs too, and Q R reproduces the permuted matrix"""
    rng = np.random.default_rng(5)
    for name, Y in [("rank-deficient", rng.standard_normal((300, 9)) @ rng.standard_normal((9, 120))),
                    ("cond 1e12", (np.linalg.qr(rng.standard_normal((400, 80)))[0] * np.logspace(0, -12, 80)) @ np.linalg.qr(rng.standard_normal((80, 80)))[0]),
                    ("wide", rng.standard_normal((60, 500))), ("tall > 4096 rows", rng.standard_normal((5000, 40)))]:
        m, n = Y.shape
        k = min(m, n)
        Yd, Qd = D.from_numpy_cm(Y), D.new_cm(m, k)
        jp = torch.empty(n, dtype=torch.float64, device="cuda")
        native.check(lib.rsvd_b200_geqp3_q(Yd.data_ptr(), m, m, n, jp.data_ptr(), Qd.data_ptr(), m))
        sync(lib)
        Q, R, I = D.to_numpy(Qd), np.triu(D.to_numpy(Yd)[:k, :]), jp.cpu().numpy().astype(int)
        assert np.abs(Q.T @ Q - np.eye(k)).max() < 1e-13, name
        assert np.linalg.norm(Q @ R - Y[:, I]) <= 1e-13 * np.linalg.norm(Y), name


def test_torch_wrappers_are_stream_ordered_without_host_sync(lib):
    """device.gemm / device.svd_rand order the library's stream against torch's current stream with events: inputs produced
    asynchronously by torch kernels and outputs consumed by torch kernels need no torch.cuda.synchronize() in between"""
    m, n, k = 4096, 512, 2048
    for rep in range(3):
        g = torch.Generator(device="cuda").manual_seed(rep)
        big = torch.randn((6000, 6000), dtype=torch.float64, device="cuda", generator=g)
        _ = big @ big                                                    # keeps torch's stream busy while the inputs are produced
        A = (torch.randn((k, m), dtype=torch.float64, device="cuda", generator=g) * 3.0).contiguous()     # column-major m x k
        B = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=g).contiguous()             # column-major k x n
        Cm = torch.full((n, m), 7.0, dtype=torch.float64, device="cuda")
        ref = Cm.clone()                                                 # beta != 0: the old contents are an input too
        D.gemm("N", "N", m, n, k, A, m, B, k, Cm, m, alpha=1.0, beta=0.5)
        got = Cm.clone()                                                 # consumer on torch's stream, right away
        want = (B @ A) + 0.5 * ref                                       # (A_cm B_cm)^T in torch's row-major view
        assert (got - want).abs().max().item() <= 1e-9 * want.abs().max().item()


def test_trsm_and_lu_solve(lib):
    rng = np.random.default_rng(0)
    k, nc = 300, 2000
    R = np.triu(rng.standard_normal((k, k))) + 5 * np.eye(k)
    B = rng.standard_normal((k, nc))
    Rd, Bd = D.from_numpy_cm(R), D.from_numpy_cm(B)
    native.check(lib.rsvd_b200_trsm_left_upper(Rd.data_ptr(), k, k, Bd.data_ptr(), k, nc))
    sync(lib)
    X = D.to_numpy(Bd)
    assert np.linalg.norm(R @ X - B) / np.linalg.norm(B) < 1e-12
    A = rng.standard_normal((k, k))
    Ad, Bd = D.from_numpy_cm(A), D.from_numpy_cm(B[:, :k])
    native.check(lib.rsvd_b200_lu_solve(Ad.data_ptr(), k, k, Bd.data_ptr(), k, k))
    sync(lib)
    X = D.to_numpy(Bd)
    assert np.linalg.norm(A @ X - B[:, :k]) / np.linalg.norm(B[:, :k]) < 1e-11


def test_device_resident_svd_roundtrip_at_scale(lib):
    """Device-resident entry point on a matrix generated in HBM with a KNOWN spectrum (no CPU oracle at this size):
    sigma must match the construction, U/V orthonormal, streamed residual ~ the discarded tail."""
    m, n, k, p = 30000, 12000, 200, 20
    g = torch.Generator(device="cuda").manual_seed(0)
    X, _ = torch.linalg.qr(torch.randn((m, k + 50), dtype=torch.float64, device="cuda", generator=g))
    W, _ = torch.linalg.qr(torch.randn((n, k + 50), dtype=torch.float64, device="cuda", generator=g))
    sig = torch.cat([torch.logspace(0, -4, k, dtype=torch.float64), 1e-9 * torch.ones(50, dtype=torch.float64)]).cuda()
    A_cm = ((X * sig) @ W.t()).t().contiguous()     # (n, m) tensor == column-major m x n
    torch.cuda.synchronize()
    U, S, V = D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777)
    pe = lib.rsvd_b200_svd_percent_error_dev(A_cm.data_ptr(), m, n, m, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n, k)
    sync(lib)
    assert ((S - sig[:k]).abs() / sig[:k]).max().item() < 1e-9
    assert (U @ U.t() - torch.eye(k, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-11
    assert (V @ V.t() - torch.eye(k, dtype=torch.float64, device="cuda")).abs().max().item() < 1e-11
    expect = 100 * (sig[k:].norm() / sig.norm()).item()
    assert pe == pytest.approx(expect, rel=0.05)
