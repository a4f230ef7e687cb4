"""GPU parity sweep of the device-backed helpers of matrix_vector_functions_intel_mkl.h (BLAS/LAPACK-class operations the
reference hands to MKL: MVF:538-561, 1206-1284, 1477-1531, 458-486) against the compiled reference on the same inputs."""
import ctypes as C

import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native
from oracle import ref_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def both():
    if not ref_lib.available(32):
        pytest.skip("compiled reference not present")
    lib = native.dev()
    assert lib.rsvd_b200_init(0) == 0
    return pkg.Api(32), ref_lib.RefLib(32)


def call(side, name, mats, out_shapes, extra=()):
    """fn(in mats..., out mats..., extra ints) with all-mat arguments; returns the output matrices"""
    PM = C.POINTER(side.Mat)
    fn = getattr(side.lib, name)
    ins = [side.to_mat(a) for a in mats]
    outs = [side.lib.matrix_new(*s) for s in out_shapes]
    fn.argtypes = [PM] * (len(ins) + len(outs)) + [C.c_int] * len(extra)
    fn.restype = None
    fn(*ins, *outs, *extra)
    res = [side.from_mat(o) for o in outs]
    res_in = [side.from_mat(i) for i in ins]
    return res, res_in


rng = np.random.default_rng(99)
A = rng.standard_normal((70, 40))
B = rng.standard_normal((40, 25))
T = rng.standard_normal((70, 25))


def rel(X, Y):
    return np.linalg.norm(X - Y) / max(np.linalg.norm(Y), 1e-300)


@pytest.mark.parametrize("name,mats,shape", [
    ("matrix_matrix_mult", (A, B), (70, 25)),
    ("matrix_transpose_matrix_mult", (A, T), (40, 25)),
    ("matrix_matrix_transpose_mult", (T, B), (70, 40)),
])
def test_gemm_helpers(both, name, mats, shape):
    ours, ref = both
    (c0,), _ = call(ours, name, mats, [shape])
    (c1,), _ = call(ref, name, mats, [shape])
    ours.check()
    assert rel(c0, c1) < 1e-14


def test_matrix_vector_helpers(both):
    x, y = rng.standard_normal(40), rng.standard_normal(70)
    res = []
    for side in both:
        PM, PV = C.POINTER(side.Mat), C.POINTER(side.Vec)
        side.lib.vector_new.restype = PV
        side.lib.vector_new.argtypes = [side.I]
        out = []
        for name, v, n_out in (("matrix_vector_mult", x, 70), ("matrix_transpose_vector_mult", y, 40)):
            fn = getattr(side.lib, name)
            fn.argtypes = [PM, PV, PV]
            vi, vo = side.lib.vector_new(len(v)), side.lib.vector_new(n_out)
            np.ctypeslib.as_array(vi.contents.d, shape=(len(v),))[:] = v
            fn(side.to_mat(A), vi, vo)
            out.append(side.from_vec(vo))
        res.append(out)
    assert rel(res[0][0], res[1][0]) < 1e-14 and rel(res[0][1], res[1][1]) < 1e-14
    assert rel(res[0][0], A @ x) < 1e-14


def test_qr_helpers(both):
    ours, ref = both
    (Q0, R0), _ = call(ours, "compact_QR_factorization", (A,), [(70, 40), (40, 40)])
    (Q1, R1), _ = call(ref, "compact_QR_factorization", (A,), [(70, 40), (40, 40)])
    ours.check()
    assert rel(Q0 @ R0, A) < 1e-13 and np.abs(Q0.T @ Q0 - np.eye(40)).max() < 1e-13 and np.abs(np.tril(R0, -1)).max() == 0
    assert rel(np.abs(R0), np.abs(R1)) < 1e-12                     # same factor up to the sign convention of each row
    (Qg,), _ = call(ours, "QR_factorization_getQ", (A,), [(70, 40)])
    assert np.linalg.norm(Qg - Q1 @ (Q1.T @ Qg)) < 1e-12


def test_svd_and_eig_helpers(both):
    ours, ref = both
    for M_ in (rng.standard_normal((30, 30)), A, A.T.copy()):
        U0, S0, V0 = ours.gesvd(M_)
        U1, S1, V1 = ref.gesvd(M_)
        ours.check()
        assert np.allclose(np.diag(S0), np.diag(S1), rtol=1e-12) and rel(U0 @ S0 @ V0, M_) < 1e-13
    Ssym = A.T @ A
    out = []
    for side in both:
        PM, PV = C.POINTER(side.Mat), C.POINTER(side.Vec)
        side.lib.vector_new.restype = PV
        side.lib.vector_new.argtypes = [side.I]
        fn = side.lib.compute_evals_and_evecs_of_symm_matrix
        fn.argtypes = [PM, PV]
        M = side.to_mat(Ssym)
        w = side.lib.vector_new(40)
        fn(M, w)
        out.append((side.from_vec(w), side.from_mat(M)))
    (w0, V0), (w1, V1) = out
    assert np.all(np.diff(w0) >= 0) and np.max(np.abs(w0 - w1)) < 1e-12 * w1.max()
    assert rel(Ssym @ V0, V0 * w0) < 1e-12


def test_solve_helpers(both):
    ours, ref = both
    R = np.triu(rng.standard_normal((40, 40))) + 6 * np.eye(40)
    (X0,), _ = call(ours, "upper_triangular_system_solve", (R, B), [(40, 25)], extra=(1,))
    (X1,), _ = call(ref, "upper_triangular_system_solve", (R, B), [(40, 25)], extra=(1,))
    assert rel(X0, X1) < 1e-12
    _, (Ri0,) = call(ours, "invert_upper_triangular_matrix", (R,), [])
    _, (Ri1,) = call(ref, "invert_upper_triangular_matrix", (R,), [])
    assert rel(np.triu(Ri0), np.triu(Ri1)) < 1e-12
    Asq = rng.standard_normal((40, 40)) + 5 * np.eye(40)
    # square_matrix_system_solve(A, X, B): A X = B by dgesv.  The reference passes ldb = B->ncols (MVF:1528), so it only works
    # for a SQUARE right-hand side (its one call site, the CUR tail RRA:2247-2250, has a k x k one): compare on that shape
    Bsq = rng.standard_normal((40, 40))
    sols = []
    for side in both:
        PM = C.POINTER(side.Mat)
        fn = side.lib.square_matrix_system_solve
        fn.argtypes = [PM, PM, PM]
        X = side.lib.matrix_new(40, 40)
        fn(side.to_mat(Asq), X, side.to_mat(Bsq))
        sols.append(side.from_mat(X))
    ours.check()
    assert rel(sols[0], sols[1]) < 1e-11 and rel(Asq @ sols[0], Bsq) < 1e-12
    (Xr,), _ = call(ours, "upper_triangular_system_solve", (R, B), [(40, 25)], extra=(2,))   # the other solve_type values: same result
    assert rel(Xr, X0) < 1e-12


def test_product_and_random_helpers(both):
    ours, ref = both
    U_, S_, V_ = rng.standard_normal((70, 6)), np.diag(rng.random(6) + 1), rng.standard_normal((40, 6))
    (P0,), _ = call(ours, "form_svd_product_matrix", (U_, S_, V_), [(70, 40)])
    (P1,), _ = call(ref, "form_svd_product_matrix", (U_, S_, V_), [(70, 40)])
    assert rel(P0, P1) < 1e-14
    Cc, Uu, Rr = rng.standard_normal((70, 6)), rng.standard_normal((6, 6)), rng.standard_normal((6, 40))
    (P0,), _ = call(ours, "form_cur_product_matrix", (Cc, Uu, Rr), [(70, 40)])
    (P1,), _ = call(ref, "form_cur_product_matrix", (Cc, Uu, Rr), [(70, 40)])
    assert rel(P0, P1) < 1e-14
    assert np.array_equal(ours.omega(33, 21, seed=4), ref.omega(33, 21, seed=4))
    ours.check()


def test_a_failed_helper_call_does_not_poison_the_next_one(both):
    """The out-of-band error status is sticky inside the device layer; every top-level helper call clears it on entry, so an
    unmodified C driver that never heard of rsvd_b200_api_clear_error() keeps working after one bad call."""
    ours, _ = both
    wide = rng.standard_normal((10, 30))
    (Qbad,), _ = call(ours, "QR_factorization_getQ", (wide,), [(10, 30)])       # m < n: reported, outputs untouched (zeros)
    assert ours.lib.rsvd_b200_api_status() != 0 and np.abs(Qbad).max() == 0
    # no explicit clear: the next helper call starts clean and computes
    (P,), _ = call(ours, "matrix_matrix_mult", (A, B), [(70, 25)])
    assert ours.lib.rsvd_b200_api_status() == 0
    assert rel(P, A @ B) < 1e-14
    # a composite helper keeps ONE status across its nested calls
    S = np.diag(np.arange(1.0, 6.0))
    U5, V5 = np.linalg.qr(rng.standard_normal((20, 5)))[0], np.linalg.qr(rng.standard_normal((12, 5)))[0]
    (P2,), _ = call(ours, "form_svd_product_matrix", (U5, S, V5), [(20, 12)])
    ours.check()
    assert rel(P2, U5 @ S @ V5.T) < 1e-14


def test_allocation_helpers_are_thread_safe(both):
    """Reference drivers call vector_new / vector_delete inside omp parallel loops: the pinned-block tables behind matrix_new
    are guarded, small allocations never touch them."""
    import threading
    ours, _ = both
    errs = []

    def worker(seed):
        try:
            r = np.random.default_rng(seed)
            for _ in range(200):
                n = int(r.integers(1, 2000))
                v = ours.lib.vector_new(n)
                v.contents.d[n - 1] = 1.0
                ours.lib.vector_delete(v)
            for _ in range(3):
                M = ours.lib.matrix_new(3000, 3000)          # 72 MB: pinned path, goes through the tables
                M.contents.d[0] = 2.0
                ours.lib.matrix_delete(M)
        except Exception as e:     # pragma: no cover
            errs.append(e)
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
