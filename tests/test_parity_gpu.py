"""GPU parity tests: the CUDA path, called through the C-ABI exactly as a C driver would (host mat/vec structs),
against (i) the committed golden vectors, (ii) the numpy twin and (iii) the unmodified reference C code
(oracle/_ref) on the same seeded inputs and the same Omega.

Tolerances (BASELINE.json north_star): identical Omega => singular values to 1e-10 relative, subspaces by principal
angle, ID/CUR index sets bit-exact; reconstruction error within 1% of the reference's."""
import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native
from oracle import ref_lib, rsvd_numpy as O
from helpers import subspace_sin, rel_sigma_err, recon_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    lib = native.dev()
    assert lib.rsvd_b200_init(0) == 0, lib.rsvd_b200_last_error().decode()
    return pkg.Api(32)


@pytest.fixture(scope="module")
def api64(api):
    return pkg.Api(64)


@pytest.fixture(scope="module")
def oracle():
    """The checker: compiled reference when present (GPU box gets the prebuilt oracle/_ref), else the numpy twin."""
    if ref_lib.available(32):
        return ref_lib.RefLib(32)

    class Twin:
        svd_rand = staticmethod(lambda A, k, p, vnum=1, q=2, s=1, seed=777: O.low_rank_svd_rand_decomp_fixed_rank(A, k, p, vnum, q, s, seed))
        svd_blockrand = staticmethod(lambda A, k, p, TOL, vnum, kstep, q, s, seed=777: O.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(A, k, p, TOL, vnum, kstep, q, s, seed))
        randQB_pb_new = staticmethod(lambda A, kstep, nstep, TOL, q, s, seed=777: O.randQB_pb_new(A, kstep, nstep, TOL, q, s, seed))
        id_rand = staticmethod(lambda A, k, p, q, s, seed=777: O.id_rand_decomp_fixed_rank(A, k, p, q, s, seed))
        id_two_sided_rand = staticmethod(lambda A, k, p, q, s, seed=777: O.id_two_sided_rand_decomp_fixed_rank(A, k, p, q, s, seed))
        cur_rand = staticmethod(lambda A, k, p, q, s, seed=777: O.cur_rand_decomp_fixed_rank(A, k, p, q, s, seed))
        omega = staticmethod(lambda r, c, seed=777: O.initialize_random_matrix(r, c, seed))
    return Twin()


def test_native_library_is_the_one_running(api):
    lib = native.dev()
    before = lib.rsvd_b200_launch_count()
    api.svd_rand(np.random.default_rng(0).standard_normal((300, 200)), 10, 5)
    assert lib.rsvd_b200_launch_count() > before


def test_omega_bit_exact_device_vs_cpu(api, oracle, golden):
    assert np.array_equal(api.omega(7, 5, seed=777), golden["omega_7x5_seed777"])
    assert np.array_equal(api.omega(301, 77, seed=5), oracle.omega(301, 77, seed=5))


# ---- SVD -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,mat", [("svdA_v1", "A"), ("svdA_v2", "A"), ("svdG_v1", "G")])
def test_svd_vs_golden(api, golden, name, mat):
    k, p, vnum, q, s, seed = [int(x) for x in golden[name + "_params"]]
    A = golden[mat]
    U, S, V = api.svd_rand(A, k, p, vnum, q, s, seed=seed)
    tol = 1e-10 if vnum == 1 else 1e-7   # vnum=2 squares the spectrum (eig of B B^T): eps*(s1/sk)^2
    assert rel_sigma_err(S, golden[name + "_S"]) < tol
    assert subspace_sin(U, golden[name + "_U"]) < 1e-6 and subspace_sin(V, golden[name + "_V"]) < 1e-6
    assert U.shape == (A.shape[0], k) and S.shape == (k, k) and V.shape == (A.shape[1], k)
    d = np.diag(S)
    assert np.all(np.diff(d) <= 0) if vnum == 1 else np.all(np.diff(d) >= 0)   # RRA:171-174 vs RRA:220-223
    assert np.count_nonzero(S - np.diag(d)) == 0                               # S is a full diagonal mat


def test_svd_with_a_sketch_wider_than_2048(api):
    """k + p = 2120: every stage beyond the round-1 limits (one-sided Jacobi above 2048 columns, the dataflow Cholesky kernel with
    67 x 67 blocks, GEMM tiles for 2120 columns) against the numpy twin with the same Omega."""
    m, n, k, p = 3200, 2600, 2100, 20
    rng = np.random.default_rng(12)
    A = (rng.standard_normal((m, 2300)) * np.logspace(0, -5, 2300)) @ rng.standard_normal((2300, n)) / 50.0
    U, S, V = api.svd_rand(A, k, p, 1, 2, 1, seed=777)
    Ur, Sr, Vr = O.low_rank_svd_rand_decomp_fixed_rank(A, k, p, 1, 2, 1, seed=777)
    assert rel_sigma_err(S, Sr) < 1e-10
    e, er = recon_err(A, U, S, V), recon_err(A, Ur, Sr, Vr)
    assert abs(e - er) <= 0.01 * er
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-10 and np.abs(V.T @ V - np.eye(k)).max() < 1e-10


@pytest.mark.parametrize("m,n,k,p,vnum,q,s,spec", [
    (2000, 1500, 100, 20, 1, 2, 1, "logspace"),    # BASELINE config 0 (reference's own spectrum)
    (2000, 1500, 100, 20, 1, 2, 1, "exp"),         # BASELINE config 0 ("exp-decaying spectrum")
    (2000, 1500, 100, 20, 1, 2, 1, "gap"),
    (1500, 2000, 64, 16, 1, 3, 2, "gap"),          # s=2: orthonormalise every other product
    (999, 701, 33, 7, 1, 1, 1, "exp"),             # q=1: no power iteration (RRA:101 `j<q`); odd leading dimensions
    (1200, 900, 50, 10, 2, 2, 1, "logspace"),      # eig of B B^T branch
    (640, 4000, 40, 24, 1, 2, 1, "gap"),           # wide
])
def test_svd_vs_reference_same_omega(api, oracle, m, n, k, p, vnum, q, s, spec):
    A, sig = O.make_matrix(m, n, spec, seed=m + n, k=k, tail=1e-7)
    U, S, V = api.svd_rand(A, k, p, vnum, q, s, seed=31)
    Ur, Sr, Vr = oracle.svd_rand(A, k, p, vnum, q, s, seed=31)
    cond = np.diag(Sr).max() / np.diag(Sr).min()
    tol = 1e-10 if vnum == 1 else max(1e-10, 1e-15 * cond ** 2)
    assert rel_sigma_err(S, Sr) < tol
    e, er = recon_err(A, U, S, V), recon_err(A, Ur, Sr, Vr)
    assert abs(e - er) <= 0.01 * er
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-10 and np.abs(V.T @ V - np.eye(k)).max() < 1e-8
    if spec == "gap":   # well-separated leading subspace: compare by principal angle
        assert subspace_sin(U, Ur) < 1e-6 and subspace_sin(V, Vr) < 1e-6


@pytest.mark.parametrize("m,n,k,p,q", [(10, 8, 3, 2, 2), (64, 64, 20, 44, 1), (33, 500, 16, 17, 2), (500, 33, 30, 3, 3), (257, 129, 1, 0, 2)])
def test_svd_edge_shapes(api, oracle, m, n, k, p, q):
    """tiny / ragged / extreme shapes: k+p == min(m,n), p == 0, k == 1, odd leading dimensions (no TMA alignment),
    matrices smaller than one GEMM tile."""
    rng = np.random.default_rng(m * n)
    A = rng.standard_normal((m, n)) * np.logspace(0, -3, n)[None, :]
    U, S, V = api.svd_rand(A, k, p, 1, q, 1, seed=4)
    Ur, Sr, Vr = oracle.svd_rand(A, k, p, 1, q, 1, seed=4)
    assert U.shape == (m, k) and S.shape == (k, k) and V.shape == (n, k)
    assert rel_sigma_err(S, Sr) < 1e-9
    assert abs(recon_err(A, U, S, V) - recon_err(A, Ur, Sr, Vr)) <= 0.01 * recon_err(A, Ur, Sr, Vr) + 1e-12


def test_id_edge_shapes(api, oracle):
    rng = np.random.default_rng(5)
    for (m, n, k, p) in [(12, 9, 3, 2), (40, 300, 7, 3), (300, 40, 20, 20)]:
        A = rng.standard_normal((m, n)) * np.logspace(0, -2, n)[None, :]
        Ic, Ir, T, S = api.id_two_sided_rand(A, k, p, 1, 1, seed=6)
        Icr, Irr, Tr, Sr = oracle.id_two_sided_rand(A, k, p, 1, 1, seed=6)
        assert np.array_equal(Ic, Icr) and np.array_equal(Ir, Irr)
        assert np.abs(T - Tr).max() < 1e-9 and np.abs(S - Sr).max() < 1e-9


def test_svd_device_philox_different_seed_agrees_on_gap_spectrum(api, oracle):
    """'device Philox' mode of the north star: a different Omega must still give sigma to 1e-8 on a gap matrix."""
    A, sig = O.make_matrix(2000, 1500, "gap", seed=0, k=100, tail=1e-8)
    U, S, V = api.svd_rand(A, 100, 20, 1, 2, 1, seed=1)
    Ur, Sr, Vr = oracle.svd_rand(A, 100, 20, 1, 2, 1, seed=2)
    assert rel_sigma_err(S, Sr) < 1e-8
    assert rel_sigma_err(S, sig[:100]) < 1e-8
    assert abs(recon_err(A, U, S, V) - recon_err(A, Ur, Sr, Vr)) <= 0.01 * recon_err(A, Ur, Sr, Vr)


def test_svd_ill_conditioned_sketch_takes_tsqr_fallback(api, oracle):
    """tail 1e-10 => cond(Y) ~ 1e10: the Gram matrix is not Cholesky-factorable (SURVEY.md hard part 2)."""
    A, sig = O.make_matrix(2000, 1500, "gap", seed=0, k=100, tail=1e-10)
    lib = native.dev()
    before = lib.rsvd_b200_get_option(b"qr_fallbacks")
    U, S, V = api.svd_rand(A, 100, 20, 1, 2, 1, seed=777)
    assert lib.rsvd_b200_get_option(b"qr_fallbacks") > before
    Ur, Sr, Vr = oracle.svd_rand(A, 100, 20, 1, 2, 1, seed=777)
    assert rel_sigma_err(S, Sr) < 1e-10
    assert np.abs(U.T @ U - np.eye(100)).max() < 1e-10


def test_svd_64bit_abi(api64, oracle):
    A, _ = O.make_matrix(700, 500, "gap", seed=9, k=30, tail=1e-7)
    U, S, V = api64.svd_rand(A, 30, 10, 1, 2, 1, seed=3)
    Ur, Sr, Vr = oracle.svd_rand(A, 30, 10, 1, 2, 1, seed=3)
    assert rel_sigma_err(S, Sr) < 1e-10


def test_invalid_parameters_reported_out_of_band(api):
    A = np.random.default_rng(0).standard_normal((50, 40))
    with pytest.raises(RuntimeError):
        api.svd_rand(A, 38, 5)          # k+p > min(m,n): dorgqr/dgesvd would fail silently in the reference (Q7)
    with pytest.raises(RuntimeError):
        api.svd_rand(A, 10, 5, 1, 2, 0)  # s = 0: modulo by zero in the reference (RRA:104)


# ---- blocked QB --------------------------------------------------------------------------------------------------
def test_randqb_vs_golden_and_reference(api, oracle, golden):
    A = golden["A"]
    f, Q, B = api.randQB_pb_new(A, 4, 3, 0.0, 2, 1, seed=777)
    assert f == int(golden["qbA_rank_frank"])
    assert np.allclose(Q @ B, golden["qbA_rank_Q"] @ golden["qbA_rank_B"], atol=1e-11)
    f, Q, B = api.randQB_pb_new(A, 4, 0, 2.0, 1, 1, seed=777)        # tolerance mode, evaluated on the device
    assert f == int(golden["qbA_tol_frank"]) and Q.shape == golden["qbA_tol_Q"].shape and B.shape == golden["qbA_tol_B"].shape
    assert np.allclose(Q @ B, golden["qbA_tol_Q"] @ golden["qbA_tol_B"], atol=1e-11)
    assert np.linalg.norm(A - Q @ B) < 2.0
    A2, _ = O.make_matrix(1200, 900, "logspace", seed=2)
    f, Q, B = api.randQB_pb_new(A2, 40, 5, 0.0, 2, 1, seed=8)       # includes the even-step re-orthogonalisation
    fr, Qr, Br = oracle.randQB_pb_new(A2, 40, 5, 0.0, 2, 1, seed=8)
    assert f == fr == 200
    assert np.linalg.norm(Q @ B - Qr @ Br) <= 1e-12 * np.linalg.norm(A2)
    assert np.abs(Q.T @ Q - np.eye(200)).max() < 1e-12
    # kstep clamp (RRA:1589-1592): kstep > min(m,n)/2 is replaced by min(m,n)/10
    f, Q, B = api.randQB_pb_new(golden["A"], 30, 1, 0.0, 1, 1, seed=777)
    fr, Qr, Br = oracle.randQB_pb_new(golden["A"], 30, 1, 0.0, 1, 1, seed=777)
    assert f == fr and Q.shape == Qr.shape


def test_blockrand_svd_vs_golden_including_quirk_q1(api, golden):
    A = golden["A"]
    f, U, S, V = api.svd_blockrand(A, 8, 4, 0.0, 1, 4, 2, 1, seed=777)
    assert f == int(golden["blkA_rank_frank"]) and rel_sigma_err(S, golden["blkA_rank_S"]) < 1e-10
    f, U, S, V = api.svd_blockrand(A, 0, 4, 1.0, 1, 4, 2, 1, seed=777)   # k=0 never reaches tolerance mode (Q1)
    assert f == int(golden["blkA_tol_frank"]) and rel_sigma_err(S, golden["blkA_tol_S"]) < 1e-10


def test_blockrand_svd_vs_reference(api, oracle):
    A, _ = O.make_matrix(1500, 1100, "gap", seed=4, k=90, tail=1e-7)
    f, U, S, V = api.svd_blockrand(A, 90, 30, 0.0, 1, 30, 2, 1, seed=12)
    fr, Ur, Sr, Vr = oracle.svd_blockrand(A, 90, 30, 0.0, 1, 30, 2, 1, seed=12)
    assert f == fr and rel_sigma_err(S, Sr) < 1e-10 and subspace_sin(U, Ur) < 1e-6


# ---- ID / CUR ------------------------------------------------------------------------------------------------------
def test_blockrand_svd_tolerance_mode_when_integer_division_gives_no_blocks(api, oracle):
    """RRA:251-266: nstep = (k+p)/kstep by C integer division; with k = -1 and p <= kstep it is 0, and randQB_pb_new then runs
    TOLERANCE-driven — the one way this entry point reaches tolerance mode.  Same frank, sigma, subspaces."""
    if not hasattr(oracle, "lib"):
        pytest.skip("needs the compiled reference")
    A, _ = O.make_matrix(900, 700, "logspace", seed=9)
    tol = 0.3 * float(np.linalg.norm(A))
    f, U, S, V = api.svd_blockrand(A, -1, 20, tol, 1, 20, 2, 1, seed=12)
    fr, Ur, Sr, Vr = oracle.svd_blockrand(A, -1, 20, tol, 1, 20, 2, 1, seed=12)
    assert f == fr and 0 < f < 700 and S.shape == Sr.shape
    assert rel_sigma_err(S, Sr) < 1e-10 and subspace_sin(U, Ur) < 1e-6 and subspace_sin(V, Vr) < 1e-6


def test_id_cur_vs_golden(api, golden):
    A = golden["A"]
    I, T = api.id_rand(A, 8, 4, 2, 1, seed=777)
    assert np.array_equal(I, golden["idA_I"])                 # bit-exact index set
    assert np.allclose(T, golden["idA_T"], rtol=0, atol=1e-11)
    Ic, Ir, T, S = api.id_two_sided_rand(A, 8, 4, 2, 1, seed=777)
    assert np.array_equal(Ic, golden["id2A_Icol"]) and np.array_equal(Ir, golden["id2A_Irow"])
    assert np.allclose(T, golden["id2A_T"], atol=1e-11) and np.allclose(S, golden["id2A_S"], atol=1e-11)
    Cm, U, R = api.cur_rand(A, 8, 4, 2, 1, seed=777)
    assert np.array_equal(Cm, golden["curA_C"]) and np.array_equal(R, golden["curA_R"])
    assert np.allclose(U, golden["curA_U"], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("m,n,k,p,q,s,spec", [
    (2000, 1500, 100, 20, 2, 1, "logspace"),     # BASELINE config 0 shape, config 3's functions
    (800, 1200, 40, 10, 1, 2, "exp"),            # driver_multi_core_mkl3.c's q=1, s=2
    (1500, 700, 150, 20, 2, 1, "logspace"),      # min(m,n)-k > 128: exercises the dlaqps/dlaqp2 formula switch
])
def test_id_two_sided_and_cur_vs_reference(api, oracle, m, n, k, p, q, s, spec):
    A, _ = O.make_matrix(m, n, spec, seed=n)
    Ic, Ir, T, S = api.id_two_sided_rand(A, k, p, q, s, seed=21)
    Icr, Irr, Tr, Sr = oracle.id_two_sided_rand(A, k, p, q, s, seed=21)
    assert np.array_equal(Ic, Icr) and np.array_equal(Ir, Irr)   # full-length permutations, bit-exact
    assert np.abs(T - Tr).max() < 1e-10 and np.abs(S - Sr).max() < 1e-10
    assert sorted(Ic.astype(int)) == list(range(n)) and sorted(Ir.astype(int)) == list(range(m))
    Cm, U, R = api.cur_rand(A, k, p, q, s, seed=21)
    Cr, Ur, Rr = oracle.cur_rand(A, k, p, q, s, seed=21)
    assert np.array_equal(Cm, Cr) and np.array_equal(R, Rr)
    e, er = O.get_percent_error_between_two_mats(A, Cm @ U @ R), O.get_percent_error_between_two_mats(A, Cr @ Ur @ Rr)
    assert abs(e - er) <= 0.01 * er


def test_blockrand_id_cur_vs_reference(api, oracle):
    """id_blockrand / id_two_sided_blockrand / cur_blockrand (RRA:1969-2027, 2086-2111, 2262-2332),
    driver_multi_core_mkl3.c's parameters scaled down (p = kstep)."""
    A, _ = O.make_matrix(1200, 900, "logspace", seed=5)
    k, kstep, q, s = 80, 20, 1, 2
    ref = oracle if hasattr(oracle, "id_blockrand") else None
    f, I, T = api.id_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3)
    fr, Ir_, Tr = (ref.id_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3) if ref else O.id_blockrand_decomp_fixed_rank_or_prec(A, k, kstep, 0.0, kstep, q, s, 3))
    assert f == fr == k and np.array_equal(I, Ir_) and np.abs(T - Tr).max() < 1e-10
    f, Ic, Irow, T, S = api.id_two_sided_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3)
    fr, Icr, Irr, Tr, Sr = (ref.id_two_sided_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3) if ref else
                            O.id_two_sided_blockrand_decomp_fixed_rank_or_prec(A, k, kstep, 0.0, kstep, q, s, 3))
    assert f == fr and np.array_equal(Ic, Icr) and np.array_equal(Irow, Irr) and np.abs(S - Sr).max() < 1e-10
    f, Cm, U, R = api.cur_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3)
    fr, Cr, Ur, Rr = (ref.cur_blockrand(A, k, kstep, 0.0, kstep, q, s, seed=3) if ref else
                      O.cur_blockrand_decomp_fixed_rank_or_prec(A, k, kstep, 0.0, kstep, q, s, 3))
    assert f == fr and np.array_equal(Cm, Cr) and np.array_equal(R, Rr)
    e, er = O.get_percent_error_between_two_mats(A, Cm @ U @ R), O.get_percent_error_between_two_mats(A, Cr @ Ur @ Rr)
    assert abs(e - er) <= 0.01 * er
    # tolerance mode (k = 0): the QB loop stops on the device, frank = round(f/(f+p) f)  (this entry point DOES honour TOL)
    tol = 0.35 * np.linalg.norm(A)
    f, I, T = api.id_blockrand(A, 0, kstep, tol, kstep, q, s, seed=3)
    fr, Ir_, Tr = (ref.id_blockrand(A, 0, kstep, tol, kstep, q, s, seed=3) if ref else O.id_blockrand_decomp_fixed_rank_or_prec(A, 0, kstep, tol, kstep, q, s, 3))
    assert f == fr and 0 < f < 900 and np.array_equal(I, Ir_) and np.abs(T - Tr).max() < 1e-10


def test_svd_and_id_from_qb(api):
    """low_rank_svd_rand_decomp_fromQB / id_rand_decomp_fromQB (oneapi_code/…one_api.c:244-304, 421-444) in FP64,
    checked against the numpy restatement (the reference ships these in float32 only) and against the direct routines."""
    A, sig = O.make_matrix(900, 700, "gap", seed=8, k=40, tail=1e-7)
    f, Q, B = api.randQB_pb_new(A, 20, 3, 0.0, 2, 1, seed=5)
    U, S, V = api.svd_from_qb(Q, B)
    Ur, Sr, Vr = O.low_rank_svd_rand_decomp_fromQB(Q, B)
    assert rel_sigma_err(np.diag(S)[:40], np.diag(Sr)[:40]) < 1e-7      # sigma^2 route: eps*(s1/sk)^2
    assert rel_sigma_err(np.diag(S)[:40], sig[:40]) < 1e-6
    assert np.linalg.norm(A - U[:, :40] @ S[:40, :40] @ V[:, :40].T) / np.linalg.norm(A) < 1e-6
    I, T = api.id_from_qb(Q, B)
    Ir, Tr = O.id_rand_decomp_fromQB(Q, B)
    assert np.array_equal(I, Ir) and np.abs(T - Tr).max() < 1e-9


def test_evaluation_helpers_print_reference_metric(api, capfd):
    A, _ = O.make_matrix(400, 300, "exp", seed=1)
    U, S, V = api.svd_rand(A, 20, 5, 1, 2, 1, seed=1)
    M, Um, Sm, Vm = api.to_mat(A), api.to_mat(U), api.to_mat(S), api.to_mat(V)
    api.lib.use_low_rank_svd_for_approximation(M, Um, Sm, Vm)
    pe = api.lib.rsvd_b200_api_last_percent_error()
    assert pe == pytest.approx(100 * recon_err(A, U, S, V), rel=1e-9)
    assert "percent_error" in capfd.readouterr().out
    I, T = api.id_rand(A, 20, 5, 1, 1, seed=1)
    api.lib.use_id_decomp_for_approximation(M, api.to_mat(T), api.to_vec(I), 20)
    idx = I.astype(int)
    P = np.zeros_like(A)
    P[:, idx] = np.hstack([A[:, idx[:20]], A[:, idx[:20]] @ T])
    assert api.lib.rsvd_b200_api_last_percent_error() == pytest.approx(O.get_percent_error_between_two_mats(A, P), rel=1e-9)


def test_relinked_reference_driver_runs_end_to_end(tmp_path, oracle):
    """The reference's own driver_multi_core_mkl1.c, compiled UNMODIFIED against include/ and linked with the B200 libraries
    (oracle/build_ref.sh), loads ../../matrix_data/A_mat_1kx2k.bin and prints the percent error of a k=200,p=10,q=2 SVD."""
    import os
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "relink", "driver_multi_core_mkl1")
    if not os.path.exists(exe):
        pytest.skip("relinked driver not built")
    A, _ = O.make_matrix(1000, 2000, "logspace", seed=7)
    (tmp_path / "matrix_data").mkdir()
    O.write_matrix_binary(A, str(tmp_path / "matrix_data" / "A_mat_1kx2k.bin"), 32)
    cwd = tmp_path / "x" / "y"
    cwd.mkdir(parents=True)
    out = subprocess.run([os.path.abspath(exe)], cwd=str(cwd), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    m = re.search(r"percent_error between M and U S V\^T = ([0-9.eE+-]+)", out.stdout)
    assert m, out.stdout
    Ur, Sr, Vr = oracle.svd_rand(A, 200, 10, 1, 2, 1, seed=777)
    assert float(m.group(1)) == pytest.approx(100 * recon_err(A, Ur, Sr, Vr), rel=1e-4)
    assert "normM = %f" % np.linalg.norm(A) in out.stdout


def test_pivotedQR_mkl_explicit_q_from_reflectors(api):
    """pivotedQR_mkl (RRA:924-976, dgeqp3 + dorgqr): I bit-exact and |R| equal to LAPACK's, Q orthonormal and Q R = M(:, I) —
    also for a rank-deficient M and one with cond 1e12, where a Q rebuilt as M(:,I) R11^-1 would be garbage"""
    import ctypes as C
    from scipy.linalg import lapack
    rng = np.random.default_rng(8)
    cases = {"full rank": rng.standard_normal((200, 150)) * np.logspace(0, -3, 150),
             "rank-deficient": rng.standard_normal((180, 7)) @ rng.standard_normal((7, 90)),
             "cond 1e12": (np.linalg.qr(rng.standard_normal((300, 60)))[0] * np.logspace(0, -12, 60)) @ np.linalg.qr(rng.standard_normal((60, 60)))[0],
             "wide": rng.standard_normal((50, 400))}
    for name, A in cases.items():
        m, n = A.shape
        k = min(m, n)
        M = api.to_mat(A)
        Q, R, I = api.PM(), api.PM(), api.PV()
        api.lib.pivotedQR_mkl(M, C.byref(Q), C.byref(R), C.byref(I))
        api.check()
        api.lib.matrix_delete(M)
        Q, R, I = api.from_mat(Q), api.from_mat(R), api.from_vec(I).astype(int)
        assert Q.shape == (m, k) and R.shape == (k, n if m <= n else k)
        assert np.abs(Q.T @ Q - np.eye(k)).max() < 1e-13, name
        assert np.linalg.norm(Q @ R - A[:, I][:, :R.shape[1]]) <= 1e-13 * np.linalg.norm(A), name
        qr, jpvt, _, _, _ = lapack.dgeqp3(np.asfortranarray(A))
        if name != "rank-deficient":
            assert np.array_equal(I, jpvt - 1), name
            assert np.abs(np.abs(R) - np.abs(np.triu(qr[:k, :R.shape[1]]))).max() < 1e-11 * np.abs(qr).max(), name
