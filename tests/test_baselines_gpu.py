"""GPU parity tests for the deterministic baselines and legacy entry points, called through the
C-ABI exactly as the reference's drivers 2 and 4 do, against the unmodified reference C code (oracle/_ref) on the same
inputs and the same Omega.

Tolerances: pivot/index vectors and ranks bit-exact; factors with a fixed sign convention (the reference's own pivoted QR,
randQB_p, ID/CUR factors) elementwise to 1e-9 relative; SVD factors by singular values (1e-10, 1e-7 for the eig(BB^T)
variants that square the spectrum) and subspace angle, never elementwise (sign conventions are library-dependent)."""
import ctypes as C

import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native
from oracle import ref_lib
from helpers import subspace_sin, rel_sigma_err, recon_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    lib = native.dev()
    assert lib.rsvd_b200_init(0) == 0, lib.rsvd_b200_last_error().decode()
    return pkg.Api(32)


@pytest.fixture(scope="module")
def ref():
    if not ref_lib.available(32):
        pytest.skip("compiled reference (oracle/_ref) not present")
    return ref_lib.RefLib(32)


def decaying(m, n, r, lo=-6, seed=0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    return (U * np.logspace(0, lo, r)) @ V.T


def relerr(X, Y):
    return float(np.linalg.norm(X - Y) / max(np.linalg.norm(Y), 1e-300))


# ---- the reference's own partial pivoted QR (RRA:1012-1334) --------------------------------------------------------
@pytest.mark.parametrize("m,n,r,k,TOL", [
    (60, 80, 25, 10, None),        # pivoted_QR_of_specified_rank (driver 2)
    (300, 200, 120, 50, None),
    (2000, 300, 150, 40, None),    # more than 1280 rows: tall-column kernels
    (60, 80, 25, 10, 0.0),         # pivoted_QR_of_specified_rank_or_prec, rank mode (driver 4)
    (300, 500, 100, 0, 1e-4),      # tolerance mode: R22norm refreshed every 5th step
    (200, 150, 60, 0, 1e-6),
])
def test_partial_pivoted_qr_vs_reference(api, ref, m, n, r, k, TOL):
    A = decaying(m, n, r, seed=m + n)
    f, Q, R, I = api.pqr(A, k, TOL)
    api.check()
    f0, Q0, R0, I0 = ref.pqr(A, k, TOL)
    assert f == f0 and (k <= 0 or f == k)
    assert np.array_equal(I, I0)
    assert Q.shape == (m, f) and R.shape == (f, n)
    assert relerr(R, R0) < 1e-9 and relerr(Q, Q0) < 1e-9          # same reflector convention: R(i,i) = +||x||
    assert np.all(np.diag(R) >= 0)
    assert np.abs(Q.T @ Q - np.eye(f)).max() < 1e-12
    # A(:, I) - Qk Rk is exactly the trailing block R22: compare its size with the reference's
    e, e0 = relerr(Q @ R, A[:, I.astype(int)]), relerr(Q0 @ R0, A[:, I0.astype(int)])
    assert abs(e - e0) <= 1e-9 + 1e-6 * e0


def test_partial_pivoted_qr_stops_on_exhausted_rank(api, ref):
    A = decaying(120, 90, 7, lo=-1, seed=3)                         # exactly rank 7
    f, Q, R, I = api.pqr(A, 30, 0.0)                                # |norm^2| < 1e-10 break (RRA:1269-1272)
    f0, _, _, I0 = ref.pqr(A, 30, 0.0)
    assert f == f0 == 7 and np.array_equal(I, I0)
    assert relerr(Q @ R, A[:, I.astype(int)]) < 1e-12


def test_use_pivoted_qr_for_approximation(api):
    A = decaying(150, 100, 40, lo=-8, seed=5)
    L = api.lib
    L.use_pivoted_QR_decomp_for_approximation.argtypes = [api.PM, api.PM, api.PM, api.PV]
    M = api.to_mat(A)
    Q, R, I_ = api.PM(), api.PM(), api.PV()
    fr = api.I(0)
    L.pivoted_QR_of_specified_rank_or_prec(M, 40, 0.0, C.byref(fr), C.byref(Q), C.byref(R), C.byref(I_))
    L.use_pivoted_QR_decomp_for_approximation(M, Q, R, I_)          # unlike the reference (RRA:2378) it does not free Q and R
    # the scan stops once the largest remaining squared column norm is below 1e-10 (RRA:1269), i.e. at ~1e-5 relative error
    f, Qn, Rn, In = api.pqr(A, 40, 0.0)
    P = np.empty_like(A)
    P[:, In.astype(int)] = Qn @ Rn
    expect = 100.0 * np.linalg.norm(A - P) / np.linalg.norm(A)
    assert f == fr.value < 40
    assert abs(L.rsvd_b200_api_last_percent_error() - expect) < 1e-9 and expect < 1e-2
    L.matrix_delete(M)
    L.matrix_delete(Q)
    L.matrix_delete(R)
    L.vector_delete(I_)


# ---- deterministic SVD / ID / CUR baselines ----------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(300, 120), (120, 300), (200, 200), (1500, 64)])
def test_full_svd_baseline(api, ref, m, n):
    A = decaying(m, n, min(m, n), lo=-5, seed=m)
    k = min(m, n) // 3
    f, U, S, V = api.svd_decomp(A, k, 0.0)
    api.check()
    s = np.linalg.svd(A, compute_uv=False)
    assert f == k and U.shape == (m, k) and S.shape == (k, k) and V.shape == (n, k)
    assert np.max(np.abs(np.diag(S) - s[:k])) / s[0] < 1e-13
    assert abs(recon_err(A, U, S, V) - np.linalg.norm(s[k:]) / np.linalg.norm(s)) < 1e-12
    f0, U0, S0, V0 = ref.svd_decomp(A, k, 0.0)
    assert rel_sigma_err(S, S0) < 1e-10 and subspace_sin(U, U0) < 1e-6 and subspace_sin(V, V0) < 1e-6
    # tolerance mode keeps the reference's integer abs() (RRA:49)
    B = 7.5 * A
    assert api.svd_decomp(B, 0, 3.0)[0] == ref.svd_decomp(B, 0, 3.0)[0]
    assert api.svd_decomp(B, 0, 0.5)[0] == ref.svd_decomp(B, 0, 0.5)[0]


@pytest.mark.parametrize("m,n", [(90, 200), (300, 120)])
def test_gesvd_nonsquare(api, m, n):
    A = decaying(m, n, min(m, n), lo=-4, seed=n)
    U, S, Vt = api.gesvd(A)
    api.check()
    assert relerr(U @ S @ Vt, A) < 1e-12
    assert np.max(np.abs(np.diag(S) - np.linalg.svd(A, compute_uv=False))) < 1e-13


@pytest.mark.parametrize("m,n,r,k,TOL", [(200, 300, 80, 30, 0.0), (400, 250, 100, 45, 0.0), (200, 300, 80, 0, 1e-5)])
def test_id_baselines_vs_reference(api, ref, m, n, r, k, TOL):
    A = decaying(m, n, r, seed=7 * m + n)
    f, I, T = api.id_decomp(A, k, TOL)
    api.check()
    f0, I0, T0 = ref.id_decomp(A, k, TOL)
    assert f == f0 and np.array_equal(I, I0) and T.shape == T0.shape == (f, n - f)
    assert relerr(T, T0) < 1e-8
    f, Ic, Ir, T, S = api.id_two_sided_decomp(A, k, TOL)
    api.check()
    f0, Ic0, Ir0, T0, S0 = ref.id_two_sided_decomp(A, k, TOL)
    assert f == f0 and np.array_equal(Ic, Ic0) and np.array_equal(Ir, Ir0)
    assert relerr(T, T0) < 1e-8 and relerr(S, S0) < 1e-8
    f, Cm, U, R = api.cur_decomp(A, k, TOL)
    api.check()
    f0, Cm0, U0, R0 = ref.cur_decomp(A, k, TOL)
    assert f == f0 and np.array_equal(Cm, Cm0) and np.array_equal(R, R0)      # C, R are gathers: bit-exact
    assert relerr(Cm @ U @ R, Cm0 @ U0 @ R0) < 1e-7
    assert abs(relerr(Cm @ U @ R, A) - relerr(Cm0 @ U0 @ R0, A)) < 1e-6


# ---- legacy randQB / randomized SVD entry points -----------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,p", [(300, 200, 12, 0), (250, 400, 10, 1), (500, 300, 8, 2)])
def test_randqb_p_vs_reference(api, ref, m, n, k, p):
    A = decaying(m, n, 60, lo=-4, seed=k)
    Q, B = api.randQB_p(A, k, p, seed=11)
    api.check()
    Q0, B0 = ref.randQB_p(A, k, p, seed=11)
    assert Q.shape == (m, k) and B.shape == (k, n)
    assert relerr(Q, Q0) < 1e-8 and relerr(B, B0) < 1e-8              # q_j = y_j/||y_j||: no sign freedom
    assert np.abs(Q.T @ Q - np.eye(k)).max() < 1e-10


@pytest.mark.parametrize("m,n,kstep,nstep,p,s", [(400, 300, 8, 4, 0, 1), (300, 500, 10, 3, 1, 1), (600, 400, 16, 3, 2, 2)])
def test_randqb_pb_vs_reference(api, ref, m, n, kstep, nstep, p, s):
    A = decaying(m, n, 120, lo=-5, seed=kstep)
    Q, B = api.randQB_pb(A, kstep, nstep, p, s, seed=5)
    api.check()
    Q0, B0 = ref.randQB_pb(A, kstep, nstep, p, s, seed=5)
    l = kstep * nstep
    assert Q.shape == (m, l) and B.shape == (l, n)
    assert np.abs(Q.T @ Q - np.eye(l)).max() < 1e-12
    assert subspace_sin(Q, np.linalg.qr(Q0)[0]) < 1e-6
    assert relerr(Q @ B, Q0 @ B0) < 1e-9                               # block signs differ (Householder vs Cholesky QR), QB does not


@pytest.mark.parametrize("which", ["svd1", "svd2", "svd3", "svd4"])
def test_legacy_randomized_svd_vs_reference(api, ref, which):
    m, n, k = 500, 400, 24
    s = np.concatenate([np.logspace(0, -3, k), 1e-9 * np.logspace(0, -2, 100 - k)])   # gap after k (risk R1)
    rng = np.random.default_rng(21)
    U0, _ = np.linalg.qr(rng.standard_normal((m, 100)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, 100)))
    A = (U0 * s) @ V0.T
    args = {"svd1": (k,), "svd2": (k,), "svd3": (k, 3, 1), "svd4": (8, 3, 1)}[which]
    U, S, V = getattr(api, which)(A, *args, seed=9)
    api.check()
    Ur, Sr, Vr = getattr(ref, which)(A, *args, seed=9)
    assert U.shape == Ur.shape and S.shape == Sr.shape and V.shape == Vr.shape
    d, dr = np.diag(S), np.diag(Sr)
    if which in ("svd1", "svd4"):                                       # eig(B B^T): ascending, spectrum squared
        assert np.all(np.diff(d) >= 0)
        assert np.max(np.abs(d - dr)) / dr.max() < 1e-7
    else:
        assert np.all(np.diff(d) <= 0)
        assert rel_sigma_err(S, Sr) < 1e-9
    assert abs(recon_err(A, U, S, V) - recon_err(A, Ur, Sr, Vr)) < 1e-6
    assert subspace_sin(np.linalg.qr(U)[0], np.linalg.qr(Ur)[0]) < 1e-5


def test_estimate_rank_and_autorank1_vs_reference(api, ref):
    A = decaying(300, 200, 30, lo=-3, seed=2)                           # exactly rank 30
    r, Q = api.estimate_rank1(A, 0.5, 1e-8, seed=3)
    api.check()
    r0, Q0 = ref.estimate_rank1(A, 0.5, 1e-8, seed=3)
    assert r == r0 and Q.shape == (300, r)
    assert np.abs(Q.T @ Q - np.eye(r)).max() < 1e-10
    assert subspace_sin(Q[:, :30], np.linalg.qr(Q0[:, :30])[0]) < 1e-5
    U, S, V = api.svd2_autorank1(A, 0.5, 1e-8, seed=3)
    api.check()
    Ur, Sr, Vr = ref.svd2_autorank1(A, 0.5, 1e-8, seed=3)
    assert S.shape == Sr.shape
    assert np.max(np.abs(np.diag(S)[:30] - np.diag(Sr)[:30])) / Sr[0, 0] < 1e-9
    assert recon_err(A, U, S, V) < 1e-9


def test_autorank2(api, ref):
    A = decaying(400, 300, 40, lo=-2, seed=4)                           # exactly rank 40
    # one block suffices: identical to the reference (whose later blocks would repeat the first)
    r, Y, Q = api.estimate_rank2(A, 50, 1e-6, seed=8)
    api.check()
    r0, Y0, Q0 = ref.estimate_rank2(A, 50, 1e-6, seed=8)
    assert r == r0 == 50 and relerr(Y, Y0) < 1e-10
    U, S, V = api.svd2_autorank2(A, 50, 1e-6, seed=8)
    Ur, Sr, Vr = ref.svd2_autorank2(A, 50, 1e-6, seed=8)
    assert np.max(np.abs(np.diag(S)[:40] - np.diag(Sr)[:40])) / Sr[0, 0] < 1e-9
    U, S, V = api.svd3_autorank2(A, 50, 1e-6, 2, 1, seed=8)
    Ur, Sr, Vr = ref.svd3_autorank2(A, 50, 1e-6, 2, 1, seed=8)
    assert np.max(np.abs(np.diag(S)[:40] - np.diag(Sr)[:40])) / Sr[0, 0] < 1e-9
    # several blocks: fresh Omega per block, stop as soon as ||QQ^T A - A||/||QQ^T A|| <= TOL
    r, Y, Q = api.estimate_rank2(A, 16, 1e-6, seed=8)
    api.check()
    assert r == 48 and np.abs(Q.T @ Q - np.eye(r)).max() < 1e-10
    assert relerr(Q @ (Q.T @ A), A) <= 1e-6
    U, S, V = api.svd2_autorank2(A, 16, 1e-6, seed=8)
    assert S.shape == (48, 48) and recon_err(A, U, S, V) < 1e-6


# ---- driver-level parity: the reference's drivers 2 and 4, unmodified, relinked against the B200 libraries ------------------
def _driver_numbers(lines):
    import re
    ranks = [int(re.search(r"output rank is:\s*(-?\d+)", l).group(1)) for l in lines if "output rank is" in l]
    errs = [float(l.rsplit("=", 1)[1]) for l in lines if "percent" in l and "=" in l]
    return ranks, errs


@pytest.mark.parametrize("which", [2, 4])
def test_relinked_drivers_2_and_4_match_the_reference_output(tmp_path, which):
    """tests/golden/driver{2,4}_ref.txt hold what the reference's own binaries print (tests/golden/make_driver_golden.py);
    the same sources linked with librsvd_b200_api32.so must print the same ranks and percent errors (Omega seed 777 on both
    sides).  Driver 4's reference run aborts inside cur_decomp_fixed_rank_or_prec, so its list is a prefix of ours."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(here, "..", "oracle", "_ref", "relink", "driver_multi_core_mkl%d" % which)
    if not os.path.exists(exe):
        pytest.skip("relinked driver not built")
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_driver_golden as G
    rc, lines = G.run_driver(os.path.abspath(exe), which, str(tmp_path))
    assert rc == 0
    ranks, errs = _driver_numbers(lines)
    gold = [l.strip() for l in open(os.path.join(here, "golden", "driver%d_ref.txt" % which)) if not l.startswith("#")]
    granks, gerrs = _driver_numbers(gold)
    assert len(granks) <= len(ranks) and len(gerrs) <= len(errs) and len(gerrs) >= 6
    assert ranks[:len(granks)] == granks
    for e, g in zip(errs, gerrs):
        assert e == pytest.approx(g, rel=2e-5, abs=2e-6)


def test_baselines_through_the_int64_abi(api):
    """multi_core_mkl_code_64bit: the same entry points with int64_t indices (rank_revealing_algorithms_intel_mkl.h:3-91 there)."""
    if not ref_lib.available(64):
        pytest.skip("compiled 64-bit reference not present")
    api64, ref64 = pkg.Api(64), ref_lib.RefLib(64)
    A = decaying(180, 240, 70, seed=64)
    f, Q, R, I = api64.pqr(A, 25, 0.0)
    f0, Q0, R0, I0 = ref64.pqr(A, 25, 0.0)
    assert f == f0 == 25 and np.array_equal(I, I0) and relerr(R, R0) < 1e-9 and relerr(Q, Q0) < 1e-9
    f, Ic, Ir, T, S = api64.id_two_sided_decomp(A, 0, 1e-5)
    f0, Ic0, Ir0, T0, S0 = ref64.id_two_sided_decomp(A, 0, 1e-5)
    assert f == f0 and np.array_equal(Ic, Ic0) and np.array_equal(Ir, Ir0) and relerr(T, T0) < 1e-8 and relerr(S, S0) < 1e-8
    Qb, Bb = api64.randQB_pb(A, 8, 3, 1, 1, seed=5)
    Q0, B0 = ref64.randQB_pb(A, 8, 3, 1, 1, seed=5)
    assert relerr(Qb @ Bb, Q0 @ B0) < 1e-9
    fr, U, S_, V = api64.svd_decomp(A, 20, 0.0)
    assert fr == 20 and np.max(np.abs(np.diag(S_) - np.linalg.svd(A, compute_uv=False)[:20])) < 1e-13


def test_baselines_vs_committed_golden_vectors(api):
    """Same checks against tests/golden/golden_baselines.npz (outputs of the compiled reference, committed): runs even where
    oracle/_ref is absent."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_baselines.npz"))
    A = g["A"]
    f, Q, R, I = api.pqr(A, 10)
    assert f == int(g["pqr_k10_frank"]) and np.array_equal(I, g["pqr_k10_I"])
    assert relerr(R, g["pqr_k10_R"]) < 1e-10 and relerr(Q, g["pqr_k10_Q"]) < 1e-10
    f, Q, R, I = api.pqr(A, 0, 0.05)
    assert f == int(g["pqr_tol_frank"]) and np.array_equal(I, g["pqr_tol_I"]) and relerr(R, g["pqr_tol_R"]) < 1e-10
    f, I, T = api.id_decomp(A, 12, 0.0)
    assert f == int(g["id_k12_frank"]) and np.array_equal(I, g["id_k12_I"]) and relerr(T, g["id_k12_T"]) < 1e-9
    f, Ic, Ir, T, S = api.id_two_sided_decomp(A, 0, 0.05)
    assert f == int(g["id2_tol_frank"]) and np.array_equal(Ic, g["id2_tol_Icol"]) and np.array_equal(Ir, g["id2_tol_Irow"])
    assert relerr(T, g["id2_tol_T"]) < 1e-9 and relerr(S, g["id2_tol_S"]) < 1e-9
    f, Cm, U, R = api.cur_decomp(A, 9, 0.0)
    assert np.array_equal(Cm, g["cur_k9_C"]) and np.array_equal(R, g["cur_k9_R"])
    assert relerr(Cm @ U @ R, g["cur_k9_C"] @ g["cur_k9_U"] @ g["cur_k9_R"]) < 1e-8
    assert rel_sigma_err(api.svd_decomp(A, 7, 0.0)[2], g["svd_k7_S"]) < 1e-12
    assert api.svd_decomp(A, 0, 2.0)[0] == int(g["svd_tol_frank"])
    Q, B = api.randQB_p(A, 6, 1, seed=777)
    assert relerr(Q, g["qbp_Q"]) < 1e-9 and relerr(B, g["qbp_B"]) < 1e-9
    Q, B = api.randQB_pb(A, 4, 3, 1, 1, seed=777)
    assert relerr(Q @ B, g["qbpb_QB"]) < 1e-10
    for name, args, tol in [("svd1", (8,), 1e-7), ("svd2", (8,), 1e-10), ("svd3", (8, 3, 1), 1e-10), ("svd4", (4, 2, 1), 1e-7)]:
        S = getattr(api, name)(A, *args, seed=777)[1]
        assert np.max(np.abs(np.diag(S) - np.diag(g[name + "_S"]))) / np.diag(g[name + "_S"]).max() < tol, name
    assert api.estimate_rank1(A, 0.5, 1e-3, seed=777)[0] == int(g["rank1"])
    api.check()


# ---- edge cases: the smallest shapes every new entry point accepts -----------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (1, 5, 1), (5, 1, 1), (6, 9, 1), (9, 6, 6), (7, 7, 7)])
def test_partial_pivoted_qr_edge_shapes(api, ref, m, n, k):
    A = np.random.default_rng(10 * m + n).standard_normal((m, n)) + 0.1
    f, Q, R, I = api.pqr(A, k, 0.0)
    api.check()
    f0, Q0, R0, I0 = ref.pqr(A, k, 0.0)
    assert f == f0 and np.array_equal(I, I0)
    assert relerr(R, R0) < 1e-10 and relerr(Q, Q0) < 1e-10
    assert sorted(I.astype(int).tolist()) == list(range(n))


@pytest.mark.parametrize("m,n", [(1, 1), (5, 1), (1, 5), (2, 2), (40, 3), (3, 40)])
def test_full_svd_edge_shapes(api, m, n):
    A = np.random.default_rng(m + 7 * n).standard_normal((m, n))
    r = min(m, n)
    f, U, S, V = api.svd_decomp(A, r, 0.0)
    api.check()
    assert f == r and relerr(U @ S @ V.T, A) < 1e-12
    assert np.allclose(np.diag(S), np.linalg.svd(A, compute_uv=False), rtol=1e-12, atol=1e-14)


def test_rank_deficient_inputs(api, ref):
    u, v = np.arange(1.0, 31.0), np.cos(np.arange(20.0))
    A = np.outer(u, v)                                                  # exactly rank 1
    f, U, S, V = api.svd_decomp(A, 3, 0.0)
    api.check()
    assert abs(S[0, 0] - np.linalg.norm(u) * np.linalg.norm(v)) < 1e-10 * S[0, 0] and S[1, 1] < 1e-10 * S[0, 0]
    f, Q, R, I = api.pqr(A, 5, 0.0)                                     # stops after one step: every remaining norm is ~0
    f0, _, _, I0 = ref.pqr(A, 5, 0.0)
    assert f == f0 == 1 and np.array_equal(I, I0)
    f, I, T = api.id_decomp(A, 5, 0.0)
    api.check()
    assert f == 1 and T.shape == (1, 19) and relerr(A[:, I[:1].astype(int)] @ T, A[:, I[1:].astype(int)]) < 1e-12


def test_legacy_entry_points_smallest_parameters(api, ref):
    A = decaying(40, 30, 12, lo=-3, seed=1)
    Q, B = api.randQB_p(A, 1, 0, seed=3)
    Q0, B0 = ref.randQB_p(A, 1, 0, seed=3)
    assert relerr(Q, Q0) < 1e-10 and relerr(B, B0) < 1e-10
    Q, B = api.randQB_p(A, 2, 1, seed=3)                                # the Gram-Schmidt loop i < j-1 is still empty
    Q0, B0 = ref.randQB_p(A, 2, 1, seed=3)
    assert relerr(Q, Q0) < 1e-9 and relerr(B, B0) < 1e-9
    Q, B = api.randQB_pb(A, 5, 1, 0, 1, seed=3)                         # one block, no power iteration
    Q0, B0 = ref.randQB_pb(A, 5, 1, 0, 1, seed=3)
    assert relerr(Q @ B, Q0 @ B0) < 1e-10
    U, S, V = api.svd2(A, 1, seed=3)
    assert S.shape == (1, 1) and abs(S[0, 0] - ref.svd2(A, 1, seed=3)[1][0, 0]) < 1e-10
    r, Q = api.estimate_rank1(A, 1.0 / 30, 1e-8, seed=3)                # maxdim = 1
    assert r == ref.estimate_rank1(A, 1.0 / 30, 1e-8, seed=3)[0] == 1 and Q.shape == (40, 1)
    api.check()


def test_invalid_parameters_are_reported_not_fatal(api):
    A = decaying(20, 15, 8, seed=2)
    for call in (lambda: api.pqr(A, 0), lambda: api.randQB_p(A, 0, 1), lambda: api.randQB_pb(A, 4, 0, 1, 1),
                 lambda: api.randQB_pb(A, 10, 4, 1, 1), lambda: api.svd3(A, 4, 2, 0), lambda: api.estimate_rank1(A, 0.0, 1e-3),
                 lambda: api.estimate_rank2(A, 0, 0.1)):
        call()                                                           # returns (void API), outputs allocated
        with pytest.raises(RuntimeError):
            api.check()
    f, Q, R, I = api.pqr(A, 99, 0.0)                                     # k > min(m,n): clamped — a warning, not an error:
    assert f <= 15 and np.abs(R).max() > 0 and np.abs(Q).max() > 0       # every output carries the (clamped) result
    api.check()
    api.lib.rsvd_b200_api_last_warning.restype = __import__("ctypes").c_char_p
    assert b"clamped" in api.lib.rsvd_b200_api_last_warning()
