"""CPU tests: pin the oracle.  (i) Philox known-answer vectors (Random123 kat_vectors);
(ii) numpy twin == golden vectors produced by the unmodified reference C code;
(iii) when oracle/_ref is present, numpy twin == live reference on fresh inputs."""
import numpy as np
import pytest

from oracle import ref_lib, rsvd_numpy as O
from helpers import subspace_sin, rel_sigma_err


def test_philox_known_answers():
    assert ref_lib.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ref_lib.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ref_lib.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normal_stream_golden_and_moments(golden):
    z = ref_lib.normal_stream(42, 1000, 64)
    assert np.array_equal(z, golden["normals_seed42_first1000_64"])
    assert np.array_equal(O.initialize_random_matrix(7, 5, 777), golden["omega_7x5_seed777"])
    z = ref_lib.normal_stream(1, 0, 2_000_000)
    assert np.all(z == z.astype(np.float32))  # float32-valued like the reference's `float *r`
    assert abs(z.mean()) < 3e-3 and abs(z.var() - 1) < 4e-3 and abs((z ** 4).mean() - 3) < 3e-2
    # random access == streaming
    assert np.array_equal(ref_lib.normal_stream(1, 12345, 7), z[12345:12352])


@pytest.mark.parametrize("name,mat", [("svdA_v1", "A"), ("svdA_v2", "A"), ("svdG_v1", "G")])
def test_twin_svd_vs_golden(golden, name, mat):
    k, p, vnum, q, s, seed = [int(x) for x in golden[name + "_params"]]
    U, S, V = O.low_rank_svd_rand_decomp_fixed_rank(golden[mat], k, p, vnum, q, s, seed)
    assert rel_sigma_err(S, golden[name + "_S"]) < 1e-10
    assert subspace_sin(U, golden[name + "_U"]) < 1e-6
    assert subspace_sin(V, golden[name + "_V"]) < 1e-6
    if vnum == 2:  # ascending order (RRA:220-223)
        assert np.all(np.diff(np.diag(S)) >= 0)
    else:
        assert np.all(np.diff(np.diag(S)) <= 0)


def test_golden_64bit_abi_matches_32bit(golden):
    assert np.allclose(golden["svdA_v1_S_64bit"], golden["svdA_v1_S"], rtol=1e-13, atol=0)


def test_twin_qb_vs_golden(golden):
    A = golden["A"]
    f, Q, B = O.randQB_pb_new(A, 4, 3, 0.0, 2, 1, 777)
    assert f == int(golden["qbA_rank_frank"]) == 12
    assert np.allclose(Q @ B, golden["qbA_rank_Q"] @ golden["qbA_rank_B"], atol=1e-12)
    f, Q, B = O.randQB_pb_new(A, 4, 0, 2.0, 1, 1, 777)
    assert f == int(golden["qbA_tol_frank"]) and Q.shape == golden["qbA_tol_Q"].shape
    assert np.allclose(Q @ B, golden["qbA_tol_Q"] @ golden["qbA_tol_B"], atol=1e-12)
    assert np.linalg.norm(A - Q @ B) < 2.0  # absolute Frobenius tolerance (RRA:1773-1775)


def test_twin_blockrand_vs_golden_including_quirk_q1(golden):
    A = golden["A"]
    f, U, S, V = O.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(A, 8, 4, 0.0, 1, 4, 2, 1, 777)
    assert f == int(golden["blkA_rank_frank"]) == 8
    assert rel_sigma_err(S, golden["blkA_rank_S"]) < 1e-10
    # k=0 "tolerance mode": reference runs ONE block and reports round(kstep/(kstep+p)*kstep)
    f, U, S, V = O.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(A, 0, 4, 1.0, 1, 4, 2, 1, 777)
    assert f == int(golden["blkA_tol_frank"]) == 2
    assert rel_sigma_err(S, golden["blkA_tol_S"]) < 1e-10


def test_twin_id_cur_vs_golden(golden):
    A = golden["A"]
    I, T = O.id_rand_decomp_fixed_rank(A, 8, 4, 2, 1, 777)
    assert np.array_equal(I, golden["idA_I"])            # bit-exact pivots
    assert np.allclose(T, golden["idA_T"], rtol=0, atol=1e-11)
    Ic, Ir, T, S = O.id_two_sided_rand_decomp_fixed_rank(A, 8, 4, 2, 1, 777)
    assert np.array_equal(Ic, golden["id2A_Icol"]) and np.array_equal(Ir, golden["id2A_Irow"])
    assert np.allclose(S, golden["id2A_S"], rtol=0, atol=1e-11)
    assert sorted(Ic.astype(int)) == list(range(A.shape[1]))   # full-length permutations
    assert sorted(Ir.astype(int)) == list(range(A.shape[0]))
    Cm, U, R = O.cur_rand_decomp_fixed_rank(A, 8, 4, 2, 1, 777)
    assert np.array_equal(Cm, golden["curA_C"]) and np.array_equal(R, golden["curA_R"])
    assert np.allclose(U, golden["curA_U"], rtol=1e-8, atol=1e-10)


def test_twin_vs_live_reference(ref32):
    """Fresh inputs (not in the golden file) through the compiled reference, both ABIs' twin."""
    A, _ = O.make_matrix(150, 220, "exp", seed=11)
    U, S, V = ref32.svd_rand(A, 15, 5, 1, 3, 1, seed=9)
    U2, S2, V2 = O.low_rank_svd_rand_decomp_fixed_rank(A, 15, 5, 1, 3, 1, seed=9)
    assert rel_sigma_err(S2, S) < 1e-10 and subspace_sin(U2, U) < 1e-7
    Ic, Ir, T, Sm = ref32.id_two_sided_rand(A, 15, 5, 1, 2, seed=9)
    Ic2, Ir2, T2, Sm2 = O.id_two_sided_rand_decomp_fixed_rank(A, 15, 5, 1, 2, seed=9)
    assert np.array_equal(Ic, Ic2) and np.array_equal(Ir, Ir2)
    assert np.allclose(T, T2, atol=1e-11) and np.allclose(Sm, Sm2, atol=1e-11)


def test_binary_io_formats_roundtrip_with_reference(ref32, tmp_path):
    """File format of MVF:77-133: int32 m,n then ROW-major doubles (64-bit ABI: int64 header)."""
    A = np.arange(12, dtype=np.float64).reshape(3, 4) + 0.5
    f = str(tmp_path / "a.bin")
    O.write_matrix_binary(A, f, 32)
    M = ref32.lib.matrix_load_from_binary_file(f.encode())
    assert np.array_equal(ref32.from_mat(M), A)
    g = str(tmp_path / "b.bin")
    M = ref32.to_mat(A)
    ref32.lib.matrix_write_to_binary_file(M, g.encode())
    ref32.lib.matrix_delete(M)
    assert np.array_equal(O.read_matrix_binary(g, 32), A)
    assert open(f, "rb").read() == open(g, "rb").read()


# ---- twins of the deterministic baselines / legacy entry points, pinned against the compiled reference --------------------
def _decaying(m, n, r, lo=-6, seed=0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    return (U * np.logspace(0, lo, r)) @ V.T


@pytest.mark.parametrize("m,n,r,k,TOL", [(60, 80, 25, 10, None), (90, 70, 40, 20, 0.0), (80, 120, 50, 0, 1e-4), (50, 40, 6, 20, 0.0)])
def test_twin_partial_pivoted_qr_matches_reference(ref32, m, n, r, k, TOL):
    A = _decaying(m, n, r, lo=-5 if r > 10 else -1, seed=m)
    f, Q, R, I = O.pivoted_QR_of_specified_rank_or_prec(A, k, TOL)
    f0, Q0, R0, I0 = ref32.pqr(A, k, TOL)
    assert f == f0 and np.array_equal(I, I0)
    assert np.linalg.norm(R - R0) <= 1e-11 * np.linalg.norm(R0) and np.linalg.norm(Q - Q0) <= 1e-11 * np.linalg.norm(Q0)


def test_twin_deterministic_id_and_svd_match_reference(ref32):
    A = _decaying(70, 90, 40, seed=5)
    for k, TOL in [(15, 0.0), (0, 1e-4)]:
        f, I, T = O.id_decomp_fixed_rank_or_prec(A, k, TOL)
        f0, I0, T0 = ref32.id_decomp(A, k, TOL)
        assert f == f0 and np.array_equal(I, I0) and np.linalg.norm(T - T0) <= 1e-9 * np.linalg.norm(T0)
        f, Ic, Ir, T, S = O.id_two_sided_decomp_fixed_rank_or_prec(A, k, TOL)
        f0, Ic0, Ir0, T0, S0 = ref32.id_two_sided_decomp(A, k, TOL)
        assert f == f0 and np.array_equal(Ic, Ic0) and np.array_equal(Ir, Ir0) and np.linalg.norm(S - S0) <= 1e-9 * np.linalg.norm(S0)
    for k, TOL in [(12, 0.0), (0, 3.0), (0, 0.5)]:
        B = 7.5 * A
        f, U, S, V = O.low_rank_svd_decomp_fixed_rank_or_prec(B, k, TOL)
        f0, U0, S0, V0 = ref32.svd_decomp(B, k, TOL)
        assert f == f0 and np.allclose(np.diag(S), np.diag(S0), rtol=1e-12)


def test_twin_legacy_randqb_and_rank_estimate_match_reference(ref32):
    A = _decaying(120, 90, 40, lo=-4, seed=9)
    Q, B = O.randQB_p(A, 8, 1, seed=11)
    Q0, B0 = ref32.randQB_p(A, 8, 1, seed=11)
    assert np.linalg.norm(Q - Q0) < 1e-9 and np.linalg.norm(B - B0) < 1e-9
    Q, B = O.randQB_pb(A, 6, 3, 1, 1, seed=5)
    Q0, B0 = ref32.randQB_pb(A, 6, 3, 1, 1, seed=5)
    assert np.linalg.norm(Q @ B - Q0 @ B0) <= 1e-10 * np.linalg.norm(A)
    A = _decaying(100, 80, 12, lo=-2, seed=2)
    r, Qe = O.estimate_rank_and_buildQ(A, 0.5, 1e-8, seed=3)
    r0, Qe0 = ref32.estimate_rank1(A, 0.5, 1e-8, seed=3)
    assert r == r0


def test_twins_match_golden_baselines():
    """tests/golden/golden_baselines.npz was produced by the compiled reference (tests/golden/make_golden.py); the numpy twins
    must reproduce it without oracle/_ref being present."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_baselines.npz"))
    A = g["A"]
    f, Q, R, I = O.pivoted_QR_of_specified_rank_or_prec(A, 10, None)
    assert f == int(g["pqr_k10_frank"]) and np.array_equal(I, g["pqr_k10_I"])
    assert np.allclose(R, g["pqr_k10_R"], rtol=0, atol=1e-12) and np.allclose(Q, g["pqr_k10_Q"], rtol=0, atol=1e-12)
    f, Q, R, I = O.pivoted_QR_of_specified_rank_or_prec(A, 0, 0.05)
    assert f == int(g["pqr_tol_frank"]) and np.array_equal(I, g["pqr_tol_I"])
    f, I, T = O.id_decomp_fixed_rank_or_prec(A, 12, 0.0)
    assert f == int(g["id_k12_frank"]) and np.array_equal(I, g["id_k12_I"]) and np.allclose(T, g["id_k12_T"], atol=1e-10)
    f, Ic, Ir, T, S = O.id_two_sided_decomp_fixed_rank_or_prec(A, 0, 0.05)
    assert f == int(g["id2_tol_frank"]) and np.array_equal(Ic, g["id2_tol_Icol"]) and np.array_equal(Ir, g["id2_tol_Irow"])
    assert np.allclose(np.diag(O.low_rank_svd_decomp_fixed_rank_or_prec(A, 7, 0.0)[2]), np.diag(g["svd_k7_S"]), rtol=1e-12)
    assert O.low_rank_svd_decomp_fixed_rank_or_prec(A, 0, 2.0)[0] == int(g["svd_tol_frank"])
    Q, B = O.randQB_p(A, 6, 1, seed=777)
    assert np.allclose(Q, g["qbp_Q"], atol=1e-10) and np.allclose(B, g["qbp_B"], atol=1e-10)
    Q, B = O.randQB_pb(A, 4, 3, 1, 1, seed=777)
    assert np.allclose(Q @ B, g["qbpb_QB"], atol=1e-11)
    assert O.estimate_rank_and_buildQ(A, 0.5, 1e-3, seed=777)[0] == int(g["rank1"])


def test_legacy_svd_and_cur_twins_match_golden_baselines():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_baselines.npz"))
    A = g["A"]
    for name, args, tol in [("randomized_low_rank_svd1", (8,), 1e-7), ("randomized_low_rank_svd2", (8,), 1e-11),
                            ("randomized_low_rank_svd3", (8, 3, 1), 1e-11), ("randomized_low_rank_svd4", (4, 2, 1), 1e-7)]:
        S = getattr(O, name)(A, *args, seed=777)[1]
        ref = np.diag(g[name.replace("randomized_low_rank_", "") + "_S"])
        assert np.max(np.abs(np.diag(S) - ref)) / ref.max() < tol, name
    f, Cm, U, R = O.cur_decomp_fixed_rank_or_prec(A, 9, 0.0)
    assert np.array_equal(Cm, g["cur_k9_C"]) and np.array_equal(R, g["cur_k9_R"])
    assert np.linalg.norm(Cm @ U @ R - g["cur_k9_C"] @ g["cur_k9_U"] @ g["cur_k9_R"]) <= 1e-9 * np.linalg.norm(A)
