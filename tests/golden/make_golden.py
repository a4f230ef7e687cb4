"""tests/golden/make_golden.py — regenerates tests/golden/golden_small.npz.

Runs the UNMODIFIED reference C code (oracle/_ref/libref32.so and libref64.so, built from
/root/reference by oracle/build_ref.sh) on small seeded inputs with the shared Philox Omega and
stores inputs + outputs.  Needs /root/reference only through the prebuilt oracle/_ref libraries.
Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ref_lib, rsvd_numpy as O  # noqa: E402


def main():
    L = ref_lib.RefLib(32)
    L64 = ref_lib.RefLib(64)
    out = {}
    A, sig = O.make_matrix(60, 40, "logspace", seed=3)
    G, gsig = O.make_matrix(96, 64, "gap", seed=4, k=12, tail=1e-7)
    out["A"], out["A_sigma"], out["G"], out["G_sigma"] = A, sig, G, gsig
    out["omega_7x5_seed777"] = L.omega(7, 5, seed=777)
    out["normals_seed42_first1000_64"] = ref_lib.normal_stream(42, 1000, 64)
    # SVD, vnum 1 and 2
    for name, M, k, p, vnum, q, s in [("svdA_v1", A, 8, 4, 1, 2, 1), ("svdA_v2", A, 8, 4, 2, 3, 2),
                                      ("svdG_v1", G, 12, 6, 1, 2, 1)]:
        U, S, V = L.svd_rand(M, k, p, vnum, q, s, seed=777)
        out[name + "_U"], out[name + "_S"], out[name + "_V"] = U, S, V
        out[name + "_params"] = np.array([k, p, vnum, q, s, 777])
    U, S, V = L64.svd_rand(A, 8, 4, 1, 2, 1, seed=777)
    out["svdA_v1_S_64bit"] = S
    # QB, rank mode and tolerance mode
    f, Q, B = L.randQB_pb_new(A, 4, 3, 0.0, 2, 1, seed=777)
    out["qbA_rank_frank"], out["qbA_rank_Q"], out["qbA_rank_B"] = np.array(f), Q, B
    f, Q, B = L.randQB_pb_new(A, 4, 0, 2.0, 1, 1, seed=777)
    out["qbA_tol_frank"], out["qbA_tol_Q"], out["qbA_tol_B"] = np.array(f), Q, B
    # blockrand SVD incl. quirk Q1 (k=0 never reaches tolerance mode)
    f, U, S, V = L.svd_blockrand(A, 8, 4, 0.0, 1, 4, 2, 1, seed=777)
    out["blkA_rank_frank"], out["blkA_rank_S"] = np.array(f), S
    f, U, S, V = L.svd_blockrand(A, 0, 4, 1.0, 1, 4, 2, 1, seed=777)
    out["blkA_tol_frank"], out["blkA_tol_S"] = np.array(f), S
    # ID / two-sided ID / CUR
    I, T = L.id_rand(A, 8, 4, 2, 1, seed=777)
    out["idA_I"], out["idA_T"] = I, T
    Ic, Ir, T, S = L.id_two_sided_rand(A, 8, 4, 2, 1, seed=777)
    out["id2A_Icol"], out["id2A_Irow"], out["id2A_T"], out["id2A_S"] = Ic, Ir, T, S
    Cm, U, R = L.cur_rand(A, 8, 4, 2, 1, seed=777)
    out["curA_C"], out["curA_U"], out["curA_R"] = Cm, U, R
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "golden_small.npz"), **out)
    print("wrote golden_small.npz with", len(out), "arrays")


def main_baselines():
    """tests/golden/golden_baselines.npz: deterministic baselines and legacy entry points."""
    L = ref_lib.RefLib(32)
    out = {}
    A, _ = O.make_matrix(60, 40, "logspace", seed=3)
    out["A"] = A
    f, Q, R, I = L.pqr(A, 10)
    out["pqr_k10_frank"], out["pqr_k10_Q"], out["pqr_k10_R"], out["pqr_k10_I"] = np.array(f), Q, R, I
    f, Q, R, I = L.pqr(A, 0, 0.05)
    out["pqr_tol_frank"], out["pqr_tol_R"], out["pqr_tol_I"] = np.array(f), R, I
    f, I, T = L.id_decomp(A, 12, 0.0)
    out["id_k12_frank"], out["id_k12_I"], out["id_k12_T"] = np.array(f), I, T
    f, Ic, Ir, T, S = L.id_two_sided_decomp(A, 0, 0.05)
    out["id2_tol_frank"], out["id2_tol_Icol"], out["id2_tol_Irow"], out["id2_tol_T"], out["id2_tol_S"] = np.array(f), Ic, Ir, T, S
    f, Cm, U, R = L.cur_decomp(A, 9, 0.0)
    out["cur_k9_C"], out["cur_k9_U"], out["cur_k9_R"] = Cm, U, R
    f, U, S, V = L.svd_decomp(A, 7, 0.0)
    out["svd_k7_S"] = S
    out["svd_tol_frank"] = np.array(L.svd_decomp(A, 0, 2.0)[0])
    Q, B = L.randQB_p(A, 6, 1, seed=777)
    out["qbp_Q"], out["qbp_B"] = Q, B
    Q, B = L.randQB_pb(A, 4, 3, 1, 1, seed=777)
    out["qbpb_QB"] = Q @ B
    for name, args in [("svd1", (8,)), ("svd2", (8,)), ("svd3", (8, 3, 1)), ("svd4", (4, 2, 1))]:
        U, S, V = getattr(L, name)(A, *args, seed=777)
        out[name + "_S"] = S
    out["rank1"] = np.array(L.estimate_rank1(A, 0.5, 1e-3, seed=777)[0])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "golden_baselines.npz"), **out)
    print("wrote golden_baselines.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
    main_baselines()
