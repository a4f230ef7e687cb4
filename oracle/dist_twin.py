"""oracle/dist_twin.py — TEST INFRASTRUCTURE ONLY.

numpy restatement of the ROW-PARTITIONED form of the hot path exactly as csrc/device/pipeline.cu sequences it
(one process per rank, A_g = this rank's rows): which products stay local, which are summed across ranks, how Omega is
indexed by global column (SVD) / global row (ID).  `allreduce(x)` sums an ndarray over the ranks in place.
Used by tests/test_dist_cpu.py over a world_size-2 gloo group to check the sharding logic against the single-process twin
(oracle/rsvd_numpy.py).  CholeskyQR2 replaces Householder QR as on the device (same range, different basis)."""
import numpy as np
from scipy.linalg import lapack, solve_triangular

from . import ref_lib, rsvd_numpy as O


def cholqr2(Y, allreduce, sharded=True):
    """Q with orthonormal columns spanning range(Y) and R with Y = Q R; Gram matrices are all-reduced when Y is row-sharded."""
    R_acc = np.eye(Y.shape[1])
    Q = Y
    for _ in range(2):
        G = Q.T @ Q
        if sharded:
            allreduce(G)
        R = np.linalg.cholesky(G).T
        Q = solve_triangular(R, Q.T, trans="T", lower=False).T
        R_acc = R @ R_acc
    return Q, R_acc


def svd_rand_sharded(A_loc, k, p, q, s, seed, allreduce):
    """pipeline.cu: svd_rand + svd_from_q (vnum = 1).  Returns U_loc (rows of this rank), S, V (replicated)."""
    m_loc, n = A_loc.shape
    l = k + p
    Omega = O.initialize_random_matrix(n, l, seed)        # indexed by global column: identical on every rank
    Y = A_loc @ Omega
    for j in range(1, q):
        if (2 * j - 2) % s == 0:
            Y, _ = cholqr2(Y, allreduce, True)
        Z = A_loc.T @ Y
        allreduce(Z)                                      # sum of n x l partial products
        if (2 * j - 1) % s == 0:
            Z, _ = cholqr2(Z, allreduce, False)           # replicated panel: no communication
        Y = A_loc @ Z
    Q, _ = cholqr2(Y, allreduce, True)
    Bt = A_loc.T @ Q
    allreduce(Bt)
    Qhat, Rhat = cholqr2(Bt, allreduce, False)
    Uhat, sv, Vhat_t = np.linalg.svd(Rhat)
    return (Q @ Vhat_t.T)[:, :k], sv[:k], (Qhat @ Uhat)[:, :k]


def id_rand_sharded(A_loc, row0, m_global, k, p, q, s, seed, allreduce):
    """pipeline.cu: id_rand.  Left sketch: Omega is l x m_global, this rank uses columns row0 .. row0+m_loc."""
    m_loc, n = A_loc.shape
    l = k + p
    Omega = O.initialize_random_matrix(l, m_global, seed)
    Yt = A_loc.T @ Omega[:, row0:row0 + m_loc].T          # (Omega A)^T, n x l
    allreduce(Yt)
    for j in range(1, q + 1):
        if (2 * j - 2) % s == 0:
            Yt, _ = cholqr2(Yt, allreduce, False)
        Wt = A_loc @ Yt                                   # m_loc x l, row-sharded
        if (2 * j - 1) % s == 0:
            Wt, _ = cholqr2(Wt, allreduce, True)
        Yt = A_loc.T @ Wt
        allreduce(Yt)
    R, I = O.pivotedQR_mkl(np.ascontiguousarray(Yt.T))
    T = O.upper_triangular_system_solve(np.triu(R[:k, :k]), R[:k, k:])
    return I, T
