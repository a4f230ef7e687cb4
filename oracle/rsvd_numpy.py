"""oracle/rsvd_numpy.py — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + scipy LAPACK) of the reference's randomized range-finder / QB hot path.
Never imported by the product package (lowrankmatrixdecompositioncodes_b200/); only tests/,
bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it, and only as
the checker.

PINNING: every function here is checked in tests/test_oracle.py against (i) outputs of the
UNMODIFIED reference C code compiled in this container (oracle/_ref/libref32.so, see
oracle/build_ref.sh) on the same inputs and the same Omega, and (ii) the golden vectors committed
under tests/golden/ that were produced by that reference build (tests/golden/make_golden.py).
The reference itself ships no golden vectors for this path, so the pin is
"reference C code + OpenBLAS 0.3.15 LAPACK, zero-initialising malloc, shared Philox Omega";
parity w.r.t. Intel MKL's own LAPACK is unpinned (MKL is not available offline).

All file:line citations are relative to /root/reference/multi_core_mkl_code/ :
  RRA = rank_revealing_algorithms_intel_mkl.c,  MVF = matrix_vector_functions_intel_mkl.c
"""
import numpy as np
from scipy.linalg import lapack

from . import ref_lib


# ------------------------------------------------------------------------------------------------
# primitives (MVF)
# ------------------------------------------------------------------------------------------------
def initialize_random_matrix(nrows, ncols, seed=777):
    """MVF:458-486 — float32 Gaussians filled in linear column-major order, widened to double.
    The VSL stream is replaced by the shared Philox map (seed, linear index) -> value
    (include/rsvd_b200_rng.h), the same one oracle/shim/vsl_shim.c feeds to the C reference."""
    z = ref_lib.normal_stream(seed, 0, nrows * ncols)
    return z.reshape((nrows, ncols), order="F")


def QR_factorization_getQ(M):
    """MVF:1251-1263 — dgeqrf + dorgqr, explicit thin Q (LAPACK sign convention)."""
    m, n = M.shape
    qr, tau, _, info = lapack.dgeqrf(np.asfortranarray(M))
    assert info == 0
    q, _, info = lapack.dorgqr(qr[:, :n] if m >= n else qr, tau)
    assert info == 0
    return q


def compact_QR_factorization(M):
    """MVF:1214-1245 — thin Q (m x n) and upper-triangular R (k x k), k = min(m,n)."""
    m, n = M.shape
    k = min(m, n)
    qr, tau, _, info = lapack.dgeqrf(np.asfortranarray(M))
    assert info == 0
    R = np.triu(qr[:k, :k])
    q, _, info = lapack.dorgqr(qr, tau)
    assert info == 0
    return q, R


def singular_value_decomposition(M):
    """MVF:1270-1284 — dgesvd 'S','S'; S returned as a full diagonal matrix, descending."""
    u, s, vt, info = lapack.dgesvd(np.asfortranarray(M), compute_uv=1, full_matrices=0)
    assert info == 0
    return u, np.diag(s), vt


def compute_evals_and_evecs_of_symm_matrix(S):
    """MVF:1206-1209 — dsyev 'V','U', ascending eigenvalues."""
    w, v, info = lapack.dsyev(np.asfortranarray(S), compute_v=1, lower=0)
    assert info == 0
    return w, v


def pivotedQR_mkl(M):
    """RRA:924-976 — dgeqp3 (all columns free: the zero-malloc fix of SURVEY.md Q4), R upper
    trapezoidal k x n (m<=n) or k x k, I = jpvt-1 stored as doubles.  The explicit Q the reference
    also forms (RRA:957-964) is never read on the hot path and is not restated."""
    m, n = M.shape
    k = min(m, n)
    qr, jpvt, tau, _, info = lapack.dgeqp3(np.asfortranarray(M))
    assert info == 0
    Rcols = n if m <= n else k
    R = np.triu(qr[:k, :Rcols])
    return R, (jpvt - 1).astype(np.float64)


def upper_triangular_system_solve(Rk1, Rk2):
    """MVF:1477-1493 (solve_type 1) — dtrsm Left/Upper/NoTrans/NonUnit."""
    x, info = lapack.dtrtrs(np.asfortranarray(Rk1), np.asfortranarray(Rk2), lower=0, trans=0, unitdiag=0)
    assert info == 0
    return x


def square_matrix_system_solve(A, B):
    """MVF:1525-1531 — dgesv."""
    _, _, x, info = lapack.dgesv(np.asfortranarray(A), np.asfortranarray(B))
    assert info == 0
    return x


def get_percent_error_between_two_mats(A, B):
    """MVF:391-405 — 100*||A-B||_F/||A||_F."""
    return 100.0 * np.linalg.norm(A - B) / np.linalg.norm(A)


# ------------------------------------------------------------------------------------------------
# SVD tail shared by RRA:133-225 and RRA:289-380
# ------------------------------------------------------------------------------------------------
def _svd_tail(M, Q, k, vnum):
    l = Q.shape[1]
    if vnum == 1 or vnum > 2:
        Bt = M.T @ Q                                   # RRA:139
        Qhat, Rhat = compact_QR_factorization(Bt)      # RRA:146
        Uhat, S, Vhat_t = singular_value_decomposition(Rhat)  # RRA:152
        U = Q @ Vhat_t.T                               # RRA:156
        V = Qhat @ Uhat                                # RRA:160
        return U[:, :k], S[:k, :k], V[:, :k]           # RRA:171-174
    B = Q.T @ M                                        # RRA:180
    BBt = B @ B.T                                      # RRA:183
    evals, Uhat = compute_evals_and_evecs_of_symm_matrix(np.triu(BBt) + np.triu(BBt, 1).T)  # RRA:189-190
    sing = np.sqrt(evals)                              # RRA:196-198
    S = np.diag(sing)
    U = Q @ Uhat                                       # RRA:203
    V = B.T @ (Uhat @ np.diag(1.0 / sing))             # RRA:208-212
    return U[:, l - k:], S[l - k:, l - k:], V[:, l - k:]  # RRA:220-223 (ascending, last k)


def low_rank_svd_rand_decomp_fixed_rank(M, k, p, vnum=1, q=2, s=1, seed=777, omega=None):
    """RRA:73-234.  NOTE the power loop runs j = 1 .. q-1 (RRA:101)."""
    m, n = M.shape
    l = k + p
    RN = initialize_random_matrix(n, l, seed) if omega is None else omega   # RRA:90-91
    Y = M @ RN                                                              # RRA:95
    for j in range(1, q):                                                   # RRA:101
        if (2 * j - 2) % s == 0:
            Z = M.T @ QR_factorization_getQ(Y)                              # RRA:106-108
        else:
            Z = M.T @ Y                                                     # RRA:112
        if (2 * j - 1) % s == 0:
            Y = M @ QR_factorization_getQ(Z)                                # RRA:118-120
        else:
            Y = M @ Z                                                       # RRA:124
    Q = QR_factorization_getQ(Y)                                            # RRA:129-130
    return _svd_tail(M, Q, k, vnum)


def randQB_pb_new(M, kstep, nstep, TOL, q, s, seed=777, omega=None):
    """RRA:1576-1801.  Power loop j = 1 .. q (RRA:1652); re-orthogonalisation against previous
    blocks only on even steps > 0 (RRA:1703); absolute Frobenius tolerance (RRA:1773-1775)."""
    m, n = M.shape
    if kstep > int(min(m, n) / 2):                                          # RRA:1589-1592
        kstep = int(min(m, n) / 10)
    tolMode = nstep <= 0
    if tolMode:
        nstep = int(min(m, n) / kstep)                                      # RRA:1595-1602
    l = kstep * nstep
    RN = initialize_random_matrix(n, l, seed) if omega is None else omega   # RRA:1607-1609
    A = M.copy()                                                            # RRA:1630
    Q = np.zeros((m, l))
    B = np.zeros((l, n))
    frank = 0
    for step in range(nstep):                                               # RRA:1635
        sl = slice(kstep * step, kstep * (step + 1))
        Yp = A @ RN[:, sl]                                                  # RRA:1643-1644
        for j in range(1, q + 1):                                           # RRA:1652
            if (2 * j - 2) % s == 0:
                Qp = QR_factorization_getQ(Yp)                              # RRA:1655
                AtQp = (Qp.T @ A).T                                         # RRA:1657-1658
            else:
                AtQp = A.T @ Yp                                             # RRA:1665
            if (2 * j - 1) % s == 0:
                Yp = A @ QR_factorization_getQ(AtQp)                        # RRA:1673-1674
            else:
                Yp = A @ AtQp                                               # RRA:1681
        Qp = QR_factorization_getQ(Yp)                                      # RRA:1690
        if step > 0 and step % 2 == 0:                                      # RRA:1703
            Qj = Q[:, :step * kstep]
            Qp = QR_factorization_getQ(Qp - Qj @ (Qj.T @ Qp))               # RRA:1716-1722
        Bp = Qp.T @ A                                                       # RRA:1741
        A = A - Qp @ Bp                                                     # RRA:1750-1751
        Q[:, sl] = Qp                                                       # RRA:1760
        B[sl, :] = Bp                                                       # RRA:1761
        frank = (step + 1) * kstep                                          # RRA:1770
        if tolMode and np.linalg.norm(A) < TOL:                             # RRA:1771-1777
            break
    if tolMode:
        Q = Q[:, :frank]                                                    # RRA:1784-1785
        B = B[:frank, :]
    return frank, Q, B


def low_rank_svd_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, vnum, kstep, q, s, seed=777):
    """RRA:239-381, INCLUDING quirk Q1: nstep=0 set for k<=0 at RRA:253 is
    overwritten at RRA:263, so the routine never reaches tolerance mode."""
    m, n = M.shape
    rankMode = k > 0
    if p < kstep and (p + kstep) < min(m, n):                               # RRA:260-262
        p = kstep
    nstep = int((k + p) // kstep)                                           # RRA:263 (integer division)
    frank, Q, B = randQB_pb_new(M, kstep, nstep, TOL, q, s, seed)           # RRA:266
    if rankMode:
        frank = k                                                           # RRA:275
    else:
        frank = int(round((frank / (frank + p + 1e-6)) * frank))            # RRA:278
    U, S, V = _svd_tail(M, Q, frank, vnum)                                  # RRA:289-380
    return frank, U, S, V


def id_rand_decomp_fixed_rank(M, k, p, q, s, seed=777, omega=None):
    """RRA:1863-1965.  LEFT sketch RN = (k+p) x m; power loop j = 1 .. q (RRA:1882)."""
    m, n = M.shape
    l = k + p
    RN = initialize_random_matrix(l, m, seed) if omega is None else omega   # RRA:1871-1872
    Y = RN @ M                                                              # RRA:1877
    for j in range(1, q + 1):                                               # RRA:1882
        Z = QR_factorization_getQ(Y.T).T if (2 * j - 2) % s == 0 else Y     # RRA:1889-1898
        Y = Z @ M.T                                                         # RRA:1908
        Z = QR_factorization_getQ(Y.T).T if (2 * j - 1) % s == 0 else Y     # RRA:1912-1921
        Y = Z @ M                                                           # RRA:1927
    Rd, I = pivotedQR_mkl(Y)                                                # RRA:1938
    Rk = Rd[:k, :]                                                          # RRA:1943
    Rk1 = np.triu(Rk[:, :k])                                                # RRA:1947,1955
    Rk2 = Rk[:, k:]                                                         # RRA:1949
    T = upper_triangular_system_solve(Rk1, Rk2)                             # RRA:1956
    return I, T


def id_decomp_full(M, k):
    """RRA:1807-1854 as reached from the two-sided ID (k == min(m,n) branch, RRA:1830-1834)."""
    Rk, I = pivotedQR_mkl(M)
    Rk1 = np.triu(Rk[:k, :k])
    Rk2 = Rk[:k, k:]
    return I, upper_triangular_system_solve(Rk1, Rk2)


def id_two_sided_rand_decomp_fixed_rank(M, k, p, q, s, seed=777):
    """RRA:2060-2082."""
    Icol, T = id_rand_decomp_fixed_rank(M, k, p, q, s, seed)                # RRA:2068
    MI = M[:, Icol[:k].astype(np.int64)]                                    # RRA:2073
    Irow, S = id_decomp_full(np.ascontiguousarray(MI.T), k)                 # RRA:2074-2078
    return Icol, Irow, T, S


def cur_rand_decomp_fixed_rank(M, k, p, q, s, seed=777):
    """RRA:2191-2258."""
    Icol, Irow, T, S = id_two_sided_rand_decomp_fixed_rank(M, k, p, q, s, seed)   # RRA:2198
    n = M.shape[1]
    Icolinv = np.empty(n, dtype=np.int64)
    Icolinv[Icol.astype(np.int64)] = np.arange(n)                           # RRA:2205-2206, MVF:1195-1201
    V1 = np.vstack([np.eye(k), T.T])                                        # RRA:2209-2214
    V = V1[Icolinv, :]                                                      # RRA:2218-2219
    R = M[Irow[:k].astype(np.int64), :]                                     # RRA:2230-2231
    Cm = M[:, Icol[:k].astype(np.int64)]                                    # RRA:2236-2237
    Ut = square_matrix_system_solve(R @ R.T, R @ V)                         # RRA:2247-2250
    return Cm, Ut.T, R                                                      # RRA:2252


def id_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, kstep, q, s, seed=777):
    """RRA:1969-2027 — column ID from the pivoted QR of the QB factor B."""
    nstep = int((k + p) // kstep)                                           # RRA:1974
    rankMode = k > 0
    if not rankMode:
        nstep = 0                                                           # RRA:1980-1983
    frank, Q, B = randQB_pb_new(M, kstep, nstep, TOL, q, s, seed)           # RRA:1989
    Rk, I = pivotedQR_mkl(B)                                                # RRA:1998
    frank = k if rankMode else int(round((frank / (frank + p + 1e-6)) * frank))   # RRA:2001-2006
    T = upper_triangular_system_solve(np.triu(Rk[:frank, :frank]), Rk[:frank, frank:])   # RRA:2009-2020
    return frank, I, T


def id_two_sided_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, kstep, q, s, seed=777):
    """RRA:2086-2111."""
    frank, Icol, T = id_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, kstep, q, s, seed)
    MI = M[:, Icol[:frank].astype(np.int64)]
    Irow, S = id_decomp_full(np.ascontiguousarray(MI.T), frank)
    return frank, Icol, Irow, T, S


def cur_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, kstep, q, s, seed=777):
    """RRA:2262-2332."""
    frank, Icol, Irow, T, S = id_two_sided_blockrand_decomp_fixed_rank_or_prec(M, k, p, TOL, kstep, q, s, seed)
    n = M.shape[1]
    Icolinv = np.empty(n, dtype=np.int64)
    Icolinv[Icol.astype(np.int64)] = np.arange(n)
    V = np.vstack([np.eye(frank), T.T])[Icolinv, :]
    R = M[Irow[:frank].astype(np.int64), :]
    Cm = M[:, Icol[:frank].astype(np.int64)]
    Ut = square_matrix_system_solve(R @ R.T, R @ V)
    return frank, Cm, Ut.T, R


def low_rank_svd_rand_decomp_fromQB(Q, B):
    """oneapi_code/rank_revealing_algorithms_one_api.c:244-304, restated in FP64 (the reference ships it in float32 only):
    SVD of B B^T (descending), S = sqrt, U = Q Uhat, V = B^T Uhat S^{-1}."""
    Uhat, St, _ = singular_value_decomposition(B @ B.T)
    sing = np.sqrt(np.diag(St))
    return Q @ Uhat, np.diag(sing), B.T @ (Uhat / sing)


def id_rand_decomp_fromQB(Q, B):
    """oneapi_code/rank_revealing_algorithms_one_api.c:421-444 in FP64: pivoted QR of B, k = cols(Q)."""
    k = Q.shape[1]
    Rk, I = pivotedQR_mkl(B)
    return I, upper_triangular_system_solve(np.triu(Rk[:k, :k]), Rk[:k, k:])


# ------------------------------------------------------------------------------------------------
# synthetic inputs + binary file formats
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# deterministic baselines and legacy entry points
# ------------------------------------------------------------------------------------------------
def pivoted_QR_of_specified_rank_or_prec(M, k, TOL=None):
    """RRA:1159-1334 (TOL given) / RRA:1012-1155 (TOL None): the reference's OWN partial Householder QR.
    Squared column norms (MVF:445-454), first index of the maximum by strict '>', reflector v = x - ||x|| e_i scaled to
    ||v||^2 = 2 (RRA:982-1008, so R(i,i) = +||x||), plain downdate norm -= R(i,j)^2 (RRA:1316-1319); stop when the largest
    remaining squared norm is 0 (fixed-rank variant, RRA:1066) / below 1e-10 in magnitude (RRA:1269) for i > 0, or — tolerance
    mode, k <= 0 — when sqrt(sum of remaining squared norms)/n, refreshed every 5th step only, is below TOL (RRA:1243-1275).
    Returns frank, Qk (m x frank), Rk (frank x n), I (n, 0-based)."""
    m, n = M.shape
    tol_mode = k <= 0
    kmax = min(m, n) if tol_mode else min(k, min(m, n))
    R = np.array(M, dtype=np.float64, order="F")
    Q = np.eye(m)
    norms = np.sum(R * R, axis=0)
    I = np.arange(n, dtype=np.float64)
    frank, r22 = 0, 0.0
    for i in range(kmax):
        j = i + int(np.argmax(norms[i:]))                   # first index of the maximum
        if i % 5 == 0:
            r22 = np.sqrt(np.sum(norms[i:])) / n if tol_mode else 0.0
        zero = (norms[j] == 0.0) if TOL is None else (abs(norms[j]) < 1e-10)
        if zero and i > 0:
            break
        if tol_mode and TOL is not None and r22 < TOL:
            break
        frank = i + 1
        I[[i, j]] = I[[j, i]]
        R[:, [i, j]] = R[:, [j, i]]
        norms[[i, j]] = norms[[j, i]]
        v = np.zeros(m)
        v[i:] = R[i:, i]
        v[i] -= np.linalg.norm(R[i:, i])
        nv = np.linalg.norm(v)
        if nv > 0:
            v *= np.sqrt(2.0) / nv
        R -= np.outer(v, v @ R)
        Q -= np.outer(Q @ v, v)
        if i != n - 1:
            norms[i + 1:] -= R[i, i + 1:] ** 2
    return frank, Q[:, :frank].copy(), R[:frank, :].copy(), I


def id_decomp_fixed_rank_or_prec(M, k, TOL):
    """RRA:1807-1854: partial pivoted QR when k < min(m,n) (or tolerance mode), dgeqp3 otherwise; T = triu(Rk1)^{-1} Rk2."""
    m, n = M.shape
    if k < min(m, n):
        frank, _, Rk, I = pivoted_QR_of_specified_rank_or_prec(M, k, TOL)
    else:
        frank = k
        Rk, I = pivotedQR_mkl(M)
    T = upper_triangular_system_solve(np.triu(Rk[:frank, :frank]), Rk[:frank, frank:])
    return frank, I, T


def id_two_sided_decomp_fixed_rank_or_prec(M, k, TOL):
    """RRA:2034-2056"""
    frank, Icol, T = id_decomp_fixed_rank_or_prec(M, k, TOL)
    MI = M[:, Icol[:frank].astype(int)]
    frank, Irow, S = id_decomp_fixed_rank_or_prec(np.ascontiguousarray(MI.T), frank, 0.0)
    return frank, Icol, Irow, T, S


def low_rank_svd_decomp_fixed_rank_or_prec(M, k, TOL):
    """RRA:7-69: dgesvd, then truncation to k or — k <= 0 — to the first i with abs((int)sigma_i) < TOL (the C integer abs
    of RRA:49)."""
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    r = len(s)
    frank = k if k > 0 else r
    if k <= 0:
        for i in range(r):
            if abs(int(s[i])) < TOL and i < r - 1:
                frank = i + 1
                break
    return frank, U[:, :frank], np.diag(s[:frank]), Vt[:frank].T


def randQB_p(M, k, p, seed=777):
    """RRA:1343-1421: single-vector randQB; the Gram-Schmidt loop runs over i < j-1 (RRA:1387)."""
    m, n = M.shape
    RN = initialize_random_matrix(n, k, seed)
    A = np.array(M, dtype=np.float64)
    Q, B = np.zeros((m, k)), np.zeros((k, n))
    for j in range(k):
        y = A @ RN[:, j]
        for _ in range(p):
            y = A @ (A.T @ y)
        for i in range(j - 1):
            y = y - (Q[:, i] @ y) * Q[:, i]
        q = y / np.linalg.norm(y)
        b = A.T @ q
        Q[:, j], B[j] = q, b
        A -= np.outer(q, b)
    return Q, B


def randQB_pb(M, kstep, nstep, p, s, seed=777):
    """RRA:1425-1572: blocked randQB, power loop j <= p, re-orthogonalisation against all previous blocks on EVERY step > 0."""
    m, n = M.shape
    l = kstep * nstep
    RN = initialize_random_matrix(n, l, seed)
    A = np.array(M, dtype=np.float64)
    Q, B = np.zeros((m, l)), np.zeros((l, n))
    for step in range(nstep):
        sl = slice(kstep * step, kstep * (step + 1))
        Yp = A @ RN[:, sl]
        for j in range(1, p + 1):
            W = A.T @ (QR_factorization_getQ(Yp) if (2 * j - 2) % s == 0 else Yp)
            Yp = A @ (QR_factorization_getQ(W) if (2 * j - 1) % s == 0 else W)
        Qp = QR_factorization_getQ(Yp)
        if step > 0:
            Qj = Q[:, :kstep * step]
            Qp = QR_factorization_getQ(Qp - Qj @ (Qj.T @ Qp))
        Bp = Qp.T @ A
        A -= Qp @ Bp
        Q[:, sl], B[sl] = Qp, Bp
    return Q, B


def estimate_rank_and_buildQ(M, frac_of_max_rank, TOL, seed=777):
    """MVF:1339-1400: Y = M RN, sequential modified Gram-Schmidt; stop at column j when two consecutive projections (p1
    persists across columns, starts at 0) are both shorter than TOL; Q = orthonormal basis of the first good_rank columns."""
    m, n = M.shape
    maxdim = int(round(min(m, n) * frac_of_max_rank))
    Qb = M @ initialize_random_matrix(n, maxdim, seed)
    good, p1 = maxdim, 0.0
    stop = False
    for j in range(maxdim):
        vj = Qb[:, j].copy()
        for i in range(j):
            vi = Qb[:, i]
            coef = (vj @ vi) / (vi @ vi)
            vj -= coef * vi
            pn = abs(coef) * np.linalg.norm(vi)
            if pn < TOL and p1 < TOL:
                good, stop = j, True
                break
            p1 = pn
        if stop:
            break
        Qb[:, j] = vj / np.linalg.norm(vj)
    return good, QR_factorization_getQ(Qb[:, :good])


def _cur_tail(M, Icol, Irow, T, k):
    """RRA:2131-2177 / 2200-2252: C, U, R from a two-sided ID."""
    n = M.shape[1]
    Icolinv = np.empty(n, dtype=np.int64)
    Icolinv[Icol.astype(np.int64)] = np.arange(n)
    V = np.vstack([np.eye(k), T.T])[Icolinv, :]
    R = M[Irow[:k].astype(np.int64), :]
    Cm = M[:, Icol[:k].astype(np.int64)]
    return Cm, square_matrix_system_solve(R @ R.T, R @ V).T, R


def cur_decomp_fixed_rank_or_prec(M, k, TOL):
    """RRA:2115-2187"""
    frank, Icol, Irow, T, S = id_two_sided_decomp_fixed_rank_or_prec(M, k, TOL)
    return (frank,) + _cur_tail(M, Icol, Irow, T, frank)


def randomized_low_rank_svd1(M, k, seed=777):
    """RRA:385-461 = the eig(B B^T) variant without oversampling or power iterations"""
    return low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 2, 1, 1, seed)


def randomized_low_rank_svd2(M, k, seed=777):
    """RRA:465-532"""
    return low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 1, 1, 1, seed)


def randomized_low_rank_svd3(M, k, q, s, seed=777):
    """RRA:536-634 (power loop j < q, the same orthonormalisation schedule as RRA:101-125)"""
    return low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 1, q, s, seed)


def randomized_low_rank_svd4(M, kstep, nstep, p, seed=777):
    """RRA:638-697: randQB_pb(M, kstep, nstep, p, 1), eig of B B^T, ascending singular values"""
    Q, B = randQB_pb(M, kstep, nstep, p, 1, seed)
    BBt = B @ B.T
    evals, Uhat = compute_evals_and_evecs_of_symm_matrix(np.triu(BBt) + np.triu(BBt, 1).T)
    sing = np.sqrt(evals)
    return Q @ Uhat, np.diag(sing), B.T @ (Uhat @ np.diag(1.0 / sing))


def make_matrix(m, n, spectrum="logspace", seed=0, k=None, tail=1e-8, rho=1.0 / 1.02):
    """make_matrix_binary.m:11-25 — A = U diag(sigma) V^T with Haar U, V.
    spectrum: 'logspace' (the reference's logspace(1,-3,p)), 'exp' (sigma_i = rho^i, the oneAPI
    drivers' choice, oneapi_code/driver1.c:47-50), or 'gap' (logspace(0,-4,k) then
    tail*logspace(0,-2) — the matrix on which two different Omegas agree to 1e-13)."""
    rng = np.random.default_rng(seed)
    r = min(m, n)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    if spectrum == "logspace":
        sig = np.logspace(1, -3, r)
    elif spectrum == "exp":
        sig = rho ** np.arange(r)
    elif spectrum == "gap":
        sig = np.concatenate([np.logspace(0, -4, k), tail * np.logspace(0, -2, r - k)])
    else:
        raise ValueError(spectrum)
    return (U * sig) @ V.T, sig


def write_matrix_binary(A, fname, bits=32):
    """MVF:114-133 / 64-bit MVF:114-135: header (int32|int64) m, n; payload ROW-major doubles."""
    m, n = A.shape
    with open(fname, "wb") as f:
        np.array([m, n], dtype=np.int32 if bits == 32 else np.int64).tofile(f)
        np.ascontiguousarray(A, dtype=np.float64).tofile(f)


def read_matrix_binary(fname, bits=32):
    """MVF:77-101."""
    with open(fname, "rb") as f:
        m, n = np.fromfile(f, dtype=np.int32 if bits == 32 else np.int64, count=2)
        return np.fromfile(f, dtype=np.float64, count=int(m) * int(n)).reshape((int(m), int(n)))
