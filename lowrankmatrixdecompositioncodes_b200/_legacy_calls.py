"""ctypes call wrappers for the deterministic baselines and legacy entry points of the reference API
(rank_revealing_algorithms_intel_mkl.h:3,12-37,44-57,65,75,85; matrix_vector_functions_intel_mkl.h:342-345).
A mixin: the host class provides lib, I, Mat, Vec, to_mat, from_mat, from_vec, set_seed — so the same wrappers serve this
package's Api and the parity tests' binding of the compiled reference."""
import ctypes as C


def bind(L, I, PM, PV):
    PPM, PPV, PI = C.POINTER(PM), C.POINTER(PV), C.POINTER(I)
    L.low_rank_svd_decomp_fixed_rank_or_prec.argtypes = [PM, I, C.c_double, PI, PPM, PPM, PPM]
    L.pivoted_QR_of_specified_rank.argtypes = [PM, I, PI, PPM, PPM, PPV]
    L.pivoted_QR_of_specified_rank_or_prec.argtypes = [PM, I, C.c_double, PI, PPM, PPM, PPV]
    L.id_decomp_fixed_rank_or_prec.argtypes = [PM, I, C.c_double, PI, PPV, PPM]
    L.id_two_sided_decomp_fixed_rank_or_prec.argtypes = [PM, I, C.c_double, PI, PPV, PPV, PPM, PPM]
    L.cur_decomp_fixed_rank_or_prec.argtypes = [PM, I, C.c_double, PI, PPM, PPM, PPM]
    L.randomized_low_rank_svd1.argtypes = [PM, I, PPM, PPM, PPM]
    L.randomized_low_rank_svd2.argtypes = [PM, I, PPM, PPM, PPM]
    L.randomized_low_rank_svd3.argtypes = [PM, I, I, I, PPM, PPM, PPM]
    L.randomized_low_rank_svd4.argtypes = [PM, I, I, I, PPM, PPM, PPM]
    L.randomized_low_rank_svd2_autorank1.argtypes = [PM, C.c_double, C.c_double, PPM, PPM, PPM]
    L.randomized_low_rank_svd2_autorank2.argtypes = [PM, I, C.c_double, PPM, PPM, PPM]
    L.randomized_low_rank_svd3_autorank2.argtypes = [PM, I, C.c_double, I, I, PPM, PPM, PPM]
    L.randQB_p.argtypes = [PM, I, I, PPM, PPM]
    L.randQB_pb.argtypes = [PM, I, I, I, I, PPM, PPM]
    L.estimate_rank_and_buildQ.argtypes = [PM, C.c_double, C.c_double, PPM, PI]
    L.estimate_rank_and_buildQ2.argtypes = [PM, I, C.c_double, PPM, PPM, PI]
    L.build_orthonormal_basis_from_mat.argtypes = [PM, PM]
    L.singular_value_decomposition.argtypes = [PM, PM, PM, PM]
    L.matrix_new.restype = PM
    L.matrix_new.argtypes = [I, I]


class LegacyCalls:
    def _pm(self):
        return C.POINTER(self.Mat)

    def _usv(self, fn, M, *args, frank=False):
        U, S, V = self._pm()(), self._pm()(), self._pm()()
        fr = self.I(0)
        tail = ([C.byref(fr)] if frank else []) + [C.byref(U), C.byref(S), C.byref(V)]
        fn(M, *args, *tail)
        self.lib.matrix_delete(M)
        out = (self.from_mat(U), self.from_mat(S), self.from_mat(V))
        return ((int(fr.value),) + out) if frank else out

    def svd_decomp(self, A, k, TOL):
        return self._usv(self.lib.low_rank_svd_decomp_fixed_rank_or_prec, self.to_mat(A), k, float(TOL), frank=True)

    def svd1(self, A, k, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd1, self.to_mat(A), k)

    def svd2(self, A, k, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd2, self.to_mat(A), k)

    def svd3(self, A, k, q, s, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd3, self.to_mat(A), k, q, s)

    def svd4(self, A, kstep, nstep, p, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd4, self.to_mat(A), kstep, nstep, p)

    def svd2_autorank1(self, A, frac, TOL, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd2_autorank1, self.to_mat(A), float(frac), float(TOL))

    def svd2_autorank2(self, A, kblock, TOL, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd2_autorank2, self.to_mat(A), kblock, float(TOL))

    def svd3_autorank2(self, A, kblock, TOL, q, s, seed=777):
        self.set_seed(seed)
        return self._usv(self.lib.randomized_low_rank_svd3_autorank2, self.to_mat(A), kblock, float(TOL), q, s)

    def pqr(self, A, k, TOL=None):
        """pivoted_QR_of_specified_rank (TOL is None) or pivoted_QR_of_specified_rank_or_prec -> frank, Qk, Rk, I"""
        M = self.to_mat(A)
        Q, R, I_ = self._pm()(), self._pm()(), C.POINTER(self.Vec)()
        fr = self.I(0)
        if TOL is None:
            self.lib.pivoted_QR_of_specified_rank(M, k, C.byref(fr), C.byref(Q), C.byref(R), C.byref(I_))
        else:
            self.lib.pivoted_QR_of_specified_rank_or_prec(M, k, float(TOL), C.byref(fr), C.byref(Q), C.byref(R), C.byref(I_))
        self.lib.matrix_delete(M)
        return int(fr.value), self.from_mat(Q), self.from_mat(R), self.from_vec(I_)

    def id_decomp(self, A, k, TOL):
        M = self.to_mat(A)
        I_, T = C.POINTER(self.Vec)(), self._pm()()
        fr = self.I(0)
        self.lib.id_decomp_fixed_rank_or_prec(M, k, float(TOL), C.byref(fr), C.byref(I_), C.byref(T))
        self.lib.matrix_delete(M)
        return int(fr.value), self.from_vec(I_), self.from_mat(T)

    def id_two_sided_decomp(self, A, k, TOL):
        M = self.to_mat(A)
        PV = C.POINTER(self.Vec)
        Ic, Ir, T, S = PV(), PV(), self._pm()(), self._pm()()
        fr = self.I(0)
        self.lib.id_two_sided_decomp_fixed_rank_or_prec(M, k, float(TOL), C.byref(fr), C.byref(Ic), C.byref(Ir), C.byref(T), C.byref(S))
        self.lib.matrix_delete(M)
        return int(fr.value), self.from_vec(Ic), self.from_vec(Ir), self.from_mat(T), self.from_mat(S)

    def cur_decomp(self, A, k, TOL):
        M = self.to_mat(A)
        Cm, U, R = self._pm()(), self._pm()(), self._pm()()
        fr = self.I(0)
        self.lib.cur_decomp_fixed_rank_or_prec(M, k, float(TOL), C.byref(fr), C.byref(Cm), C.byref(U), C.byref(R))
        self.lib.matrix_delete(M)
        return int(fr.value), self.from_mat(Cm), self.from_mat(U), self.from_mat(R)

    def randQB_p(self, A, k, p, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Q, B = self._pm()(), self._pm()()
        self.lib.randQB_p(M, k, p, C.byref(Q), C.byref(B))
        self.lib.matrix_delete(M)
        return self.from_mat(Q), self.from_mat(B)

    def randQB_pb(self, A, kstep, nstep, p, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Q, B = self._pm()(), self._pm()()
        self.lib.randQB_pb(M, kstep, nstep, p, s, C.byref(Q), C.byref(B))
        self.lib.matrix_delete(M)
        return self.from_mat(Q), self.from_mat(B)

    def estimate_rank1(self, A, frac, TOL, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Q = self._pm()()
        r = self.I(0)
        self.lib.estimate_rank_and_buildQ(M, float(frac), float(TOL), C.byref(Q), C.byref(r))
        self.lib.matrix_delete(M)
        return int(r.value), self.from_mat(Q)

    def estimate_rank2(self, A, kblock, TOL, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Y, Q = self._pm()(), self._pm()()
        r = self.I(0)
        self.lib.estimate_rank_and_buildQ2(M, kblock, float(TOL), C.byref(Y), C.byref(Q), C.byref(r))
        self.lib.matrix_delete(M)
        return int(r.value), self.from_mat(Y), self.from_mat(Q)

    def gesvd(self, A):
        """singular_value_decomposition (dgesvd 'S','S'): U m x r, S r x r, Vt r x n"""
        M = self.to_mat(A)
        m, n = A.shape
        r = min(m, n)
        U, S, Vt = self.lib.matrix_new(m, r), self.lib.matrix_new(r, r), self.lib.matrix_new(r, n)
        self.lib.singular_value_decomposition(M, U, S, Vt)
        self.lib.matrix_delete(M)
        return self.from_mat(U), self.from_mat(S), self.from_mat(Vt)
