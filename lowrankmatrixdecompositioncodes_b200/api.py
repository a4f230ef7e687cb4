"""The reference's C API bound from Python exactly as a C driver would use it: mat/vec structs in host memory,
callee-allocated outputs, matrix_delete to free.  Mirrors rank_revealing_algorithms_intel_mkl.h:5-8,60,68,78,88
(argument order and meaning) so parity tests read like calls into the reference (oracle/ref_lib.RefLib has the same
method names)."""
import ctypes as C

import numpy as np

from . import native
from . import _legacy_calls


class Api(_legacy_calls.LegacyCalls):
    def __init__(self, bits=32):
        native.dev()  # device layer first (RTLD_GLOBAL)
        self.bits = bits
        self.I = C.c_int if bits == 32 else C.c_int64
        I = self.I

        class Mat(C.Structure):
            _fields_ = [("nrows", I), ("ncols", I), ("d", C.POINTER(C.c_double))]

        class Vec(C.Structure):
            _fields_ = [("nrows", I), ("d", C.POINTER(C.c_double))]

        self.Mat, self.Vec = Mat, Vec
        self.lib = native._load(native.API32_PATH if bits == 32 else native.API64_PATH)
        L = self.lib
        PM, PV = C.POINTER(Mat), C.POINTER(Vec)
        PPM, PPV = C.POINTER(PM), C.POINTER(PV)
        self.PM, self.PV = PM, PV
        L.matrix_new.restype = PM
        L.matrix_new.argtypes = [I, I]
        L.vector_new.restype = PV
        L.vector_new.argtypes = [I]
        L.matrix_delete.argtypes = [PM]
        L.vector_delete.argtypes = [PV]
        L.low_rank_svd_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.low_rank_svd_blockrand_decomp_fixed_rank_or_prec.argtypes = [
            PM, I, I, C.c_double, I, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.randQB_pb_new.argtypes = [PM, I, I, C.c_double, I, I, C.POINTER(I), PPM, PPM]
        L.id_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPV, PPM]
        L.id_decomp_fixed_rank_or_prec.argtypes = [PM, I, C.c_double, C.POINTER(I), PPV, PPM]
        L.id_two_sided_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPV, PPV, PPM, PPM]
        L.cur_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPM, PPM, PPM]
        L.id_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPV, PPM]
        L.id_two_sided_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPV, PPV, PPM, PPM]
        L.cur_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.low_rank_svd_rand_decomp_fromQB.argtypes = [PM, PM, PPM, PPM, PPM]
        L.id_rand_decomp_fromQB.argtypes = [PM, PM, PPV, PPM]
        L.pivotedQR_mkl.argtypes = [PM, PPM, PPM, PPV]
        L.matrix_load_from_binary_file.restype = PM
        L.matrix_load_from_binary_file.argtypes = [C.c_char_p]
        L.matrix_write_to_binary_file.argtypes = [PM, C.c_char_p]
        L.get_matrix_frobenius_norm.restype = C.c_double
        L.get_matrix_frobenius_norm.argtypes = [PM]
        L.get_percent_error_between_two_mats.restype = C.c_double
        L.get_percent_error_between_two_mats.argtypes = [PM, PM]
        L.form_svd_product_matrix.argtypes = [PM, PM, PM, PM]
        L.form_cur_product_matrix.argtypes = [PM, PM, PM, PM]
        L.initialize_random_matrix.argtypes = [PM]
        L.QR_factorization_getQ.argtypes = [PM, PM]
        L.compact_QR_factorization.argtypes = [PM, PM, PM]
        L.singular_value_decomposition.argtypes = [PM, PM, PM, PM]
        L.matrix_matrix_mult.argtypes = [PM, PM, PM]
        L.matrix_transpose_matrix_mult.argtypes = [PM, PM, PM]
        L.matrix_matrix_transpose_mult.argtypes = [PM, PM, PM]
        L.use_low_rank_svd_for_approximation.argtypes = [PM, PM, PM, PM]
        L.use_id_decomp_for_approximation.argtypes = [PM, PM, PV, I]
        L.use_id_two_sided_decomp_for_approximation.argtypes = [PM, PM, PM, PV, PV, I]
        L.use_cur_decomp_for_approximation.argtypes = [PM, PM, PM, PM]
        _legacy_calls.bind(L, I, PM, PV)
        L.rsvd_b200_api_status.restype = C.c_int
        L.rsvd_b200_api_last_error.restype = C.c_char_p
        L.rsvd_b200_api_last_percent_error.restype = C.c_double

    # ---- marshalling ----
    def set_seed(self, seed):
        native.dev().rsvd_b200_set_option(b"seed", int(seed))

    def check(self):
        """Raise if the last API call reported an error (out-of-band: the reference API is void); reading clears the status."""
        if self.lib.rsvd_b200_api_status():
            msg = self.lib.rsvd_b200_api_last_error().decode()
            self.lib.rsvd_b200_api_clear_error()
            raise RuntimeError("rsvd_b200 api: " + msg)

    def to_mat(self, a):
        a = np.asarray(a, dtype=np.float64)
        m, n = a.shape
        M = self.lib.matrix_new(m, n)
        if m * n:
            np.ctypeslib.as_array(M.contents.d, shape=(m * n,))[:] = np.asfortranarray(a).ravel(order="F")
        return M

    def to_vec(self, v):
        v = np.asarray(v, dtype=np.float64)
        V = self.lib.vector_new(len(v))
        if len(v):
            np.ctypeslib.as_array(V.contents.d, shape=(len(v),))[:] = v
        return V

    def from_mat(self, M, free=True):
        m, n = int(M.contents.nrows), int(M.contents.ncols)
        if m * n:
            out = np.ctypeslib.as_array(M.contents.d, shape=(m * n,)).copy().reshape((m, n), order="F")
        else:
            out = np.zeros((m, n))
        if free:
            self.lib.matrix_delete(M)
        return out

    def from_vec(self, v, free=True):
        n = int(v.contents.nrows)
        out = np.ctypeslib.as_array(v.contents.d, shape=(n,)).copy() if n else np.zeros(0)
        if free:
            self.lib.vector_delete(v)
        return out

    # ---- the hot-path API (reference semantics, numpy in/out) ----
    def svd_rand(self, A, k, p, vnum=1, q=2, s=1, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        U, S, V = self.PM(), self.PM(), self.PM()
        frank = self.I(0)
        self.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, vnum, q, s, C.byref(frank), C.byref(U), C.byref(S), C.byref(V))
        self.lib.matrix_delete(M)
        out = self.from_mat(U), self.from_mat(S), self.from_mat(V)
        self.check()
        return out

    def svd_blockrand(self, A, k, p, TOL, vnum, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        U, S, V = self.PM(), self.PM(), self.PM()
        frank = self.I(0)
        self.lib.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(
            M, k, p, float(TOL), vnum, kstep, q, s, C.byref(frank), C.byref(U), C.byref(S), C.byref(V))
        self.lib.matrix_delete(M)
        out = int(frank.value), self.from_mat(U), self.from_mat(S), self.from_mat(V)
        self.check()
        return out

    def randQB_pb_new(self, A, kstep, nstep, TOL, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Q, B = self.PM(), self.PM()
        frank = self.I(0)
        self.lib.randQB_pb_new(M, kstep, nstep, float(TOL), q, s, C.byref(frank), C.byref(Q), C.byref(B))
        self.lib.matrix_delete(M)
        out = int(frank.value), self.from_mat(Q), self.from_mat(B)
        self.check()
        return out

    def id_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        I_, T = self.PV(), self.PM()
        self.lib.id_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(I_), C.byref(T))
        self.lib.matrix_delete(M)
        out = self.from_vec(I_), self.from_mat(T)
        self.check()
        return out

    def id_two_sided_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Ic, Ir, T, S = self.PV(), self.PV(), self.PM(), self.PM()
        self.lib.id_two_sided_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(Ic), C.byref(Ir), C.byref(T), C.byref(S))
        self.lib.matrix_delete(M)
        out = self.from_vec(Ic), self.from_vec(Ir), self.from_mat(T), self.from_mat(S)
        self.check()
        return out

    def cur_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Cm, U, R = self.PM(), self.PM(), self.PM()
        self.lib.cur_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(Cm), C.byref(U), C.byref(R))
        self.lib.matrix_delete(M)
        out = self.from_mat(Cm), self.from_mat(U), self.from_mat(R)
        self.check()
        return out


    def id_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        I_, T = self.PV(), self.PM()
        frank = self.I(0)
        self.lib.id_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(I_), C.byref(T))
        self.lib.matrix_delete(M)
        out = int(frank.value), self.from_vec(I_), self.from_mat(T)
        self.check()
        return out

    def id_two_sided_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Ic, Ir, T, S = self.PV(), self.PV(), self.PM(), self.PM()
        frank = self.I(0)
        self.lib.id_two_sided_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(Ic), C.byref(Ir),
                                                                  C.byref(T), C.byref(S))
        self.lib.matrix_delete(M)
        out = int(frank.value), self.from_vec(Ic), self.from_vec(Ir), self.from_mat(T), self.from_mat(S)
        self.check()
        return out

    def cur_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        Cm, U, R = self.PM(), self.PM(), self.PM()
        frank = self.I(0)
        self.lib.cur_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(Cm), C.byref(U), C.byref(R))
        self.lib.matrix_delete(M)
        out = int(frank.value), self.from_mat(Cm), self.from_mat(U), self.from_mat(R)
        self.check()
        return out

    def svd_from_qb(self, Q, B):
        Qm, Bm = self.to_mat(Q), self.to_mat(B)
        U, S, V = self.PM(), self.PM(), self.PM()
        self.lib.low_rank_svd_rand_decomp_fromQB(Qm, Bm, C.byref(U), C.byref(S), C.byref(V))
        self.lib.matrix_delete(Qm); self.lib.matrix_delete(Bm)
        out = self.from_mat(U), self.from_mat(S), self.from_mat(V)
        self.check()
        return out

    def id_from_qb(self, Q, B):
        Qm, Bm = self.to_mat(Q), self.to_mat(B)
        I_, T = self.PV(), self.PM()
        self.lib.id_rand_decomp_fromQB(Qm, Bm, C.byref(I_), C.byref(T))
        self.lib.matrix_delete(Qm); self.lib.matrix_delete(Bm)
        out = self.from_vec(I_), self.from_mat(T)
        self.check()
        return out

    def omega(self, nrows, ncols, seed=777):
        self.set_seed(seed)
        M = self.lib.matrix_new(nrows, ncols)
        self.lib.initialize_random_matrix(M)
        out = self.from_mat(M)
        self.check()
        return out
