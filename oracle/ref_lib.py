"""oracle/ref_lib.py — TEST INFRASTRUCTURE ONLY (imported by tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke(); never by the product package).

ctypes binding of oracle/_ref/libref{32,64}.so = the UNMODIFIED reference C sources
(/root/reference/multi_core_mkl_code{,_64bit}/*.c) compiled by oracle/build_ref.sh against an MKL
shim + OpenBLAS 0.3.15, with Omega supplied by the shared Philox generator (oracle/shim/vsl_shim.c).
Signatures follow rank_revealing_algorithms_intel_mkl.h:5-8,60,68,78,88 and
matrix_vector_functions_intel_mkl.h:18-27.
"""
import ctypes as C
import os
import numpy as np

from lowrankmatrixdecompositioncodes_b200 import _legacy_calls  # call wrappers only (no device code is touched)

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, "_ref")


def available(bits=32):
    return os.path.exists(os.path.join(_REFDIR, "libref%d.so" % bits))


class RefLib(_legacy_calls.LegacyCalls):
    """One ABI flavour (32: int indices, 64: int64_t indices) of the reference library."""

    def __init__(self, bits=32):
        self.bits = bits
        self.I = C.c_int if bits == 32 else C.c_int64
        I = self.I

        class Mat(C.Structure):
            _fields_ = [("nrows", I), ("ncols", I), ("d", C.POINTER(C.c_double))]

        class Vec(C.Structure):
            _fields_ = [("nrows", I), ("d", C.POINTER(C.c_double))]

        self.Mat, self.Vec = Mat, Vec
        self.lib = C.CDLL(os.path.join(_REFDIR, "libref%d.so" % bits), mode=C.RTLD_LOCAL)
        L = self.lib
        PM, PV = C.POINTER(Mat), C.POINTER(Vec)
        PPM, PPV = C.POINTER(PM), C.POINTER(PV)
        L.matrix_new.restype = PM
        L.matrix_new.argtypes = [I, I]
        L.vector_new.restype = PV
        L.vector_new.argtypes = [I]
        L.matrix_delete.argtypes = [PM]
        L.vector_delete.argtypes = [PV]
        L.oracle_set_seed.argtypes = [C.c_ulonglong]
        L.low_rank_svd_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.low_rank_svd_blockrand_decomp_fixed_rank_or_prec.argtypes = [
            PM, I, I, C.c_double, I, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.randQB_pb_new.argtypes = [PM, I, I, C.c_double, I, I, C.POINTER(I), PPM, PPM]
        L.id_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPV, PPM]
        L.id_two_sided_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPV, PPV, PPM, PPM]
        L.cur_rand_decomp_fixed_rank.argtypes = [PM, I, I, I, I, PPM, PPM, PPM]
        L.id_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPV, PPM]
        L.id_two_sided_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPV, PPV, PPM, PPM]
        L.cur_blockrand_decomp_fixed_rank_or_prec.argtypes = [PM, I, I, C.c_double, I, I, I, C.POINTER(I), PPM, PPM, PPM]
        L.matrix_load_from_binary_file.restype = PM
        L.matrix_load_from_binary_file.argtypes = [C.c_char_p]
        L.matrix_write_to_binary_file.argtypes = [PM, C.c_char_p]
        L.get_matrix_frobenius_norm.restype = C.c_double
        L.get_matrix_frobenius_norm.argtypes = [PM]
        L.get_percent_error_between_two_mats.restype = C.c_double
        L.get_percent_error_between_two_mats.argtypes = [PM, PM]
        L.form_svd_product_matrix.argtypes = [PM, PM, PM, PM]
        L.pivotedQR_mkl.argtypes = [PM, PPM, PPM, PPV]
        L.QR_factorization_getQ.argtypes = [PM, PM]
        L.initialize_random_matrix.argtypes = [PM]
        _legacy_calls.bind(L, I, PM, PV)

    # ---- marshalling -------------------------------------------------------------------------
    def set_seed(self, seed):
        self.lib.oracle_set_seed(int(seed))

    def to_mat(self, a):
        """numpy (m,n) -> reference mat (column-major copy owned by the reference allocator)."""
        a = np.asarray(a, dtype=np.float64)
        m, n = a.shape
        M = self.lib.matrix_new(m, n)
        buf = np.ctypeslib.as_array(M.contents.d, shape=(m * n,))
        buf[:] = np.asfortranarray(a).ravel(order="F")
        return M

    def from_mat(self, M, free=True):
        m, n = int(M.contents.nrows), int(M.contents.ncols)
        out = np.ctypeslib.as_array(M.contents.d, shape=(m * n,)).copy().reshape((m, n), order="F")
        if free:
            self.lib.matrix_delete(M)
        return out

    def from_vec(self, v, free=True):
        n = int(v.contents.nrows)
        out = np.ctypeslib.as_array(v.contents.d, shape=(n,)).copy()
        if free:
            self.lib.vector_delete(v)
        return out

    # ---- the hot-path API (reference semantics, numpy in/out) ----------------------------------
    def svd_rand(self, A, k, p, vnum=1, q=2, s=1, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PM = C.POINTER(self.Mat)
        U, S, V = PM(), PM(), PM()
        frank = self.I(0)
        self.lib.low_rank_svd_rand_decomp_fixed_rank(M, k, p, vnum, q, s, C.byref(frank),
                                                     C.byref(U), C.byref(S), C.byref(V))
        self.lib.matrix_delete(M)
        return self.from_mat(U), self.from_mat(S), self.from_mat(V)

    def svd_blockrand(self, A, k, p, TOL, vnum, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PM = C.POINTER(self.Mat)
        U, S, V = PM(), PM(), PM()
        frank = self.I(0)
        self.lib.low_rank_svd_blockrand_decomp_fixed_rank_or_prec(
            M, k, p, float(TOL), vnum, kstep, q, s, C.byref(frank), C.byref(U), C.byref(S), C.byref(V))
        self.lib.matrix_delete(M)
        return int(frank.value), self.from_mat(U), self.from_mat(S), self.from_mat(V)

    def randQB_pb_new(self, A, kstep, nstep, TOL, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PM = C.POINTER(self.Mat)
        Q, B = PM(), PM()
        frank = self.I(0)
        self.lib.randQB_pb_new(M, kstep, nstep, float(TOL), q, s, C.byref(frank), C.byref(Q), C.byref(B))
        self.lib.matrix_delete(M)
        return int(frank.value), self.from_mat(Q), self.from_mat(B)

    def id_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        I_, T = C.POINTER(self.Vec)(), C.POINTER(self.Mat)()
        self.lib.id_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(I_), C.byref(T))
        self.lib.matrix_delete(M)
        return self.from_vec(I_), self.from_mat(T)

    def id_two_sided_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PV, PM = C.POINTER(self.Vec), C.POINTER(self.Mat)
        Ic, Ir, T, S = PV(), PV(), PM(), PM()
        self.lib.id_two_sided_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(Ic), C.byref(Ir),
                                                     C.byref(T), C.byref(S))
        self.lib.matrix_delete(M)
        return self.from_vec(Ic), self.from_vec(Ir), self.from_mat(T), self.from_mat(S)

    def cur_rand(self, A, k, p, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PM = C.POINTER(self.Mat)
        Cm, U, R = PM(), PM(), PM()
        self.lib.cur_rand_decomp_fixed_rank(M, k, p, q, s, C.byref(Cm), C.byref(U), C.byref(R))
        self.lib.matrix_delete(M)
        return self.from_mat(Cm), self.from_mat(U), self.from_mat(R)


    def id_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        I_, T = C.POINTER(self.Vec)(), C.POINTER(self.Mat)()
        frank = self.I(0)
        self.lib.id_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(I_), C.byref(T))
        self.lib.matrix_delete(M)
        return int(frank.value), self.from_vec(I_), self.from_mat(T)

    def id_two_sided_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PV, PM = C.POINTER(self.Vec), C.POINTER(self.Mat)
        Ic, Ir, T, S = PV(), PV(), PM(), PM()
        frank = self.I(0)
        self.lib.id_two_sided_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(Ic), C.byref(Ir),
                                                                  C.byref(T), C.byref(S))
        self.lib.matrix_delete(M)
        return int(frank.value), self.from_vec(Ic), self.from_vec(Ir), self.from_mat(T), self.from_mat(S)

    def cur_blockrand(self, A, k, p, TOL, kstep, q, s, seed=777):
        self.set_seed(seed)
        M = self.to_mat(A)
        PM = C.POINTER(self.Mat)
        Cm, U, R = PM(), PM(), PM()
        frank = self.I(0)
        self.lib.cur_blockrand_decomp_fixed_rank_or_prec(M, k, p, float(TOL), kstep, q, s, C.byref(frank), C.byref(Cm), C.byref(U), C.byref(R))
        self.lib.matrix_delete(M)
        return int(frank.value), self.from_mat(Cm), self.from_mat(U), self.from_mat(R)

    def omega(self, nrows, ncols, seed=777):
        """The Omega the reference would draw for an nrows x ncols initialize_random_matrix."""
        self.set_seed(seed)
        M = self.lib.matrix_new(nrows, ncols)
        self.lib.initialize_random_matrix(M)
        return self.from_mat(M)


_RNG = None


def rng_lib():
    global _RNG
    if _RNG is None:
        _RNG = C.CDLL(os.path.join(_REFDIR, "liboracle_rng.so"), mode=C.RTLD_LOCAL)
        _RNG.oracle_fill_normal.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_longlong,
                                            C.POINTER(C.c_double)]
        _RNG.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint), C.c_uint, C.c_uint]
    return _RNG


def normal_stream(seed, first, count):
    """float32-valued normals (as float64) of linear entries first..first+count-1."""
    out = np.empty(int(count), dtype=np.float64)
    rng_lib().oracle_fill_normal(int(seed), int(first), int(count),
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def philox(ctr, key):
    c = (C.c_uint * 4)(*ctr)
    rng_lib().oracle_philox4x32_10(c, key[0], key[1])
    return [int(x) for x in c]
