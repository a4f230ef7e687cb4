"""CPU parity sweep of the plain-C helper layer (matrix_vector_functions_intel_mkl.h:32-360): every helper that does not need
the device is called in this build and in the compiled reference (oracle/_ref) on the same seeded inputs, both index ABIs, and
every matrix/vector argument is compared afterwards (in-place helpers modify their inputs).  No GPU involved."""
import ctypes as C

import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from oracle import ref_lib


class Side:
    """Uniform view of one library (ours or the reference's)."""

    def __init__(self, lib_obj):
        self.o, self.L, self.I = lib_obj, lib_obj.lib, lib_obj.I
        self.PM, self.PV = C.POINTER(lib_obj.Mat), C.POINTER(lib_obj.Vec)
        self.L.vector_new.restype = self.PV
        self.L.vector_new.argtypes = [self.I]

    def mat(self, a):
        return self.o.to_mat(a)

    def vec(self, v):
        p = self.L.vector_new(len(v))
        np.ctypeslib.as_array(p.contents.d, shape=(max(len(v), 1),))[:len(v)] = v
        return p

    def get(self, x):
        if isinstance(x, self.PM):
            return self.o.from_mat(x, free=False)
        return self.o.from_vec(x, free=False)


def sides(bits):
    if not ref_lib.available(bits):
        pytest.skip("compiled reference not present")
    return Side(pkg.Api(bits)), Side(ref_lib.RefLib(bits))


rng = np.random.default_rng(12345)
A = rng.standard_normal((7, 5))
B = rng.standard_normal((7, 5))
Sq = rng.standard_normal((6, 6))
v7, w7, v5 = rng.standard_normal(7), rng.standard_normal(7), rng.standard_normal(5)
perm5 = rng.permutation(5).astype(np.float64)
perm7 = rng.permutation(7).astype(np.float64)

# name, argument builders (m: matrix, v: vector, i: index, d: double, z(r,c): zero matrix, zv(n): zero vector, ia: index array), restype
CASES = [
    ("vector_scale", [("v", v7), ("d", 2.5)], None),
    ("matrix_scale", [("m", A), ("d", -0.5)], None),
    ("vector_get2norm", [("v", v7)], C.c_double),
    ("vector_copy", [("zv", 7), ("v", v7)], None),
    ("matrix_copy", [("z", (7, 5)), ("m", A)], None),
    ("matrix_hard_threshold", [("m", A), ("d", 0.7)], None),
    ("matrix_build_transpose", [("z", (5, 7)), ("m", A)], None),
    ("vector_sub", [("v", v7), ("v", w7)], None),
    ("matrix_sub", [("m", A), ("m", B)], None),
    ("matrix_sub_column_times_row_vector", [("m", A), ("v", v7), ("v", v5)], None),
    ("get_matrix_frobenius_norm", [("m", A)], C.c_double),
    ("get_matrix_max_abs_element", [("m", A)], C.c_double),
    ("vector_dot_product", [("v", v7), ("v", w7)], C.c_double),
    ("get_matrix_column_norm_squared", [("m", A), ("i", 3)], C.c_double),
    ("compute_matrix_column_norms", [("m", A), ("zv", 5)], None),
    ("get_percent_error_between_two_mats", [("m", A), ("m", B)], C.c_double),
    ("matrix_get_col", [("m", A), ("i", 2), ("zv", 7)], None),
    ("matrix_set_col", [("m", A), ("i", 4), ("v", v7)], None),
    ("matrix_get_row", [("m", A), ("i", 6), ("zv", 5)], None),
    ("matrix_set_row", [("m", A), ("i", 0), ("v", v5)], None),
    ("matrix_get_selected_columns", [("m", A), ("ia", [4, 0, 2]), ("z", (7, 3))], None),
    ("matrix_set_selected_columns", [("m", A), ("ia", [1, 3]), ("m", B[:, :2])], None),
    ("matrix_get_selected_rows", [("m", A), ("ia", [6, 1, 3, 0]), ("z", (4, 5))], None),
    ("matrix_set_selected_rows", [("m", A), ("ia", [5, 2]), ("m", B[:2, :])], None),
    ("matrix_copy_symmetric", [("z", (6, 6)), ("m", Sq)], None),
    ("matrix_keep_only_upper_triangular", [("m", Sq)], None),
    ("initialize_diagonal_matrix", [("z", (5, 5)), ("v", v5)], None),
    ("initialize_identity_matrix", [("m", Sq)], None),
    ("invert_diagonal_matrix", [("z", (5, 5)), ("m", np.diag(v5))], None),
    ("fill_vector_from_row_list", [("v", v5), ("v", perm5), ("zv", 5)], None),
    ("matrix_copy_first_rows", [("z", (3, 5)), ("m", A)], None),
    ("matrix_copy_first_columns", [("z", (7, 2)), ("m", A)], None),
    ("matrix_copy_first_columns_with_param", [("z", (7, 5)), ("m", A), ("i", 3)], None),
    ("matrix_copy_first_k_rows_and_columns", [("z", (4, 4)), ("m", Sq)], None),
    ("matrix_copy_all_rows_and_last_columns_from_indexk", [("z", (7, 3)), ("m", A), ("i", 2)], None),
    ("fill_matrix_from_first_rows", [("m", A), ("i", 4), ("z", (4, 5))], None),
    ("fill_matrix_from_last_rows", [("m", Sq), ("i", 2), ("z", (2, 6))], None),
    ("fill_matrix_from_first_columns", [("m", A), ("i", 3), ("z", (7, 3))], None),
    ("fill_matrix_from_last_columns", [("m", A), ("i", 2), ("z", (7, 2))], None),
    ("fill_matrix_from_last_columns_from_specified_one", [("m", A), ("i", 2), ("z", (7, 3))], None),
    ("fill_matrix_from_lower_right_corner", [("m", Sq), ("i", 3), ("z", (3, 3))], None),
    ("fill_matrix_from_first_columns_from_list", [("m", A), ("v", perm5), ("i", 3), ("z", (7, 3))], None),
    ("fill_matrix_from_first_rows_from_list", [("m", A), ("v", perm7), ("i", 4), ("z", (4, 5))], None),
    ("fill_matrix_from_last_columns_from_list", [("m", A), ("v", perm5), ("i", 2), ("z", (7, 3))], None),   # columns I[k:]
    ("append_matrices_horizontally", [("m", A), ("m", B[:, :3]), ("z", (7, 8))], None),
    ("append_matrices_vertically", [("m", A), ("m", B[:4, :]), ("z", (11, 5))], None),
    ("vector_build_rewrapped", [("zv", 7), ("v", perm7)], None),
    ("project_vector", [("v", v7), ("v", w7), ("zv", 7)], None),
    ("get_householder_matrix", [("v", v7), ("i", 2), ("i", 7), ("z", (7, 1))], None),
]


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("name,spec,restype", CASES, ids=[c[0] for c in CASES])
def test_host_helper_matches_reference(bits, name, spec, restype):
    ours, ref = sides(bits)
    results = []
    for side in (ours, ref):
        if not hasattr(side.L, name):
            pytest.fail("symbol %s missing" % name)
        fn = getattr(side.L, name)
        args, types, holders = [], [], []
        for kind, val in spec:
            if kind == "m":
                x = side.mat(np.array(val, dtype=np.float64)); types.append(side.PM); holders.append(x)
            elif kind == "z":
                x = side.mat(np.zeros(val)); types.append(side.PM); holders.append(x)
            elif kind == "v":
                x = side.vec(np.array(val, dtype=np.float64)); types.append(side.PV); holders.append(x)
            elif kind == "zv":
                x = side.vec(np.zeros(val)); types.append(side.PV); holders.append(x)
            elif kind == "i":
                x = side.I(val); types.append(side.I)
            elif kind == "d":
                x = C.c_double(val); types.append(C.c_double)
            elif kind == "ia":
                x = (side.I * len(val))(*val); types.append(C.POINTER(side.I))
            args.append(x)
        fn.argtypes, fn.restype = types, restype
        ret = fn(*args)
        results.append((ret, [side.get(h) for h in holders]))
    (r0, out0), (r1, out1) = results
    if restype is not None:
        assert r0 == pytest.approx(r1, rel=1e-13, abs=1e-300), name
    for a, b in zip(out0, out1):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-13, atol=1e-15), name


def test_maxcolnorm():
    """matrix_getmaxcolnorm (MVF:422-442): the reference shares one scratch vector between its OpenMP threads (a data race),
    so the comparison is against numpy."""
    ours = Side(pkg.Api(32))
    ours.L.matrix_getmaxcolnorm.restype = C.c_double
    ours.L.matrix_getmaxcolnorm.argtypes = [ours.PM]
    assert ours.L.matrix_getmaxcolnorm(ours.mat(A)) == pytest.approx(np.linalg.norm(A, axis=0).max(), rel=1e-14)


@pytest.mark.parametrize("bits", [32, 64])
def test_min_max_element_match_reference(bits):
    ours, ref = sides(bits)
    out = []
    for side in (ours, ref):
        for name in ("vector_get_min_element", "vector_get_max_element"):
            fn = getattr(side.L, name)
            fn.argtypes = [side.PV, C.POINTER(side.I), C.POINTER(C.c_double)]
            i, d = side.I(-1), C.c_double(0)
            fn(side.vec(perm7 - 3.0), C.byref(i), C.byref(d))     # integer-valued: the reference truncates through an int (MVF:315-319)
            out.append((int(i.value), d.value))
    assert out[:2] == out[2:]
