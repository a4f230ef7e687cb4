"""GPU tests of the file <-> device path: the reference's binary matrix format
(matrix_vector_functions_intel_mkl.c:77-133, 64-bit :78-135) streamed straight into / out of HBM, bit-exact against the
numpy restatement of the format (oracle.rsvd_numpy.write_matrix_binary / read_matrix_binary)."""
import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native, device as D
from oracle import rsvd_numpy as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    lib = native.dev()
    assert lib.rsvd_b200_init(0) == 0, lib.rsvd_b200_last_error().decode()
    return lib


@pytest.mark.parametrize("m,n,bits", [(1, 1, 32), (7, 3, 32), (1000, 333, 32), (333, 1000, 64), (5000, 4100, 32), (3, 70000, 64)])
def test_load_and_store_round_trip_bit_exact(lib, tmp_path, m, n, bits):
    A = np.random.default_rng(m * n).standard_normal((m, n))
    src = str(tmp_path / "a.bin")
    O.write_matrix_binary(A, src, bits)
    M = D.load_binary(src, bits)
    assert (M.m, M.n) == (m, n)
    t = M.to_torch()
    assert np.array_equal(D.to_numpy(t), A)
    dst = str(tmp_path / "b.bin")
    D.store_binary(dst, M.ptr, m, m, n, bits)
    M.free()
    assert open(dst, "rb").read() == open(src, "rb").read()


def test_truncated_and_missing_files_fail_loudly(lib, tmp_path):
    with pytest.raises(Exception):
        D.load_binary(str(tmp_path / "nope.bin"))
    lib.rsvd_b200_clear_error()
    A = np.ones((50, 40))
    p = str(tmp_path / "t.bin")
    O.write_matrix_binary(A, p, 32)
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-80])
    with pytest.raises(Exception):
        D.load_binary(p)
    lib.rsvd_b200_clear_error()


def test_file_to_factors_without_a_host_matrix(lib, tmp_path):
    """driver_multi_core_mkl1.c's job (load, rank-k SVD, percent error) with the matrix never materialised on the host."""
    A, _ = O.make_matrix(1200, 900, "logspace", seed=1)
    p = str(tmp_path / "a.bin")
    O.write_matrix_binary(A, p, 32)
    M = D.load_binary(p)
    t = M.to_torch()
    M.free()
    U, S, V = D.svd_rand(t, 60, 10, 1, 2, 1, seed=777)
    Ur, Sr, Vr = O.low_rank_svd_rand_decomp_fixed_rank(A, 60, 10, 1, 2, 1, 777)
    assert np.max(np.abs(S.cpu().numpy() - np.diag(Sr)) / np.diag(Sr)) < 1e-10
    api = pkg.Api(32)
    Ma = api.lib.matrix_load_from_binary_file(p.encode())           # host loader of the same file: identical matrix
    assert np.array_equal(api.from_mat(Ma), A)
