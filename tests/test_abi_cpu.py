"""CPU tests of the boundary: the three shared libraries load without a GPU, export every symbol the headers
declare, keep the reference's struct layout, fail loudly (no CPU fallback), and the host-only pieces (allocation,
slicing helpers, both binary file formats, row partition) behave like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lowrankmatrixdecompositioncodes_b200 as pkg
from lowrankmatrixdecompositioncodes_b200 import native
from oracle import rsvd_numpy as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", " ".join(
        l for l in txt.splitlines() if not l.strip().startswith("#")))) - {"defined", "sizeof"})


def test_device_layer_exports_every_declared_symbol():
    lib = native.dev()
    names = [n for n in _declared("rsvd_b200.h") if n.startswith("rsvd_b200_")]
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(native.SIGNATURES) == names   # the ctypes table covers the whole header


@pytest.mark.parametrize("bits", [32, 64])
def test_api_libraries_export_reference_symbols(bits):
    api = pkg.Api(bits)
    names = [n for n in _declared("rsvd_b200_rra_decl.h") + _declared("rsvd_b200_matvec_decl.h")
             if n not in ("min", "max", "x", "y")]
    assert "low_rank_svd_rand_decomp_fixed_rank" in names and "matrix_load_from_binary_file" in names
    for n in names:
        assert hasattr(api.lib, n), n


@pytest.mark.parametrize("bits", [32, 64])
def test_every_symbol_of_the_compiled_reference_is_exported(bits):
    """A driver written against the reference links against the drop-in whatever subset of the API it uses: the dynamic
    symbol table of the compiled reference (minus the oracle's own shim symbols) must be a subset of ours."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref", "libref%d.so" % bits)
    if not os.path.exists(ref):
        pytest.skip("compiled reference not present")
    def syms(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        return {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TW"}
    ours = syms(native.API32_PATH if bits == 32 else native.API64_PATH)
    theirs = {n for n in syms(ref) if not n.startswith(("oracle_", "vsl", "vsRng", "_"))}
    assert len(theirs) > 100
    assert theirs <= ours, sorted(theirs - ours)


def test_struct_layout_matches_reference():
    a32, a64 = pkg.Api(32), pkg.Api(64)
    assert C.sizeof(a32.Mat) == 16 and C.sizeof(a32.Vec) == 16    # {int,int,double*}, {int,double*}  (MVH:18-27)
    assert C.sizeof(a64.Mat) == 24 and C.sizeof(a64.Vec) == 16    # int64_t twins
    assert a32.Mat.d.offset == 8 and a64.Mat.d.offset == 16


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    api = pkg.Api(32)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        api.svd_rand(np.random.rand(30, 20), 4, 2)
    with pytest.raises(RuntimeError):
        api.id_rand(np.random.rand(30, 20), 4, 2, 1, 1)
    # the baselines and legacy entry points have no host arithmetic behind them either
    for call in (lambda: api.pqr(np.random.rand(30, 20), 5), lambda: api.svd_decomp(np.random.rand(30, 20), 5, 0.0),
                 lambda: api.randQB_p(np.random.rand(30, 20), 3, 1), lambda: api.randQB_pb(np.random.rand(30, 20), 2, 2, 1, 1),
                 lambda: api.svd3(np.random.rand(30, 20), 4, 2, 1), lambda: api.estimate_rank2(np.random.rand(30, 20), 4, 0.1)):
        call()
        with pytest.raises(RuntimeError):
            api.check()
    from lowrankmatrixdecompositioncodes_b200 import device as D
    with pytest.raises(RuntimeError):
        D.load_binary(__file__)
    native.dev().rsvd_b200_clear_error()


def test_multi_device_request_without_gpus_fails_loudly():
    """RSVD_B200_DEVICES on a box without GPUs: the worker pool is not created, the error is recorded, hot-path calls still fail
    loudly (no CPU fallback), the plain-C helpers keep working."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = ("import numpy as np, lowrankmatrixdecompositioncodes_b200 as pkg\n"
            "lib = pkg.native.dev()\n"
            "assert lib.rsvd_b200_active_devices() == 1\n"
            "assert lib.rsvd_b200_status() != 0 and b'CUDA device' in lib.rsvd_b200_last_error()\n"
            "api = pkg.Api(32)\n"
            "M = api.to_mat(np.arange(12.0).reshape(3, 4))\n"
            "assert api.lib.get_matrix_frobenius_norm(M) == float(np.linalg.norm(np.arange(12.0)))\n"
            "try:\n"
            "    api.svd_rand(np.random.rand(30, 20), 4, 2)\n"
            "    raise SystemExit('no error raised')\n"
            "except RuntimeError as e:\n"
            "    assert 'CUDA device' in str(e) or 'fallback' in str(e), e\n"
            "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, RSVD_B200_DEVICES="0-3"), timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("bits", [32, 64])
def test_binary_io_both_formats(bits, tmp_path):
    api = pkg.Api(bits)
    A = np.random.default_rng(0).standard_normal((37, 91))
    f = str(tmp_path / "a.bin")
    O.write_matrix_binary(A, f, bits)                       # the reference's format (make_matrix_binary.m:32-41)
    M = api.lib.matrix_load_from_binary_file(f.encode())
    assert np.array_equal(api.from_mat(M), A)
    g = str(tmp_path / "b.bin")
    M = api.to_mat(A)
    api.lib.matrix_write_to_binary_file(M, g.encode())
    api.lib.matrix_delete(M)
    assert open(f, "rb").read() == open(g, "rb").read()
    assert os.path.getsize(g) == (8 if bits == 32 else 16) + 37 * 91 * 8
    # ragged / empty
    E = np.zeros((0, 5))
    O.write_matrix_binary(E, f, bits)
    M = api.lib.matrix_load_from_binary_file(f.encode())
    assert api.from_mat(M).shape == (0, 5)


def test_io_file_interchange_with_reference(ref32, tmp_path):
    api = pkg.Api(32)
    A = np.random.default_rng(1).standard_normal((13, 7))
    f = str(tmp_path / "x.bin")
    M = api.to_mat(A)
    api.lib.matrix_write_to_binary_file(M, f.encode())
    api.lib.matrix_delete(M)
    R = ref32.lib.matrix_load_from_binary_file(f.encode())
    assert np.array_equal(ref32.from_mat(R), A)


def test_host_slicing_helpers_match_reference(ref32):
    api = pkg.Api(32)
    A = np.random.default_rng(2).standard_normal((9, 6))
    for name in ["resize_matrix_by_columns", "resize_matrix_by_columns_from_end", "resize_matrix_by_rows",
                 "resize_matrix_by_rows_from_end"]:
        for lib_ in (api, ref32):
            getattr(lib_.lib, name).argtypes = [C.POINTER(C.POINTER(lib_.Mat)), lib_.I]
        M1, M2 = api.to_mat(A), ref32.to_mat(A)
        getattr(api.lib, name)(C.byref(M1), 4)
        getattr(ref32.lib, name)(C.byref(M2), 4)
        assert np.array_equal(api.from_mat(M1), ref32.from_mat(M2)), name
    # vector_build_rewrapped + fill_matrix_from_first_rows_from_list (CUR tail helpers, MVF:1195-1201,1029-1042)
    perm = np.random.default_rng(3).permutation(9).astype(np.float64)
    outs = []
    for lib_ in (api, ref32):
        lib_.lib.vector_build_rewrapped.argtypes = [C.POINTER(lib_.Vec), C.POINTER(lib_.Vec)]
        lib_.lib.fill_matrix_from_first_rows_from_list.argtypes = [C.POINTER(lib_.Mat), C.POINTER(lib_.Vec), lib_.I, C.POINTER(lib_.Mat)]
        v = lib_.lib.vector_new(9)
        np.ctypeslib.as_array(v.contents.d, shape=(9,))[:] = perm
        vi = lib_.lib.vector_new(9)
        lib_.lib.vector_build_rewrapped(vi, v)
        M, Mk = lib_.to_mat(A), lib_.lib.matrix_new(5, 6)
        lib_.lib.fill_matrix_from_first_rows_from_list(M, v, 5, Mk)
        outs.append((lib_.from_vec(vi), lib_.from_mat(Mk)))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert api.lib.get_matrix_frobenius_norm(api.to_mat(A)) == pytest.approx(np.linalg.norm(A), rel=1e-14)


def test_row_partition_covers_all_rows():
    for m, w in [(50000, 8), (1000000, 8), (17, 4), (400000, 3), (5, 8)]:
        tot, prev_end = 0, 0
        for r in range(w):
            r0, rows = native.row_partition(m, w, r)
            assert r0 == min(prev_end, m) and rows >= 0
            assert r0 % 16 == 0 or rows == 0
            prev_end = r0 + rows
            tot += rows
        assert tot == m


def test_reference_drivers_relink_unchanged():
    """drivers `#include "rank_revealing_algorithms_intel_mkl.h"` and relink.  oracle/build_ref.sh compiles
    the reference's unmodified driver sources against include/ + the B200 libraries; the binaries must exist and resolve
    every symbol (ldd) on a GPU-less host."""
    import subprocess
    d = os.path.join(ROOT, "oracle", "_ref", "relink")
    if not os.path.isdir(d):
        pytest.skip("relinked drivers not built (reference tree absent)")
    for name in ["driver_multi_core_mkl1", "driver_multi_core_mkl2", "driver_multi_core_mkl3", "driver_multi_core_mkl4",
                 "driver_multi_core_mkl5", "driver1_64bit"]:
        exe = os.path.join(d, name)
        assert os.path.exists(exe), name
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "not found" not in out, out
        assert "librsvd_b200" in out


@pytest.mark.parametrize("n,bw", [(2, 1), (7, 1), (520, 1), (3, 2), (10, 2), (520, 2), (1050, 2), (9, 4), (520, 4), (64, 4)])
def test_jacobi_pair_schedule_is_a_complete_tournament(n, bw):
    """Host logic of the Jacobi kernels (csrc/device/jacobi.cu: block_pair / intra_pair / pair_at, shared by the rotation
    kernel and the replay kernel): in one sweep every pair of columns meets exactly once and the pairs of a step are disjoint."""
    lib = native.dev()
    N = lib.rsvd_b200_jacobi_schedule(n, bw, None)
    assert N >= n and N % (2 * bw) == 0 and N - n < 2 * bw
    buf = (C.c_int * ((N - 1) * (N // 2) * 2))()
    assert lib.rsvd_b200_jacobi_schedule(n, bw, buf) == N
    pairs = np.frombuffer(buf, dtype=np.int32).reshape(N - 1, N // 2, 2)
    assert pairs.min() == 0 and pairs.max() == N - 1
    for st in range(N - 1):
        assert len(set(pairs[st].ravel().tolist())) == N          # disjoint: all N columns appear once
    met = {(min(p, q), max(p, q)) for p, q in pairs.reshape(-1, 2).tolist()}
    assert len(met) == N * (N - 1) // 2 and all(p != q for p, q in met)


def test_binary_loader_multi_block_and_truncated(tmp_path):
    """The loader double-buffers 32 MB row blocks (>= 8 rows each): a wide matrix exercises several blocks and a ragged last
    one; a truncated file is reported through the status channel and still returns a matrix (rows read so far)."""
    api = pkg.Api(32)
    m, n = 21, 524288 + 3                                     # 4 MB rows -> 8-row blocks: 8 + 8 + 5
    A = np.random.default_rng(5).standard_normal((m, n))
    f = str(tmp_path / "wide.bin")
    O.write_matrix_binary(A, f, 32)
    M = api.lib.matrix_load_from_binary_file(f.encode())
    assert api.lib.rsvd_b200_api_status() == 0
    assert np.array_equal(api.from_mat(M), A)
    g = str(tmp_path / "wide_out.bin")
    M = api.to_mat(A)
    api.lib.matrix_write_to_binary_file(M, g.encode())
    api.lib.matrix_delete(M)
    assert open(f, "rb").read() == open(g, "rb").read()
    data = open(f, "rb").read()
    open(f, "wb").write(data[: 8 + 12 * n * 8 + 40])         # cut in the middle of the second block
    api.lib.rsvd_b200_api_clear_error()
    M = api.lib.matrix_load_from_binary_file(f.encode())
    assert api.lib.rsvd_b200_api_status() != 0 and b"truncated" in api.lib.rsvd_b200_api_last_error()
    B = api.from_mat(M)
    assert B.shape == (m, n) and np.array_equal(B[:8], A[:8])
    api.lib.rsvd_b200_api_clear_error()


def test_every_runtime_option_is_documented_in_the_header():
    """rsvd_b200_set_option / get_option accept exactly the names runtime.cu compares against; each must appear in the option
    comment of include/rsvd_b200.h (an undocumented switch is an interface nobody can rely on)."""
    import re
    src = open(os.path.join(ROOT, "lowrankmatrixdecompositioncodes_b200", "csrc", "device", "runtime.cu")).read()
    body = src[src.index("void rsvd_b200_set_option("):]
    body = body[:body.index("\n}\n", body.index("rsvd_i64 rsvd_b200_get_option("))]
    names = set(re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', body))
    assert {"seed", "verbose", "no_chol_dataflow", "no_live_replay", "no_block_cache", "last_qr_path"} <= names
    header = open(os.path.join(ROOT, "include", "rsvd_b200.h")).read()
    missing = sorted(n for n in names if '"%s"' % n not in header)
    assert not missing, "options missing from include/rsvd_b200.h: %s" % missing
