"""ctypes signatures of the device layer C-ABI (include/rsvd_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsvd_b200.so")
API32_PATH = os.path.join(_HERE, "librsvd_b200_api32.so")
API64_PATH = os.path.join(_HERE, "librsvd_b200_api64.so")


class NativeLibraryMissing(RuntimeError):
    pass


def _load(path):
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            "%s is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no Python/CPU fallback." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL if path == LIB_PATH else C.RTLD_LOCAL)


i64 = C.c_longlong
dp = C.c_void_p   # device pointers travel as integers
u64 = C.c_uint64

# every symbol include/rsvd_b200.h declares, with its ctypes signature (restype, argtypes)
SIGNATURES = {
    "rsvd_b200_init": (C.c_int, [C.c_int]),
    "rsvd_b200_device_count": (C.c_int, []),
    "rsvd_b200_status": (C.c_int, []),
    "rsvd_b200_last_error": (C.c_char_p, []),
    "rsvd_b200_clear_error": (None, []),
    "rsvd_b200_stream": (C.c_void_p, []),
    "rsvd_b200_sync": (None, []),
    "rsvd_b200_launch_count": (C.c_ulonglong, []),
    "rsvd_b200_set_option": (None, [C.c_char_p, i64]),
    "rsvd_b200_get_option": (i64, [C.c_char_p]),
    "rsvd_b200_dev_alloc": (dp, [i64]),
    "rsvd_b200_dev_free": (None, [dp]),
    "rsvd_b200_h2d": (C.c_int, [dp, C.c_void_p, i64]),
    "rsvd_b200_d2h": (C.c_int, [C.c_void_p, dp, i64]),
    "rsvd_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "rsvd_b200_host_free": (None, [C.c_void_p]),
    "rsvd_b200_gemm": (C.c_int, [C.c_char, C.c_char, i64, i64, i64, C.c_double, dp, i64, dp, i64, C.c_double, dp, i64]),
    "rsvd_b200_sketch": (C.c_int, [C.c_char, i64, i64, i64, dp, i64, u64, i64, i64, i64, dp, i64]),
    "rsvd_b200_fill_normal": (C.c_int, [dp, i64, u64, i64]),
    "rsvd_b200_orthonormalize": (C.c_int, [dp, i64, i64, i64, dp, i64]),
    "rsvd_b200_chol_inv": (C.c_int, [dp, i64, i64, dp, i64, dp]),
    "rsvd_b200_geqp3": (C.c_int, [dp, i64, i64, i64, dp]),
    "rsvd_b200_geqp3_q": (C.c_int, [dp, i64, i64, i64, dp, dp, i64]),
    "rsvd_b200_svd_small": (C.c_int, [dp, i64, i64, dp, i64, dp, dp, i64]),
    "rsvd_b200_eig_small": (C.c_int, [dp, i64, i64, dp]),
    "rsvd_b200_trsm_left_upper": (C.c_int, [dp, i64, i64, dp, i64, i64]),
    "rsvd_b200_lu_solve": (C.c_int, [dp, i64, i64, dp, i64, i64]),
    "rsvd_b200_frob_norm": (C.c_double, [dp, i64, i64, i64]),
    "rsvd_b200_transpose": (C.c_int, [dp, i64, dp, i64, i64, i64]),
    "rsvd_b200_svd_rand_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, u64, dp, dp, i64, dp, dp, i64]),
    "rsvd_b200_svd_rand_host": (C.c_int, [C.c_void_p, dp, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, u64, dp, i64, dp, dp, i64]),
    "rsvd_b200_randqb_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_double, C.c_int, C.c_int, u64, dp, i64, dp, i64, C.POINTER(i64)]),
    "rsvd_b200_svd_from_q_dev": (C.c_int, [dp, i64, i64, i64, dp, i64, i64, i64, C.c_int, dp, i64, dp, dp, i64]),
    "rsvd_b200_id_rand_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, dp, dp, dp, i64]),
    "rsvd_b200_id_full_dev": (C.c_int, [dp, i64, i64, i64, dp, dp, i64]),
    "rsvd_b200_id_qr_dev": (C.c_int, [dp, i64, i64, i64, i64, dp, dp, i64]),
    "rsvd_b200_id_rows_dev": (C.c_int, [dp, i64, i64, i64, dp, i64, dp, dp, i64]),
    "rsvd_b200_cur_from_id_dev": (C.c_int, [dp, i64, i64, i64, dp, dp, dp, i64, i64, dp, i64, dp, i64, dp, i64]),
    "rsvd_b200_svd_from_qb_dev": (C.c_int, [dp, i64, i64, dp, i64, i64, i64, dp, i64, dp, dp, i64]),
    "rsvd_b200_id_two_sided_rand_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, dp, dp, dp, i64, dp, i64]),
    "rsvd_b200_cur_rand_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, dp, i64, dp, i64, dp, i64]),
    "rsvd_b200_svd_rand_h": (C.c_int, [C.c_void_p, i64, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, u64, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]),
    "rsvd_b200_id_rand_h": (C.c_int, [C.c_void_p, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, C.c_void_p, C.c_void_p, i64]),
    "rsvd_b200_id_two_sided_rand_h": (C.c_int, [C.c_void_p, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, C.c_void_p, C.c_void_p, C.c_void_p, i64, C.c_void_p, i64]),
    "rsvd_b200_cur_rand_h": (C.c_int, [C.c_void_p, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64]),
    "rsvd_b200_randqb_h": (C.c_void_p, [C.c_void_p, i64, i64, i64, i64, i64, i64, C.c_double, C.c_int, C.c_int, u64, C.POINTER(i64)]),
    "rsvd_b200_qb_parts": (i64, [C.c_void_p]),
    "rsvd_b200_qb_download": (C.c_int, [C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64]),
    "rsvd_b200_qb_svd": (C.c_int, [C.c_void_p, i64, i64, C.c_int, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]),
    "rsvd_b200_qb_dev_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "rsvd_b200_qb_release_handle": (None, [C.c_void_p]),
    "rsvd_b200_qb_free": (None, [C.c_void_p]),
    "rsvd_b200_pin_matrix": (C.c_int, [C.c_void_p, i64, i64]),
    "rsvd_b200_unpin_matrix": (None, [C.c_void_p]),
    "rsvd_b200_is_resident": (C.c_int, [C.c_void_p]),
    "rsvd_b200_set_devices": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "rsvd_b200_active_devices": (C.c_int, []),
    "rsvd_b200_jacobi_schedule": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "rsvd_b200_load_binary_dev": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(i64)]),
    "rsvd_b200_store_binary_dev": (C.c_int, [C.c_char_p, C.c_int, dp, i64, i64, i64]),
    "rsvd_b200_randqb_legacy_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, C.c_int, C.c_int, u64, dp, i64, dp, i64]),
    "rsvd_b200_randqb_single_dev": (C.c_int, [dp, i64, i64, i64, i64, i64, u64, dp, i64, dp, i64]),
    "rsvd_b200_svd_from_qb_asc_dev": (C.c_int, [dp, i64, i64, dp, i64, i64, i64, dp, i64, dp, dp, i64]),
    "rsvd_b200_svd_full_dev": (C.c_int, [dp, i64, i64, i64, dp, i64, dp, dp, i64]),
    "rsvd_b200_estimate_rank1_dev": (C.c_int, [dp, i64, i64, i64, i64, C.c_double, u64, dp, i64, C.POINTER(i64)]),
    "rsvd_b200_estimate_rank2_dev": (C.c_int, [dp, i64, i64, i64, i64, C.c_double, u64, dp, i64, dp, i64, i64, C.POINTER(i64)]),
    "rsvd_b200_svd_rand_from_sketch_dev": (C.c_int, [dp, i64, i64, i64, dp, i64, i64, C.c_int, C.c_int, dp, i64, dp, dp, i64]),
    "rsvd_b200_pqr_partial_dev": (C.c_int, [dp, i64, i64, i64, i64, C.c_double, C.c_int, dp, dp, i64, dp, i64, C.POINTER(i64)]),
    "rsvd_b200_svd_percent_error_dev": (C.c_double, [dp, i64, i64, i64, dp, i64, dp, dp, i64, i64]),
    "rsvd_b200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "rsvd_b200_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "rsvd_b200_comm_destroy": (None, []),
    "rsvd_b200_allreduce_sum": (C.c_int, [dp, i64]),
    "rsvd_b200_row_partition": (None, [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]),
    "rsvd_b200_fp64_peak_tflops": (C.c_double, [C.c_int, C.c_int]),
}

_dev = None


def dev():
    """The device-layer library (loaded once, RTLD_GLOBAL so the API libraries resolve against it)."""
    global _dev
    if _dev is None:
        lib = _load(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _dev = lib
    return _dev


def check(rc=0):
    """Raise if the device layer recorded an error."""
    lib = dev()
    if rc or lib.rsvd_b200_status():
        msg = lib.rsvd_b200_last_error().decode()
        lib.rsvd_b200_clear_error()
        raise RuntimeError("rsvd_b200: " + (msg or "unknown error"))


def row_partition(m, world, rank):
    r0, rows = i64(0), i64(0)
    dev().rsvd_b200_row_partition(m, world, rank, C.byref(r0), C.byref(rows))
    return int(r0.value), int(rows.value)
