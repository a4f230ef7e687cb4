#!/usr/bin/env python
"""bench.py — randomized SVD (low_rank_svd_rand_decomp_fixed_rank) time-to-solution and FP64 TFLOP/s on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own CPU path, oracle/_ref, on host cores)

Workload (BASELINE.json configs[1]): 50000 x 20000 FP64, k=500, p=20, q=2, s=1, vnum=1 per GPU.  With N GPUs the
matrix is row-partitioned, 50000 rows per rank (weak scaling: global matrix 50000*N x 20000), NCCL only for the
n x l and l x l sums.  A "step" is one complete decomposition (2q = 4 streaming passes over A, 3 CholeskyQR2
orthonormalisations, the QR of B^T, the l x l Jacobi SVD and the products forming U and V).

  value  = GEMM FLOPs of the job (2q * 2*m*n*(k+p)) / device time, inputs resident in HBM
  e2e    = the same metric through the reference's C API (host `mat` in pinned memory -> U,S,V in host memory),
           host<->device copies inside the timed region
  roofline = the dominant kernel (gemm_tma_kernel, one streaming pass) timed alone with CUDA events on the
           library's stream vs the FP64 peak measured in-run (MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline = the reference C code (oracle/_ref: unmodified sources on OpenBLAS 0.3.15; MKL unavailable offline)
           on the box's host cores, ONE run of the full workload on the very matrix the e2e leg used (about 20 s), and
           `parity` = its U, S, V against ours for the same Omega (sigma rel. error, principal angles, percent errors)
  --impl reference = the same reference code on the full workload for every warm-up and timed step (same config)
  north_star (N = 8 only) = BASELINE configs[4] and configs[3] device-resident, as sub-records of the same JSON line
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_PER_GPU, N_COLS, K, P, Q, S_ORTH, VNUM = 50000, 20000, 500, 20, 2, 1, 1
WORKLOAD = "low_rank_svd_rand_decomp_fixed_rank 50000x20000 fp64 (per GPU), k=500 p=20 q=2 s=1 vnum=1 (BASELINE configs[1])"
# --config c5: BASELINE configs[4] (the north-star target): 1,000,000 x 100,000 over 8 GPUs = 125,000 rows (100 GB) per GPU
C5 = dict(rows=125000, n=100000, k=1000, p=50,
          workload="low_rank_svd_rand_decomp_fixed_rank 1,000,000x100,000 fp64 over 8 GPUs (125000 rows = 100 GB per GPU), k=1000 p=50 q=2 (BASELINE configs[4])")
METRIC = "randSVD FP64 TFLOP/s (GEMM flops 2q*2mn(k+p) / time-to-solution)"


def gemm_flops(m, n, l, q):
    return 2.0 * q * 2.0 * m * n * l


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle/_ref on host cores
# ------------------------------------------------------------------------------------------------------------------
def bench_config(world, rows):
    """The `config` object both arms print (identical by construction)."""
    return {"workload": WORKLOAD, "rows_per_gpu": rows, "global_shape": [rows * world, N_COLS],
            "parallelism": "row-partition x%d" % world,
            "l2": "inputs (%.1f GB per GPU) larger than L2" % (8.0 * rows * N_COLS / 1e9)}


REF_LABEL = "unmodified reference C code (oracle/_ref) on OpenBLAS 0.3.15, MKL unavailable offline"


def host_matrix_into(buf_nm, m, n, seed=0):
    """Fills buf_nm (numpy (n, m) view of a column-major m x n matrix) with the bench matrix recipe: rank-640 core with
    the reference generator's logspace(1,-3) spectrum plus a 1e-6 noise floor."""
    import numpy as np
    rng = np.random.default_rng(seed)
    r = 640
    X = rng.standard_normal((m, r)) / np.sqrt(m)
    W = rng.standard_normal((n, r)) / np.sqrt(n)
    sig = np.logspace(1, -3, r)
    for j0 in range(0, n, 2048):
        j1 = min(n, j0 + 2048)
        np.matmul(W[j0:j1] * sig, X.T, out=buf_nm[j0:j1])
        buf_nm[j0:j1] += 1e-6 * rng.standard_normal((j1 - j0, m))


def reference_call(L, M, keep=False):
    """One low_rank_svd_rand_decomp_fixed_rank of the compiled reference on the mat* M; wall clock around the API call as
    the reference drivers do (driver_multi_core_mkl5.c:33-36).  Returns (seconds, (U, S, V) numpy or None)."""
    PM = C.POINTER(L.Mat)
    U, Sg, V = PM(), PM(), PM()
    frank = L.I(0)
    L.set_seed(777)
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)      # the reference printf()s progress lines
    t0 = time.perf_counter()
    L.lib.low_rank_svd_rand_decomp_fixed_rank(M, K, P, VNUM, Q, S_ORTH, C.byref(frank), C.byref(U), C.byref(Sg), C.byref(V))
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    os.close(devnull); os.close(saved)
    out = None
    if keep:
        out = (L.from_mat(U, free=False), L.from_mat(Sg, free=False), L.from_mat(V, free=False))
    for x in (U, Sg, V):
        L.lib.matrix_delete(x)
    return dt, out


def set_host_threads():
    cores = os.cpu_count() or 1
    os.environ["OPENBLAS_NUM_THREADS"] = str(cores)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return cores


def run_reference_arm(args, emit):
    """The reference's own CPU implementation on the FULL workload (C2 runs in full), every warm-up and
    timed step a complete low_rank_svd_rand_decomp_fixed_rank call.  N > 1: rank 0 alone runs one GPU's row block."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    cores = set_host_threads()
    from oracle import ref_lib, rsvd_numpy as O
    m, n = args.rows, N_COLS
    budget = float(os.environ.get("BENCH_REF_BUDGET_S", "1500"))
    t_start = time.perf_counter()
    times, steps, warmup = [], max(1, args.steps), max(0, args.warmup)
    if ref_lib.available(32):
        kind = "reference"
        L = ref_lib.RefLib(32)
        M = L.lib.matrix_new(m, n)
        host_matrix_into(np.ctypeslib.as_array(M.contents.d, shape=(n, m)), m, n)
        warm = []
        for it in range(warmup + steps):
            dt, _ = reference_call(L, M)
            (times if it >= warmup else warm).append(dt)
            # safety net: never run into the driver's limit — stop early and report the calls actually made
            if (time.perf_counter() - t_start) + dt > budget:
                break
        if not times:
            times = [warm.pop()]
        L.lib.matrix_delete(M)
        steps, warmup = len(times), len(warm)
    else:
        kind = "port"
        A = np.empty((n, m))
        host_matrix_into(A, m, n)
        A = A.T
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.low_rank_svd_rand_decomp_fixed_rank(A, K, P, VNUM, Q, S_ORTH, 777)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    tf = gemm_flops(m, n, K + P, Q) / sec / 1e12
    sample = "full %dx%d workload (%s), %d warm-up + %d timed calls of low_rank_svd_rand_decomp_fixed_rank; %s" % (
        m, n, "one GPU's row block of the weak-scaling job" if args.gpus > 1 else "BASELINE configs[1]", warmup, steps,
        REF_LABEL if kind == "reference" else "numpy port of the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": bench_config(args.gpus, m),
        "time_to_solution_s": sec, "step_s": times,
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        load = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# north star: BASELINE configs[4] (C5) and configs[3] (C4), device resident on 8 GPUs, reported as sub-records
# ------------------------------------------------------------------------------------------------------------------
def gen_lowrank(torch, m, n, r, lo, noise, seed, rank, m_global):
    """This rank's m rows of A = X diag(logspace(1, lo, r)) W^T + noise, column-major (tensor of shape (n, m)), built in
    HBM in column slabs.  torch only draws the random factors, outside every timed region."""
    g = torch.Generator(device="cuda").manual_seed(seed + 1000 * rank)
    gw = torch.Generator(device="cuda").manual_seed(seed + 7)               # the same W on every rank
    X = torch.randn((m, r), dtype=torch.float64, device="cuda", generator=g) / (m_global ** 0.5)
    W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=gw) / (n ** 0.5)
    sig = torch.logspace(1, lo, r, dtype=torch.float64, device="cuda")
    A = torch.empty((n, m), dtype=torch.float64, device="cuda")
    from lowrankmatrixdecompositioncodes_b200 import device as D
    Xcm = X.t().contiguous()                     # column-major m x r
    Ws = (W * sig).contiguous()                  # (n, r) row-major = column-major r x n: the K-major operand of the library's own GEMM
    step = 1024
    for j0 in range(0, n, step):
        j1 = min(n, j0 + step)
        D.gemm("N", "N", m, j1 - j0, r, Xcm, m, Ws[j0:j1], r, A[j0:j1], m)      # A(:, j0:j1) = X diag(sig) W(j0:j1, :)^T — no cuBLAS anywhere in the run
        A[j0:j1].add_(torch.randn((j1 - j0, m), dtype=torch.float64, device="cuda", generator=g), alpha=noise)
    del X, W, Xcm, Ws
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return A, sig


def north_star_records(lib, torch, dist, D, native, rank, world, peak):
    st = D.stream()
    out = {}

    def sync():
        lib.rsvd_b200_sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    def timed(fn):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
        fn()
        with torch.cuda.stream(st):
            e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- C5: low_rank_svd_rand_decomp_fixed_rank 1,000,000 x 100,000, k=1000 p=50 q=2 ------------------------------------
    small = bool(os.environ.get("BENCH_NORTH_STAR_TEST"))      # code-path check on a small box: same calls, reduced shapes
    m, n, k, p, q = C5["rows"], C5["n"], C5["k"], C5["p"], 2
    if small:
        m, n = 24000, 20000
    l, mg = k + p, m * world
    lib.rsvd_b200_set_option(b"row0", rank * m); lib.rsvd_b200_set_option(b"m_global", mg)
    A, _ = gen_lowrank(torch, m, n, 1280, -3.0, 1e-6, 4321, rank, mg)
    U = D.new_cm(m, k); V = D.new_cm(n, k); Sv = torch.empty(k, dtype=torch.float64, device="cuda")

    def svd():
        native.check(lib.rsvd_b200_svd_rand_dev(A.data_ptr(), m, n, m, k, p, 1, q, 1, 777, None, U.data_ptr(), m, Sv.data_ptr(), V.data_ptr(), n))
    svd()                                   # warm-up (allocations, NCCL channels)
    t = min(timed(svd), timed(svd))
    flops = gemm_flops(mg, n, l, q)
    pe = lib.rsvd_b200_svd_percent_error_dev(A.data_ptr(), m, n, m, U.data_ptr(), m, Sv.data_ptr(), V.data_ptr(), n, k)
    Y = D.new_cm(m, l); Z = D.new_cm(n, l)
    passes = {}
    for name, fn in (("sketch", lambda: native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A.data_ptr(), m, 777, 1, n, 0, Y.data_ptr(), m))),
                     ("TN", lambda: D.gemm("T", "N", n, l, m, A, m, Y, m, Z, n)),
                     ("NN", lambda: D.gemm("N", "N", m, l, n, A, m, Z, n, Y, m))):
        fn()
        tp = timed(fn)
        passes[name] = {"ms": tp * 1e3, "tflops_per_gpu": 2.0 * m * n * l / tp / 1e12, "frac_of_fp64_peak": 2.0 * m * n * l / tp / 1e12 / peak}
    out["c5"] = {"workload": C5["workload"] if not small else "TEST SHAPE %d x %d" % (mg, n), "time_to_solution_s": t, "tflops": flops / t / 1e12,
                 "frac_of_fp64_peak_whole_job": flops / t / 1e12 / (peak * world), "passes": passes, "percent_error": pe,
                 "target": ">= 0.60 of the FP64 roofline per GEMM pass"}
    del A, U, V, Y, Z
    torch.cuda.empty_cache()

    # ---- C4: id_two_sided_rand_decomp_fixed_rank + cur_rand_decomp_fixed_rank 400,000 x 50,000, k=1000 p=20 q=2 -----------
    mg, n, k, p = 400000, 50000, 1000, 20
    if small:
        mg, n = 48000, 36000
    l = k + p
    r0, m = native.row_partition(mg, world, rank)
    lib.rsvd_b200_set_option(b"row0", r0); lib.rsvd_b200_set_option(b"m_global", mg)
    A, sig = gen_lowrank(torch, m, n, 1536, -2.0, 1e-8, 99, rank, mg)
    Icol = torch.empty(n, dtype=torch.float64, device="cuda"); Irow = torch.empty(mg, dtype=torch.float64, device="cuda")
    T = torch.empty((n - k, k), dtype=torch.float64, device="cuda"); Sm = torch.empty((mg - k, k), dtype=torch.float64, device="cuda")
    Cm = D.new_cm(m, k); Um = D.new_cm(k, k); Rm = D.new_cm(k, n)

    def two_sided():
        native.check(lib.rsvd_b200_id_two_sided_rand_dev(A.data_ptr(), m, n, m, k, p, q, 1, 777, Icol.data_ptr(), Irow.data_ptr(), T.data_ptr(), k, Sm.data_ptr(), k))

    def cur():
        native.check(lib.rsvd_b200_cur_rand_dev(A.data_ptr(), m, n, m, k, p, q, 1, 777, Cm.data_ptr(), m, Um.data_ptr(), k, Rm.data_ptr(), k))
    two_sided()
    t_id = timed(two_sided)
    cur()
    t_cur = timed(cur)
    ic = Icol.long()
    rows = torch.randperm(m, device="cuda")[:2048]
    As = A.t()[rows]
    err_id = ((As[:, ic[k:]] - As[:, ic[:k]] @ T.t()).norm() / As.norm()).item()
    err_cur = ((As - Cm.t()[rows] @ Um.t() @ Rm.t()).norm() / As.norm()).item()
    opt = (torch.sqrt((sig[k:] ** 2).sum()) / torch.sqrt((sig ** 2).sum())).item()
    id_flops = (1 + 2 * q) * 2.0 * mg * n * l
    out["c4"] = {"workload": "id_two_sided_rand_decomp_fixed_rank + cur_rand_decomp_fixed_rank %dx%d fp64, k=1000 p=20 q=2, row-partitioned x%d (BASELINE configs[3]%s)" % (mg, n, world, ", TEST SHAPE" if small else ""),
                 "id_two_sided_s": t_id, "cur_s": t_cur, "id_gemm_tflops": id_flops / t_id / 1e12,
                 "id_gemm_floor_s": id_flops / (peak * 1e12 * world),
                 "column_id_rel_err_sampled_rows": err_id, "cur_rel_err_sampled_rows": err_cur, "optimal_rank_k_rel_err": opt}
    del A, T, Sm, Cm, Rm
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated per-pass timing (used under ncu)")
    ap.add_argument("--rows", type=int, default=None, help="rows per GPU (default: the BASELINE workload)")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"], help="c2 = BASELINE configs[1] per GPU (default); c5 = configs[4] (device-resident only)")
    args = ap.parse_args()
    global N_COLS, K, P, WORKLOAD
    if args.config == "c5":
        N_COLS, K, P, WORKLOAD = C5["n"], C5["k"], C5["p"], C5["workload"]
        args.no_e2e = True          # 100 GB per rank does not fit the host; the matrix exists only in HBM
        args.no_cpu_baseline = True
        if args.rows is None:
            args.rows = C5["rows"]
    if args.rows is None:
        args.rows = M_PER_GPU
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner, the reference's progress
    # printf) are redirected to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference_arm(args, emit)

    import numpy as np
    import torch
    import torch.distributed as dist
    import lowrankmatrixdecompositioncodes_b200 as pkg
    from lowrankmatrixdecompositioncodes_b200 import device as D, native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    lib = native.dev()
    assert lib.rsvd_b200_init(local_rank) == 0, lib.rsvd_b200_last_error().decode()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = C.create_string_buffer(128)
            native.check(lib.rsvd_b200_comm_unique_id(buf))
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        ident = ident.cuda()
        dist.broadcast(ident, 0)
        native.check(lib.rsvd_b200_comm_init(rank, world, bytes(ident.cpu().numpy().tobytes())))

    m, n, l = args.rows, N_COLS, K + P
    m_global = m * world
    lib.rsvd_b200_set_option(b"row0", rank * m)
    lib.rsvd_b200_set_option(b"m_global", m_global)
    st = D.stream()

    # synthetic input generated in HBM (rank-r core with the reference generator's logspace(1,-3) spectrum + noise floor);
    # each rank builds only its own rows.  torch draws the random factors; the product is the library's own GEMM (no cuBLAS in the run).
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    r = 640
    X = torch.randn((m, r), dtype=torch.float64, device="cuda", generator=g) / (m_global ** 0.5)
    gw = torch.Generator(device="cuda").manual_seed(99)
    W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=gw) / (n ** 0.5)
    sig = torch.logspace(1, -3, r, dtype=torch.float64, device="cuda")
    A_cm = torch.empty((n, m), dtype=torch.float64, device="cuda")          # column-major m x n
    Xcm = X.t().contiguous()                     # column-major m x r
    Ws = (W * sig).contiguous()                  # (n, r) row-major = column-major r x n: the K-major operand of the library's own GEMM
    for j0 in range(0, n, 4096):
        j1 = min(n, j0 + 4096)
        D.gemm("N", "N", m, j1 - j0, r, Xcm, m, Ws[j0:j1], r, A_cm[j0:j1], m)   # A(:, j0:j1) = X diag(sig) W(j0:j1, :)^T
        A_cm[j0:j1] += 1e-6 * torch.randn((j1 - j0, m), dtype=torch.float64, device="cuda", generator=g)
    del X, W, Xcm, Ws
    U = D.new_cm(m, K); V = D.new_cm(n, K)
    Sv = torch.empty(K, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        lib.rsvd_b200_sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        rc = lib.rsvd_b200_svd_rand_dev(A_cm.data_ptr(), m, n, m, K, P, VNUM, Q, S_ORTH, 777, None,
                                        U.data_ptr(), m, Sv.data_ptr(), V.data_ptr(), n)
        native.check(rc)

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()       # before the barrier: spawning a child of a CUDA process can take 100s of ms
        time.sleep(0.3)
    barrier()
    launches0 = lib.rsvd_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with torch.cuda.stream(st):
        e0.record()
    for i_ in range(args.steps):
        step()
        with torch.cuda.stream(st):
            marks[i_].record()
    with torch.cuda.stream(st):
        e1.record()
    barrier()
    step_ms = [([e0] + marks)[i_].elapsed_time(marks[i_]) for i_ in range(args.steps)]
    launches = lib.rsvd_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_s = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec_per_step = t.item() / args.steps
    flops = gemm_flops(m_global, n, l, Q)
    value = flops / sec_per_step / 1e12
    pct_err = lib.rsvd_b200_svd_percent_error_dev(A_cm.data_ptr(), m, n, m, U.data_ptr(), m, Sv.data_ptr(), V.data_ptr(), n, K)

    # ---- roofline of the dominant kernel: one streaming pass (NN and TN), timed alone on the library's stream --------
    if args.no_roofline:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "TFLOP/s", "ms_per_step": sec_per_step * 1e3, "gpu_launches": int(launches), "step_ms": step_ms, "note": "profiling run"})
        return
    B = torch.randn((l, n), dtype=torch.float64, device="cuda")
    Y = torch.empty((l, m), dtype=torch.float64, device="cuda")
    Z = torch.empty((l, n), dtype=torch.float64, device="cuda")
    kt = {}
    for name, fn in (("NN", lambda: D.gemm("N", "N", m, l, n, A_cm, m, B, n, Y, m)),
                     ("TN", lambda: D.gemm("T", "N", n, l, m, A_cm, m, Y, m, Z, n)),
                     ("sketch", lambda: native.check(lib.rsvd_b200_sketch(b"N", m, l, n, A_cm.data_ptr(), m, 777, 1, n, 0, Y.data_ptr(), m)))):
        for _ in range(2):
            fn()
        lib.rsvd_b200_sync()
        reps = 3
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            a0.record()
        for _ in range(reps):
            fn()
        with torch.cuda.stream(st):
            a1.record()
        lib.rsvd_b200_sync()
        kt[name] = a0.elapsed_time(a1) * 1e-3 / reps
    del B, Y, Z
    pass_flops = 2.0 * m * n * l
    avg_pass = (kt["NN"] + kt["TN"]) / 2
    peak_dmma = max(lib.rsvd_b200_fp64_peak_tflops(4000, 0), lib.rsvd_b200_fp64_peak_tflops(4000, 1))
    peak_nominal = 148 * 64 * 2 * 1.965e9 / 1e12      # 64 FP64 FMA/clk/SM at clocks.max.sm
    peak = max(peak_dmma, 0.0)
    traffic = None
    tf_file = os.path.join(ROOT, "profiles", "r2_gemm_tma_dram_bytes.json")     # from the round-2 `ncu --set full` capture of the same kernels
    if os.path.exists(tf_file):
        try:
            traffic = json.load(open(tf_file)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "achieved": pass_flops / avg_pass / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": pass_flops / avg_pass / 1e12 / peak if peak > 0 else None, "traffic": traffic,
                "kernel": "gemm_tma_kernel (FP64 DMMA, TMA-staged): one pass over A, mean of A*Z (NN) and A^T*Y (TN)",
                "peak_source": "measured in-run with a register-resident FP64 DMMA/DFMA loop (MEASURED_PEAKS.json has no FP64 entry); nominal 64 FMA/clk/SM x 148 SM x 1.965 GHz = %.1f TFLOP/s" % peak_nominal,
                "ms_per_pass": {k_: v * 1e3 for k_, v in kt.items()},
                "tflops_per_pass": {k_: pass_flops / v / 1e12 for k_, v in kt.items()},
                "algorithmic_bytes_per_pass": 8.0 * m * n + 8.0 * l * (m + n),
                "hbm_gbs_implied": (8.0 * m * n + 8.0 * l * (m + n)) / avg_pass / 1e9}

    # ---- e2e: the reference's C API with host buffers -------------------------------------------------------------------
    e2e, cpu, parity = None, None, None
    if not args.no_e2e:
        api = pkg.Api(32 if m * n < 2 ** 31 else 64)
        M = api.lib.matrix_new(m, n)           # pinned host memory (>= 64 MB)
        native.check(lib.rsvd_b200_d2h(C.cast(M.contents.d, C.c_void_p), A_cm.data_ptr(), m * n))   # same matrix as the device arm
        del A_cm  # the API call allocates its own device copy
        torch.cuda.empty_cache()
        api.set_seed(777)
        times, ours = [], None
        n_e2e = 1 + max(1, min(args.steps, 3))
        for it in range(n_e2e):
            Um, Sm, Vm = api.PM(), api.PM(), api.PM()
            frank = api.I(0)
            barrier()
            t0 = time.perf_counter()
            api.lib.low_rank_svd_rand_decomp_fixed_rank(M, K, P, VNUM, Q, S_ORTH, C.byref(frank), C.byref(Um), C.byref(Sm), C.byref(Vm))
            lib.rsvd_b200_sync()
            dt = time.perf_counter() - t0
            api.check()
            if it == n_e2e - 1 and rank == 0 and world == 1 and not args.no_cpu_baseline:
                ours = (api.from_mat(Um, free=False), api.from_mat(Sm, free=False), api.from_mat(Vm, free=False))
            for x in (Um, Sm, Vm):
                api.lib.matrix_delete(x)
            if it > 0:
                times.append(dt)
        te = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": flops / te.item() / 1e12, "unit": "TFLOP/s", "time_to_solution_s": te.item(),
               "h2d_bytes_per_step": 8 * m * n, "d2h_bytes_per_step": 8 * (m * K + K + n * K),
               "api": "low_rank_svd_rand_decomp_fixed_rank(mat*) via librsvd_b200_api%d.so, M in pinned host memory" % api.bits}

        # ---- cpu_baseline + parity: the compiled reference, ONE full run on the very same host matrix and the same Omega ----
        if ours is not None:
            cores = set_host_threads()
            from oracle import ref_lib
            if ref_lib.available(api.bits):
                L = ref_lib.RefLib(api.bits)
                Mref = L.Mat(nrows=m, ncols=n, d=M.contents.d)      # the reference borrows the same buffer (inputs are never modified)
                sec, ref = reference_call(L, C.pointer(Mref), keep=True)
                cpu = {"value": flops / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference", "seconds": sec,
                       "sample": "full %dx%d workload, 1 call on the e2e leg's host matrix; %s" % (m, n, REF_LABEL)}
                (Uo, So, Vo), (Ur, Sr, Vr) = ours, ref
                so, sr = np.diag(So), np.diag(Sr)

                def sin_theta(X, Y):       # ||(I - X X^T) Y||_2: accurate for small angles
                    return float(np.linalg.norm(Y - X @ (X.T @ Y), 2))

                def pct(Uh, Sh, Vh):       # streamed on the device by the library's own evaluation helper (MVF:391-405 semantics)
                    Um, Sm, Vm = api.to_mat(Uh), api.to_mat(Sh), api.to_mat(Vh)
                    api.lib.use_low_rank_svd_for_approximation(M, Um, Sm, Vm)
                    for x in (Um, Sm, Vm):
                        api.lib.matrix_delete(x)
                    return api.lib.rsvd_b200_api_last_percent_error()

                pe_o, pe_r = pct(Uo, So, Vo), pct(Ur, Sr, Vr)
                parity = {"oracle": "oracle/_ref (compiled reference), same host matrix, same Omega (Philox seed 777)",
                          "sigma_rel_err_max": float(np.max(np.abs(so - sr) / sr)), "sin_theta_U": sin_theta(Uo, Ur), "sin_theta_V": sin_theta(Vo, Vr),
                          "percent_error": pe_o, "percent_error_reference": pe_r,
                          "tolerances": {"sigma_rel_err_max": 1e-10, "sin_theta": 1e-8, "percent_error_rel_diff": 0.01}}
                parity["pass"] = bool(parity["sigma_rel_err_max"] <= 1e-10 and parity["sin_theta_U"] <= 1e-8 and parity["sin_theta_V"] <= 1e-8
                                      and abs(pe_o - pe_r) <= 0.01 * pe_r)
                del ours, ref, Uo, So, Vo, Ur, Sr, Vr
        api.lib.matrix_delete(M)

    # ---- north star (BASELINE configs[4] and configs[3]) as sub-records: only on a full 8-GPU box, device resident ----
    north = None
    if args.no_e2e:
        del A_cm
    del U, V
    torch.cuda.empty_cache()
    if (world == 8 or (world > 1 and os.environ.get("BENCH_NORTH_STAR_TEST"))) and args.config == "c2" and not os.environ.get("BENCH_NO_NORTH_STAR"):
        try:
            north = north_star_records(lib, torch, dist, D, native, rank, world, peak)
        except Exception as exc:     # never lose the main line to the extras
            north = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(world, m), "percent_error": pct_err,
            "time_to_solution_s": sec_per_step, "gemm_flops_per_step": flops, "step_ms_rank0": step_ms,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        }
        if north is not None:
            line["north_star"] = north
        emit(line)
    if world > 1:
        lib.rsvd_b200_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
