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
           on the box's host cores, on a bounded row-subsample of the same workload
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
def cpu_reference_run(m_sample, steps, warmup):
    """Times the reference's own low_rank_svd_rand_decomp_fixed_rank (unmodified C sources, OpenBLAS) on an
    m_sample x N_COLS row-subsample with the workload's n, k, p, q.  Returns (seconds per step, cores, kind)."""
    import numpy as np
    cores = os.cpu_count() or 1
    os.environ["OPENBLAS_NUM_THREADS"] = str(cores)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import ref_lib, rsvd_numpy as O
    rng = np.random.default_rng(0)
    r = 640
    X = rng.standard_normal((m_sample, r)) / np.sqrt(m_sample)
    W = rng.standard_normal((N_COLS, r)) / np.sqrt(N_COLS)
    A = (X * np.logspace(1, -3, r)) @ W.T
    times = []
    if ref_lib.available(32):
        L = ref_lib.RefLib(32)
        kind = "reference"
        M = L.to_mat(A)
        del A
        L.set_seed(777)
        PM = C.POINTER(L.Mat)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        for it in range(warmup + steps):
            U, Sg, V = PM(), PM(), PM()
            frank = L.I(0)
            sys.stdout.flush()
            os.dup2(devnull, 1)      # the reference printf()s progress lines
            t0 = time.perf_counter()
            L.lib.low_rank_svd_rand_decomp_fixed_rank(M, K, P, VNUM, Q, S_ORTH, C.byref(frank), C.byref(U), C.byref(Sg), C.byref(V))
            dt = time.perf_counter() - t0
            os.dup2(saved, 1)
            for x in (U, Sg, V):
                L.lib.matrix_delete(x)
            if it >= warmup:
                times.append(dt)
        L.lib.matrix_delete(M)
    else:
        kind = "port"
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.low_rank_svd_rand_decomp_fixed_rank(A, K, P, VNUM, Q, S_ORTH, 777)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    return sum(times) / len(times), cores, kind


def run_reference_arm(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m_sample = 6250     # 1/8 of the workload's rows: ~5e11 GEMM flops + the full-size (n x l) QR and l x l SVD per step
    sec, cores, kind = cpu_reference_run(m_sample, max(1, args.steps), min(args.warmup, 1))
    tf = gemm_flops(m_sample, N_COLS, K + P, Q) / sec / 1e12
    sample = "%dx%d row-subsample (1/8 of the rows), same n,k,p,q; %s" % (
        m_sample, N_COLS, "unmodified reference C code on OpenBLAS 0.3.15 (MKL unavailable offline)" if kind == "reference" else "numpy port")
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "time_to_solution_s": sec,
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
    # each rank builds only its own rows.  torch (cuBLAS) is used for data generation only, outside every timed region.
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    r = 640
    X = torch.randn((m, r), dtype=torch.float64, device="cuda", generator=g) / (m_global ** 0.5)
    gw = torch.Generator(device="cuda").manual_seed(99)
    W = torch.randn((n, r), dtype=torch.float64, device="cuda", generator=gw) / (n ** 0.5)
    sig = torch.logspace(1, -3, r, dtype=torch.float64, device="cuda")
    A_cm = torch.empty((n, m), dtype=torch.float64, device="cuda")          # column-major m x n
    for j0 in range(0, n, 4096):
        j1 = min(n, j0 + 4096)
        torch.matmul(W[j0:j1] * sig, X.t(), out=A_cm[j0:j1])
        A_cm[j0:j1] += 1e-6 * torch.randn((j1 - j0, m), dtype=torch.float64, device="cuda", generator=g)
    del X, W
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
    tf_file = os.path.join(ROOT, "profiles", "r1_gemm_tma_dram_bytes.json")
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
    e2e = None
    if not args.no_e2e:
        api = pkg.Api(32 if m * n < 2 ** 31 else 64)
        M = api.lib.matrix_new(m, n)           # pinned host memory (>= 64 MB)
        native.check(lib.rsvd_b200_d2h(C.cast(M.contents.d, C.c_void_p), A_cm.data_ptr(), m * n))   # same matrix as the device arm
        del A_cm  # the API call allocates its own device copy
        torch.cuda.empty_cache()
        api.set_seed(777)
        times = []
        for it in range(1 + max(1, min(args.steps, 3))):
            Um, Sm, Vm = api.PM(), api.PM(), api.PM()
            frank = api.I(0)
            barrier()
            t0 = time.perf_counter()
            api.lib.low_rank_svd_rand_decomp_fixed_rank(M, K, P, VNUM, Q, S_ORTH, C.byref(frank), C.byref(Um), C.byref(Sm), C.byref(Vm))
            lib.rsvd_b200_sync()
            dt = time.perf_counter() - t0
            api.check()
            s0 = float(Sm.contents.d[0])
            for x in (Um, Sm, Vm):
                api.lib.matrix_delete(x)
            if it > 0:
                times.append(dt)
        api.lib.matrix_delete(M)
        te = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": flops / te.item() / 1e12, "unit": "TFLOP/s", "time_to_solution_s": te.item(),
               "h2d_bytes_per_step": 8 * m * n, "d2h_bytes_per_step": 8 * (m * K + K + n * K),
               "api": "low_rank_svd_rand_decomp_fixed_rank(mat*) via librsvd_b200_api%d.so, M in pinned host memory" % api.bits}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m_s = 6250
        sec, cores, kind = cpu_reference_run(m_s, 1, 0)
        cpu = {"value": gemm_flops(m_s, n, l, Q) / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind,
               "seconds": sec,
               "sample": "%dx%d row-subsample (1/8 of the rows), same n,k,p,q, 1 run; unmodified reference C code on OpenBLAS 0.3.15 (MKL unavailable offline); flop-proportional estimate for the full workload: %.1f s" % (m_s, n, sec * m / m_s)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": m, "global_shape": [m_global, n], "parallelism": "row-partition x%d" % world,
                       "l2": "inputs (%.1f GB per GPU) larger than L2" % (8.0 * m * n / 1e9),
                       "percent_error": pct_err},
            "time_to_solution_s": sec_per_step, "gemm_flops_per_step": flops, "step_ms_rank0": step_ms,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        lib.rsvd_b200_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
