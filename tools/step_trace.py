"""One device-resident C2 step under `ncu --metrics gpu__time_duration.sum` (use --profile-from-start off style gating via
cudaProfilerStart/Stop): prints nothing itself; the launch list is the product."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lowrankmatrixdecompositioncodes_b200 import native, device as D
lib = native.dev()
m, n, k, p = 50000, 20000, 500, 20
gen = torch.Generator(device="cuda").manual_seed(1)
X = torch.randn((m, 640), dtype=torch.float64, device="cuda", generator=gen) / m ** 0.5
W = torch.randn((n, 640), dtype=torch.float64, device="cuda", generator=gen) / n ** 0.5
A_cm = torch.matmul(W * torch.logspace(1, -3, 640, dtype=torch.float64, device="cuda"), X.t())
A_cm += 1e-6 * torch.randn((n, m), dtype=torch.float64, device="cuda", generator=gen)
del X, W
torch.cuda.synchronize()
D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777); lib.rsvd_b200_sync()
if os.environ.get("RSVD_B200_VERBOSE"):      # 3 = per-phase CUDA-event times on stderr (pipeline.cu: Phase)
    lib.rsvd_b200_set_option(b"verbose", int(os.environ["RSVD_B200_VERBOSE"]))
torch.cuda.cudart().cudaProfilerStart()
D.svd_rand(A_cm, k, p, 1, 2, 1, seed=777); lib.rsvd_b200_sync()
torch.cuda.cudart().cudaProfilerStop()
