set -x
nproc; free -g | head -2; nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -3
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# ---- phase A: C4 at full size on 1, 2 and 4 GPUs side by side (disjoint GPUs), C3 at full size on the last GPU ----
CUDA_VISIBLE_DEVICES=0 python tools/large_configs.py c4 400000 > gpurun_out/r2e_c4_n1.log 2>&1 &
CUDA_VISIBLE_DEVICES=1,2 $TR --nproc-per-node 2 --master-port 29701 tools/large_configs.py c4 400000 > gpurun_out/r2e_c4_n2.log 2>&1 &
CUDA_VISIBLE_DEVICES=3,4,5,6 $TR --nproc-per-node 4 --master-port 29702 tools/large_configs.py c4 400000 > gpurun_out/r2e_c4_n4.log 2>&1 &
CUDA_VISIBLE_DEVICES=7 python tools/large_configs.py c3 200000 > gpurun_out/r2e_c3_full.log 2>&1 &
wait
tail -9 gpurun_out/r2e_c4_n1.log; tail -7 gpurun_out/r2e_c4_n2.log; tail -7 gpurun_out/r2e_c4_n4.log; tail -8 gpurun_out/r2e_c3_full.log
# ---- phase B: all 8 GPUs ----
$TR --nproc-per-node 8 --master-port 29703 tools/large_configs.py c4 400000 > gpurun_out/r2e_c4_n8.log 2>&1; tail -7 gpurun_out/r2e_c4_n8.log
$TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err; tail -c 600 gpurun_out/r2e_bench_n8.err; cat gpurun_out/r2e_bench_n8.json | cut -c1-3000
$TR --nproc-per-node 8 --master-port 29705 tools/dist_check.py > gpurun_out/r2e_dist_check_n8.log 2>&1; grep -E "OK|FAIL|PASS|rror" gpurun_out/r2e_dist_check_n8.log | tail -14
RSVD_B200_DEVICES=0-7 python tools/mg_check.py > gpurun_out/r2e_mg_check_n8.log 2>&1; grep -E "OK|FAIL|PASS|rror|devices" gpurun_out/r2e_mg_check_n8.log | tail -18
python tools/h2d_scaling.py > gpurun_out/r2e_h2d_scaling.log 2>&1; $TR --nproc-per-node 8 --master-port 29706 tools/h2d_scaling.py >> gpurun_out/r2e_h2d_scaling.log 2>&1; grep H2D gpurun_out/r2e_h2d_scaling.log
