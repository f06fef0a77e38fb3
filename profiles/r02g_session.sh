# 2-GPU box: AL fused probe tests, row-shard parity (tests/gpu_multi.py) with the K3 probe, bench --gpus 2 (parity record inside)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "augmented_lagrangian" > gpurun_out/t_al.log 2>&1; echo "rc_al=$?"; tail -4 gpurun_out/t_al.log
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "row_sharded" > gpurun_out/t_multi.log 2>&1; echo "rc_multi=$?"; tail -4 gpurun_out/t_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02g_n2.json 2> gpurun_out/bench_r02g_n2.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/bench_r02g_n2.json; tail -3 gpurun_out/bench_r02g_n2.err
