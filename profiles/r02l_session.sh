# 2-GPU box: row-shard parity with batched probes / K3 walk, bench --gpus 2 (parity record inside)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "row_sharded" > gpurun_out/t_multi2.log 2>&1; echo "rc_multi=$?"; tail -4 gpurun_out/t_multi2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02l_n2.json 2> gpurun_out/bench_r02l_n2.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/bench_r02l_n2.json; tail -3 gpurun_out/bench_r02l_n2.err
