# 2-GPU box, 1 minute: the functor device-resident search with its in-kernel rank exchange (tests/gpu_multi.py, prebuilt user objective)
set -u
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/gpu_multi.py 16 tests/link/libuser_objective_prebuilt.so > gpurun_out/r02n_multi2.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/r02n_multi2.log | tail -16 | cut -c1-330
