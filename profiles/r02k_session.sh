set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fused_multi" > gpurun_out/t_ring.log 2>&1; rc=$?; tail -5 gpurun_out/t_ring.log | cut -c1-300; echo "rc_ring=$rc"
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/t_ring.log | head -20; exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02k_n1_ring.json 2> gpurun_out/bench_r02k_n1_ring.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02k_n1_ring.json; tail -3 gpurun_out/bench_r02k_n1_ring.err
