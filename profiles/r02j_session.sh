set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fused_direction or fused_multi or user_objective or scalar_statement or hostsim or fortran_abi" > gpurun_out/t_walk.log 2>&1; rc=$?; tail -5 gpurun_out/t_walk.log | cut -c1-300; echo "rc_walk=$rc"
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/t_walk.log | head -20; exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02j_n1.json 2> gpurun_out/bench_r02j_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02j_n1.json; tail -3 gpurun_out/bench_r02j_n1.err
