set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fused_multi" > gpurun_out/t_multi4.log 2>&1; rc=$?; tail -5 gpurun_out/t_multi4.log | cut -c1-300; echo "rc_multi=$rc"
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/t_multi4.log | head -20; exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02i_n1.json 2> gpurun_out/bench_r02i_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02i_n1.json
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu.py::test_fused_multi_is_the_same_algorithm --deselect tests/test_gpu.py::test_fused_multi_kernel_bitwise > gpurun_out/t_all6.log 2>&1; echo "rc_all=$?"; tail -3 gpurun_out/t_all6.log
