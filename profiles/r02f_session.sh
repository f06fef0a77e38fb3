set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "fused_direction or user_objective" > gpurun_out/t_dir.log 2>&1; rc=$?; tail -5 gpurun_out/t_dir.log; echo "rc_dir=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02f_n1.json 2> gpurun_out/bench_r02f_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r02f_n1.json
FLGPU_FUSED_DIRECTION=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/bench_r02f_n1_nodir.json 2> gpurun_out/bench_r02f_n1_nodir.err; echo "bench nodir rc=$?"; cut -c1-300 gpurun_out/bench_r02f_n1_nodir.json
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu.py::test_fused_direction_is_the_same_algorithm > gpurun_out/t_all5.log 2>&1; echo "rc_all=$?"; tail -3 gpurun_out/t_all5.log
