# final 1-GPU session of round 2: bench line, ncu (launch list + --set full of K1, K3-with-probe, batched probe, single probe), full GPU suite
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02m_n1.json 2> gpurun_out/bench_r02m_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_r02m_n1.json; tail -3 gpurun_out/bench_r02m_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02m_reference_arm.json 2> gpurun_out/bench_r02m_reference_arm.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_r02m_reference_arm.json
NCU_LIGHT=1 timeout 900 bash profiles/run_ncu.sh r02m; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all7.log 2>&1; echo "rc_all=$?"; tail -4 gpurun_out/t_all7.log | cut -c1-300
