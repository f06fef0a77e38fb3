set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -s -k "augmented_lagrangian_fused_probe" > gpurun_out/t_al2.log 2>&1; echo "rc_al=$?"; grep -n "outer\|passed\|failed\|FAILED" gpurun_out/t_al2.log | cut -c1-260 | tail -30
