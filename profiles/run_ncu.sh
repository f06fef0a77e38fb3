#!/bin/bash
# profiles/run_ncu.sh <tag> -- run on the GPU box through gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01a'
# 1. the bench command plain (must exit 0), 2. its launch list (gpu__time_duration per launch),
# 3. one --set full capture of the L-BFGS kernels K1/K3 and of the line-search kernels in steady state.
# Outputs land in gpurun_out/; summaries are made here with profiles/summarise.py and committed under profiles/.
set -u
TAG=${1:-r02}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary --e2e-steps 2"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
cat gpurun_out/${TAG}_plain.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1_update_dots|k3_direction' -s 24 -c 2 \
    -f -o gpurun_out/${TAG}_k1k3 $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "k1k3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'objective_kernel' -s 300 -c 5 \
    -f -o gpurun_out/${TAG}_ls $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ls rc=$?"
ls -la gpurun_out/
