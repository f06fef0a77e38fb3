#!/bin/bash
# profiles/run_ncu.sh <tag> -- run on the GPU box through gpurun:
#   gpurun --timeout 1800 -- 'bash profiles/run_ncu.sh r02b'
# 1. the bench command plain (must exit 0), 2. its launch list (gpu__time_duration per launch),
# 3. one --set full capture of the L-BFGS kernels K1/K3 and of the line-search kernels in steady state.
# The .ncu-rep files (40 MB each with the embedded source) exceed what gpurun brings back, so their raw pages are
# exported to CSV on the box (`ncu -i ... --page raw --csv`, the command B200_PROFILING.md reads them with) and the
# reports are dropped; profiles/summarise.py turns the CSVs into the committed summaries.
set -u
TAG=${1:-r02}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary --e2e-steps 2"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
cut -c1-300 gpurun_out/${TAG}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_LAUNCHES:-3000} --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
capture() {   # name, kernel regex, skip, count
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $CMD > gpurun_out/${TAG}_ncu_$1.log 2>&1
    echo "$1 rc=$?"
    ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/${TAG}_$1.source_head.csv
    rm -f gpurun_out/${TAG}_$1.ncu-rep
}
capture k1k3 'k1_update_dots|k3_direction' 24 2
capture multi 'objective_multi_kernel' 150 2     # batched probes (steady state: past the prologue's walks)
capture ls 'objective_kernel' 12 2
[ -z "${NCU_LIGHT:-}" ] && capture tree 'tree_kernel' 60 2
ls -la gpurun_out/ | grep ${TAG}
