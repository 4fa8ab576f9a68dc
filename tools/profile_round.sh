#!/bin/bash
# (every ncu run is under a hard timeout: a hung kernel replay once burned 25 GPU-minutes)
# Round profile on ONE B200 (run under gpurun): bench (plain) -> ncu launch list of the same
# command -> one `--set full` capture of the dominant kernel (k_gmres).  Outputs in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-refined"
mkdir -p gpurun_out
$CMD > gpurun_out/bench_profile_plain_$TAG.json 2> gpurun_out/bench_profile_plain_$TAG.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_profile_plain_$TAG.err; exit 1; }
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:k_gmres -s 2 -c 1 -o gpurun_out/gmres_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/gmres_$TAG.ncu-rep --page raw --csv > gpurun_out/gmres_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/gmres_$TAG.ncu-rep --page source --csv > gpurun_out/gmres_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out | tail -12
