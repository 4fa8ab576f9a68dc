#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun; outputs in gpurun_out/):
#   1. pytest -m gpu (whole suite, slowest tests listed)
#   2. bench.py with its defaults (the driver's command)
#   3. ncu launch list of the bench command (shortened: 3 steps, no CPU leg, no h=0.08 leg)
#   4. ncu single-pass metric captures of the final persistent k_gmres (h=0.04, streamed matrix) under
#      APPLICATION replay — the kernel waits on its own grid-wide flags, so it is never kernel-replayed
# Every ncu run sits under a hard timeout.
set -u
TAG=${1:-r02}
WHAT=${2:-all}
mkdir -p gpurun_out
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/gputests_$TAG.log 2>&1
  echo "tests rc=$?"; tail -4 gpurun_out/gputests_$TAG.log
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
  echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_${TAG}_n1.json
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary"
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
  echo "launch list rc=$?"
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
  M=$M,lts__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
  M=$M,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
  for r in barrier long_scoreboard short_scoreboard wait membar lg_throttle mio_throttle branch_resolving not_selected no_instruction; do
    M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio
  done
  for orth in mgs cgs2f; do
    timeout -s KILL 300 ncu --replay-mode application --clock-control none --metrics $M -k regex:k_gmres -s 1 -c 1 --csv \
        --log-file gpurun_out/ncu_gmres_${orth}_$TAG.csv python tools/ncu_gmres.py $orth 400 > gpurun_out/ncu_gmres_${orth}_$TAG.log 2>&1
    echo "k_gmres $orth capture rc=$?"; tail -2 gpurun_out/ncu_gmres_${orth}_$TAG.log
  done
fi
ls -la gpurun_out | tail -8
