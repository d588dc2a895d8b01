#!/bin/bash
# Round-2 GPU call 8 (one GPU): the whole -m gpu suite on the current tree; nonlinear step folded (4 launches) against the 6- and
# 8-launch forms; L2 policy of the table rows (ODIS_B200_L2_KEEP_MB) on the grids that fit the L2; ncu of the nonlinear step.
set -u
OUT=gpurun_out/r02h
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-4} $OUT/$name.log | cut -c1-700 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
run 900 tests_all python -m pytest tests -m gpu -q
export TAILN=7
run 300 nonlinear_l8 python scripts/nonlinear_timing.py 8
run 300 nonlinear_l7 python scripts/nonlinear_timing.py 7
export TAILN=1
for lv in 7 8; do
  for sg in 0 2; do
    ODIS_B200_L2_KEEP_MB=0 run 200 l2evict_l${lv}_sg${sg} python scripts/step_cfg_timing.py $lv $sg 0
    ODIS_B200_L2_KEEP_MB=100000 run 200 l2keep_l${lv}_sg${sg} python scripts/step_cfg_timing.py $lv $sg 0
  done
done
ODIS_B200_L2_KEEP_MB=100000 run 200 l2keep_l9_sg2 python scripts/step_cfg_timing.py 9 2 0
run 200 l9_sg2 python scripts/step_cfg_timing.py 9 2 0
export TAILN=3
run 300 bench_short python bench.py --steps 10 --warmup 3 --no-variants --no-cpu
run 300 ncu_nl ncu --set full --clock-control none --import-source on -k "regex:nl_" -s 200 -c 4 -o $OUT/nl_kernels_r02h -f python scripts/nonlinear_timing.py 8 folded
run 200 ncu_nl_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 48 --csv --log-file $OUT/launches_nl_r02h.csv python scripts/nonlinear_timing.py 8 folded
log done
