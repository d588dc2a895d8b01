#!/bin/bash
# Round-2 GPU call 13 (one GPU): what the driver runs at round end, on the final tree — smoke, the whole -m gpu suite, bench.py and its
# reference arm — plus the ncu launch list and a --set full capture of the dominant kernel for profiles/.
set -u
OUT=gpurun_out/r02m
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-4} $OUT/$name.log | cut -c1-1200 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
run 200 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run 900 tests_all python -m pytest tests -m gpu -q
export TAILN=2
run 900 bench python bench.py
grep '^{' $OUT/bench.log | tail -1 > $OUT/bench_n1.json
run 600 bench_reference python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference
grep '^{' $OUT/bench_reference.log | tail -1 > $OUT/bench_reference.json
run 200 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_default_r02m.csv python scripts/profile_default.py 9 2
run 200 ncu_edge ncu --set full --clock-control none --import-source on -k regex:edge_step_pipe16 -s 3 -c 1 -o $OUT/edge_step_pipe16_r02m -f python scripts/profile_default.py 9 2
log done
