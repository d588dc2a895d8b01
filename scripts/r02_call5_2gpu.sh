#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): where does the merged partitioned self-gravity step stall in a multi-process run? + per-step times at N = 2
set -u
OUT=gpurun_out/r02e
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    grep -E "^N=|FAILED|rror" $OUT/$name.log | tail -${TAILN:-6} | cut -c1-400 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
export ODIS_B200_WAIT_TIMEOUT_S=3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29700
next() { P=$((P + 1)); }
next; run 200 merged_l9 $TR --master-port $P scripts/partitioned_debug.py 9 2 100 15 0
next; ODIS_B200_COOP=0 run 200 merged_nocoop_l9 $TR --master-port $P scripts/partitioned_debug.py 9 2 100 15 0
next; run 200 merged_nograph_l9 $TR --master-port $P scripts/partitioned_debug.py 9 2 100 15 8
next; ODIS_B200_MERGED_SYNTH=0 run 200 unmerged_l9 $TR --master-port $P scripts/partitioned_debug.py 9 2 100 15 0
next; run 200 nosg_l9 $TR --master-port $P scripts/partitioned_debug.py 9 0 100 15 0
next; run 200 merged_l7 $TR --master-port $P scripts/partitioned_debug.py 7 2 100 15 0
next; run 200 merged_l9_fullpot $TR --master-port $P scripts/partitioned_debug.py 9 2 12 40 0
run 600 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q
tail -3 $OUT/tests_partitioned.log | sed 's/^/   | /' >> $OUT/SUMMARY.txt
log done
