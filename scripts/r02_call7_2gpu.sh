#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): partition-overhead tweaks (warp-parallel flag wait, send list prefetched, unmerged synthesis on partitioned
# solvers) against the previous build (libodis_b200_prev.so), at 82k cells per rank (level 8 on 2 GPUs = the per-rank size of 655,362 cells on
# 8) and at level 9; bench.py at N = 2.
set -u
OUT=gpurun_out/r02g
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    grep -E "^N=|FAILED|rror|^\{|passed|failed" $OUT/$name.log | tail -${TAILN:-4} | cut -c1-700 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29900
next() { P=$((P + 1)); }
PREV=$PWD/geodesicodis_b200/libodis_b200_prev.so
for lv in 8 9; do
    next; run 200 new_l${lv}_sg $TR --master-port $P scripts/partitioned_debug.py $lv 2 100 12 0
    next; ODIS_B200_LIB=$PREV ODIS_B200_MERGED_SYNTH=0 run 200 prev_l${lv}_sg $TR --master-port $P scripts/partitioned_debug.py $lv 2 100 12 0
    next; run 200 new_l${lv}_nosg $TR --master-port $P scripts/partitioned_debug.py $lv 0 100 12 0
    next; ODIS_B200_LIB=$PREV run 200 prev_l${lv}_nosg $TR --master-port $P scripts/partitioned_debug.py $lv 0 100 12 0
done
run 200 single_l8_nosg python scripts/step_cfg_timing.py 8 0 0
run 200 single_l7_nosg python scripts/step_cfg_timing.py 7 0 0
run 600 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py tests/test_step_parity_gpu.py -m gpu -q
next; run 600 bench_n2 $TR --master-port $P bench.py --gpus 2 --no-variants
grep '^{' $OUT/bench_n2.log | tail -1 > $OUT/bench_n2.json
log done
