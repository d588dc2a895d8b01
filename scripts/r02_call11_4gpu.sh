#!/bin/bash
# Round-2 GPU call 11 (4 GPUs): halo warp in the edge kernel (boundary fence off the consumers' path) and the harmonic all-reduce as
# LL lines (one NVLink crossing, no flag, no fence): partitioned tests, per-step times at N = 2 and 4, traces, bench at N = 4.
set -u
OUT=gpurun_out/r02k
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-4} $OUT/$name.log | cut -c1-900 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export ODIS_B200_WAIT_TIMEOUT_S=4
run 300 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q -x
export TAILN=1
run 200 time_n4_l9 $TR4 --master-port 29931 scripts/partitioned_debug.py 9 0 100 12 0
run 200 time_n4_l9_sg $TR4 --master-port 29932 scripts/partitioned_debug.py 9 2 100 12 0
ODIS_B200_MERGED_PART=1 run 200 time_n4_l9_sg_merged $TR4 --master-port 29933 scripts/partitioned_debug.py 9 2 100 12 0
run 200 time_n2_l9 $TR2 --master-port 29934 scripts/partitioned_debug.py 9 0 100 12 0
run 200 time_n2_l9_sg $TR2 --master-port 29935 scripts/partitioned_debug.py 9 2 100 12 0
ODIS_B200_MERGED_PART=1 run 200 time_n2_l9_sg_merged $TR2 --master-port 29936 scripts/partitioned_debug.py 9 2 100 12 0
run 200 time_n2_l8 $TR2 --master-port 29937 scripts/partitioned_debug.py 8 0 100 12 0
export TAILN=4
run 200 trace_n4_l9 $TR4 --master-port 29938 scripts/halo_trace.py 9 0 $OUT
run 200 trace_n4_l9_sg $TR4 --master-port 29939 scripts/halo_trace.py 9 2 $OUT
export TAILN=3
run 400 bench_n4 $TR4 --master-port 29940 bench.py --gpus 4 --no-variants
grep '^{' $OUT/bench_n4.log | tail -1 > $OUT/bench_n4.json
log done
