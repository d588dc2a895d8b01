#!/bin/bash
# Round-2 GPU call 10 (4 GPUs): flags published with one fence (was: a release store = a fence per neighbour, serial); per-CTA time
# stamps on partitioned runs; bench at N = 4 and N = 2 with the pipelined e2e.
set -u
OUT=gpurun_out/r02j
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
export TAILN=1
run 200 time_n4_l9 $TR4 --master-port 29921 scripts/partitioned_debug.py 9 0 100 12 0
run 200 time_n4_l9_sg $TR4 --master-port 29922 scripts/partitioned_debug.py 9 2 100 12 0
run 200 time_n2_l8 $TR2 --master-port 29923 scripts/partitioned_debug.py 8 0 100 12 0
export TAILN=4
run 200 trace_n4_l9 $TR4 --master-port 29924 scripts/halo_trace.py 9 0 $OUT
run 200 trace_n4_l9_sg $TR4 --master-port 29925 scripts/halo_trace.py 9 2 $OUT
run 200 trace_n2_l8 $TR2 --master-port 29926 scripts/halo_trace.py 8 0 $OUT
run 100 trace_n1_l8 python scripts/halo_trace.py 8 0 $OUT
run 100 trace_n1_l7 python scripts/halo_trace.py 7 0 $OUT
run 300 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q
export TAILN=3
run 400 bench_n4 $TR4 --master-port 29927 bench.py --gpus 4 --no-variants
grep '^{' $OUT/bench_n4.log | tail -1 > $OUT/bench_n4.json
log done
