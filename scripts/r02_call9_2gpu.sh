#!/bin/bash
# Round-2 GPU call 9 (2 GPUs): partitioned tests incl. the pipelined I/O, bench at N = 2 (pipelined e2e), and per-CTA time stamps of the
# staged kernels (trace variant library) on partitioned and single runs of the same per-GPU size: where do the 5-18 us per step go?
set -u
OUT=gpurun_out/r02i
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
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
run 300 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q
export TAILN=2
run 200 trace_n2_l8 $TR --master-port 29911 scripts/halo_trace.py 8 0 $OUT
run 200 trace_n2_l9 $TR --master-port 29912 scripts/halo_trace.py 9 0 $OUT
run 200 trace_n2_l9_sg $TR --master-port 29913 scripts/halo_trace.py 9 2 $OUT
run 200 trace_n2_l7 $TR --master-port 29914 scripts/halo_trace.py 7 0 $OUT
run 100 trace_n1_l7 python scripts/halo_trace.py 7 0 $OUT
run 100 trace_n1_l8 python scripts/halo_trace.py 8 0 $OUT
run 100 trace_n1_l6 python scripts/halo_trace.py 6 0 $OUT
export TAILN=1
run 200 time_n2_l8 $TR --master-port 29915 scripts/partitioned_debug.py 8 0 100 12 0
run 200 time_n2_l7 $TR --master-port 29916 scripts/partitioned_debug.py 7 0 100 12 0
export TAILN=3
run 400 bench_n2 $TR --master-port 29917 bench.py --gpus 2 --no-variants
grep '^{' $OUT/bench_n2.log | tail -1 > $OUT/bench_n2.json
log done
