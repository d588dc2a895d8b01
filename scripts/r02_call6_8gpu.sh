#!/bin/bash
# Round-2 GPU call 6 (8 GPUs): partitioned tests at world 2 and 4, per-step times of the 655,362-cell step at N = 4 and 8 (merged vs
# separate synthesis, without the term), bench.py under torchrun at N = 8 (with the 10,485,762-cell variant) and N = 4, BASELINE config 5
# (256-member sweep, 32 members per GPU).   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/r02_call6_8gpu.sh'
set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    grep -E "^N=|FAILED|rror|^\{|passed|failed" $OUT/$name.log | tail -${TAILN:-4} | cut -c1-1500 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
run 600 tests_partitioned python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29800
next() { P=$((P + 1)); }
next; run 200 n8_merged $TR --nproc-per-node 8 --master-port $P scripts/partitioned_debug.py 9 2 100 12 0
next; ODIS_B200_MERGED_SYNTH=0 run 200 n8_unmerged $TR --nproc-per-node 8 --master-port $P scripts/partitioned_debug.py 9 2 100 12 0
next; run 200 n8_nosg $TR --nproc-per-node 8 --master-port $P scripts/partitioned_debug.py 9 0 100 12 0
next; run 200 n4_merged $TR --nproc-per-node 4 --master-port $P scripts/partitioned_debug.py 9 2 100 12 0
next; run 200 n4_nosg $TR --nproc-per-node 4 --master-port $P scripts/partitioned_debug.py 9 0 100 12 0
next; run 800 bench_n8 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8
grep '^{' $OUT/bench_n8.log | tail -1 > $OUT/bench_n8.json
next; run 300 bench_n4 $TR --nproc-per-node 4 --master-port $P bench.py --gpus 4 --no-variants
grep '^{' $OUT/bench_n4.log | tail -1 > $OUT/bench_n4.json
next; run 300 ensemble_n8 $TR --nproc-per-node 8 --master-port $P scripts/ensemble_multigpu.py 7 32 8
grep '^{' $OUT/ensemble_n8.log | tail -1 > $OUT/ensemble_n8.json
log done
