#!/bin/bash
# Round-2 GPU call 4 (2 GPUs): the whole -m gpu suite (partitioned tests at world 2 included), bench.py under torchrun at N = 2 with
# the 10,485,762-cell variant, merged vs separate synthesis at N = 2.   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r02_call4_2gpu.sh'
set -u
OUT=gpurun_out/r02d
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
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
run 900 tests_all python -m pytest tests -m gpu -q
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run 300 bench_n2_short $TR --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-variants
ODIS_B200_MERGED_SYNTH=0 run 300 bench_n2_short_unmerged $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-variants
run 300 bench_n1_short python bench.py --steps 10 --warmup 3 --no-variants --no-cpu
run 900 bench_n2_full $TR --master-port 29613 bench.py --gpus 2
grep '^{' $OUT/bench_n2_full.log | tail -1 > $OUT/bench_n2.json
log done
