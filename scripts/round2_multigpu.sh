#!/bin/bash
# Second GPU call of the next round (multi-GPU box):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash scripts/round2_multigpu.sh 8'
# Partitioned tests, then the 655,362-cell bench at N = 2, 4, 8 with the default kernels, the 3-launch partitioned self-gravity step
# (kernel_select 16) and that plus 16-bit stencil ids (144). Results: gpurun_out/r02mg/.
set -u
NMAX=${1:-8}
OUT=gpurun_out/r02mg
mkdir -p $OUT
PORT=29511
timeout 900 python -m pytest tests/test_multigpu.py tests/test_variant_ids16_gpu.py tests/test_variant_sg3_gpu.py -m gpu -q -k "partitioned" > $OUT/tests.log 2>&1
echo "tests exit $?" | tee $OUT/SUMMARY.txt; tail -3 $OUT/tests.log >> $OUT/SUMMARY.txt
for n in 2 4 8; do
    [ $n -le $NMAX ] || continue
    for sel in 0 16 144; do
        PORT=$((PORT + 1))
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
            bench.py --gpus $n --steps 10 --warmup 3 --kernel-select $sel > $OUT/bench_n${n}_sel${sel}.log 2>&1
        echo "n=$n sel=$sel exit $?: $(grep '^{' $OUT/bench_n${n}_sel${sel}.log | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["value"], d["unit"], "e2e", d["e2e"]["value"])
except Exception as e: print("no line", e)')" | tee -a $OUT/SUMMARY.txt
    done
done
# BASELINE config 5: 256-member sweep, 32 members per GPU (replicas only)
PORT=$((PORT + 1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port $PORT \
    scripts/ensemble_multigpu.py 7 32 8 > $OUT/ensemble_n${NMAX}.log 2>&1
echo "ensemble n=$NMAX exit $?: $(grep '^{' $OUT/ensemble_n${NMAX}.log | tail -1)" | tee -a $OUT/SUMMARY.txt
