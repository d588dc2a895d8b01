#!/bin/bash
# Round-2 GPU call 12 (4 GPUs): partitioned tests after the epoch fix (the publisher records the epoch), small grids included.
set -u
OUT=gpurun_out/r02l
mkdir -p $OUT
export ODIS_B200_WAIT_TIMEOUT_S=4
timeout 400 python -m pytest tests/test_multigpu.py tests/test_self_gravity_step_gpu.py tests/test_run_gpu.py tests/test_variant_ids16_gpu.py -m gpu -q > $OUT/tests_partitioned.log 2>&1
echo "tests exit $?" | tee -a $OUT/SUMMARY.txt
tail -5 $OUT/tests_partitioned.log | tee -a $OUT/SUMMARY.txt
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 100 $TR4 --master-port 29951 scripts/partitioned_debug.py 5 2 100 12 0 > $OUT/time_n4_l5_sg.log 2>&1; tail -1 $OUT/time_n4_l5_sg.log | tee -a $OUT/SUMMARY.txt
timeout 100 $TR4 --master-port 29952 scripts/partitioned_debug.py 9 2 100 12 0 > $OUT/time_n4_l9_sg.log 2>&1; tail -1 $OUT/time_n4_l9_sg.log | tee -a $OUT/SUMMARY.txt
