#!/bin/bash
# Round-2 GPU call 14 (8 GPUs, short): per-step time of the headline grid on 8 GPUs with the degree-2 term after the exchange fixes.
OUT=gpurun_out/r02n
mkdir -p $OUT
export ODIS_B200_WAIT_TIMEOUT_S=4
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29961 scripts/partitioned_debug.py 9 2 100 8 0 > $OUT/time_n8_l9_sg.log 2>&1
tail -1 $OUT/time_n8_l9_sg.log | tee $OUT/SUMMARY.txt
