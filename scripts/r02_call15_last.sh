#!/bin/bash
# Round-2 GPU call 15 (one GPU, under a minute: the round's last GPU seconds): the final tree after the A/B switches were removed.
OUT=gpurun_out/r02o
mkdir -p $OUT
( timeout 15 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 45 python -m pytest tests/test_nonlinear_gpu.py tests/test_variant_nl4_gpu.py tests/test_self_gravity_step_gpu.py tests/test_step_parity_gpu.py -m gpu -q -x -k "not large_grid" 2>&1 | tail -2
  timeout 30 python bench.py --steps 5 --warmup 3 --no-variants --no-cpu 2>&1 | tail -1 | cut -c1-400 ) > $OUT/SUMMARY.txt 2>&1
cat $OUT/SUMMARY.txt
