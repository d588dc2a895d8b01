#!/bin/bash
# Round-2 GPU call 3: merged cell + solve + synthesis kernel (grid barrier, cooperative launch) against the 3-launch form; the rewritten
# bench.py (parity keys, pipelined e2e, L10 on one GPU, reference arm through the host-only library).
set -u
OUT=gpurun_out/r02c
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-4} $OUT/$name.log | cut -c1-600 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
run 300 tests_sg_merged python -m pytest tests/test_self_gravity_step_gpu.py tests/test_self_gravity_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x
ODIS_B200_MERGED_SYNTH=0 run 300 tests_sg_unmerged python -m pytest tests/test_self_gravity_step_gpu.py tests/test_self_gravity_gpu.py -m gpu -q -x
run 900 tests_rest python -m pytest tests -m gpu -q --deselect tests/test_self_gravity_step_gpu.py --deselect tests/test_self_gravity_gpu.py --deselect tests/test_baseline_sizes_gpu.py
export TAILN=1
for lv in 9 8 7; do
    run 200 merged_l${lv}_sg python scripts/step_cfg_timing.py $lv 2 0
    ODIS_B200_MERGED_SYNTH=0 run 200 unmerged_l${lv}_sg python scripts/step_cfg_timing.py $lv 2 0
done
run 200 l9_nosg python scripts/step_cfg_timing.py 9 0 0
run 200 l9_sg4 python scripts/step_cfg_timing.py 9 4 0
run 200 l9_sg8 python scripts/step_cfg_timing.py 9 8 0
export TAILN=3
run 900 bench python bench.py
grep '^{' $OUT/bench.log | tail -1 > $OUT/bench_n1.json
run 600 bench_reference python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference
run 300 ncu_cell_merged ncu --set full --clock-control none --import-source on -k regex:cell_step_pipe -s 3 -c 1 -o $OUT/cell_step_pipe_merged_r02c -f python scripts/profile_default.py 9 2
run 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_default_r02c.csv python scripts/profile_default.py 9 2
run 400 ncu_ens ncu --set full --clock-control none --import-source on -k "regex:ens_(edge|cell|sh)" -s 40 -c 8 -o $OUT/ens_kernels_r02c -f python scripts/ensemble_timing.py 7 32 l8
log done
