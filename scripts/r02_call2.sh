#!/bin/bash
# Round-2 GPU call 2: the new staged cell update (+ folded harmonic analysis) and the solve+synthesis launch: parity on hardware,
# configuration sweep, PDL experiment, ncu of the new kernels.   gpurun --timeout 1800 -- 'bash scripts/r02_call2.sh'
set -u
OUT=gpurun_out/r02b
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-4} $OUT/$name.log | cut -c1-400 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
run 600 tests_step python -m pytest tests/test_step_parity_gpu.py tests/test_self_gravity_step_gpu.py tests/test_self_gravity_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x
run 900 tests_rest python -m pytest tests -m gpu -q --deselect tests/test_step_parity_gpu.py --deselect tests/test_self_gravity_step_gpu.py --deselect tests/test_self_gravity_gpu.py --deselect tests/test_baseline_sizes_gpu.py
export TAILN=1
for cfg in 0 1 2 3; do
    ODIS_B200_CELL_CFG=$cfg run 200 cfg${cfg}_sg python scripts/step_cfg_timing.py 9 2 0
    ODIS_B200_CELL_CFG=$cfg run 200 cfg${cfg}_nosg python scripts/step_cfg_timing.py 9 0 0
done
ODIS_B200_DIRECT_CELL=1 run 200 directcell_sg python scripts/step_cfg_timing.py 9 2 0
ODIS_B200_DIRECT_CELL=1 run 200 directcell_nosg python scripts/step_cfg_timing.py 9 0 0
run 200 wideids_sg python scripts/step_cfg_timing.py 9 2 128
run 200 baseline_sg python scripts/step_cfg_timing.py 9 2 1
for cfg in 0 3; do
    ODIS_B200_PDL=1 ODIS_B200_CELL_CFG=$cfg run 200 pdl_cfg${cfg}_sg python scripts/step_cfg_timing.py 9 2 0
    ODIS_B200_PDL=1 ODIS_B200_CELL_CFG=$cfg run 200 pdl_cfg${cfg}_nosg python scripts/step_cfg_timing.py 9 0 0
done
run 200 l8_sg python scripts/step_cfg_timing.py 8 2 0
run 200 l7_sg python scripts/step_cfg_timing.py 7 2 0
export TAILN=3
for cfg in 0 3; do
    ODIS_B200_CELL_CFG=$cfg run 300 ncu_cell_cfg$cfg ncu --set full --clock-control none --import-source on -k regex:cell_step_pipe -s 3 -c 1 -o $OUT/cell_step_pipe_cfg${cfg}_r02b -f python scripts/profile_default.py 9 2
done
run 300 ncu_synth ncu --set full --clock-control none --import-source on -k regex:sh_bsolve_synthesis -s 3 -c 1 -o $OUT/sh_bsolve_synthesis_r02b -f python scripts/profile_default.py 9 2
run 300 ncu_edge16 ncu --set full --clock-control none --import-source on -k regex:edge_step_pipe16 -s 3 -c 1 -o $OUT/edge_step_pipe16_r02b -f python scripts/profile_default.py 9 2
run 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_default_r02b.csv python scripts/profile_default.py 9 2
log done
