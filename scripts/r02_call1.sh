#!/bin/bash
# Round-2 GPU call 1: baseline of the code as round 1 left it + the new BASELINE-size parity tests + ncu --set full captures of the
# kernels that only had launch lists.   gpurun --timeout 2400 -- 'bash scripts/r02_call1.sh'
set -u
OUT=gpurun_out/r02a
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    local t0=$(date +%s)
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc after $(( $(date +%s) - t0 )) s; tail:" >> $OUT/SUMMARY.txt
    tail -${TAILN:-6} $OUT/$name.log | cut -c1-400 | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
run 600 tests_baseline_sizes python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x
run 900 tests_all python -m pytest tests -m gpu -q --deselect tests/test_baseline_sizes_gpu.py
TAILN=12 run 600 variants_timing python scripts/variants_timing.py 9 2
run 300 nonlinear_timing python scripts/nonlinear_timing.py 8
run 300 ensemble_timing python scripts/ensemble_timing.py 7 32 l8
run 900 bench python bench.py
grep '^{' $OUT/bench.log | tail -1 > $OUT/bench_n1.json
run 600 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02a.csv \
    python bench.py --steps 2 --warmup 1 --substeps 5 --no-cpu --no-variants
for spec in "cell_step_kernel:profile_sh.py 9 2" "sh_analysis_mf_kernel:profile_sh.py 9 2" "sh_synthesis_mf_kernel:profile_sh.py 9 2" \
            "sh_reduce_solve_kernel:profile_sh.py 9 2" "edge_step_pipe_kernel:profile_sh.py 9 2"; do
    k=${spec%%:*}; cmd=${spec#*:}
    run 400 "ncu_$k" ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/${k}_r02a -f python scripts/$cmd
done
run 400 ncu_nl ncu --set full --clock-control none --import-source on -k regex:nl_ -s 30 -c 8 -o $OUT/nl_kernels_r02a -f python scripts/nonlinear_timing.py 8
run 400 ncu_ens ncu --set full --clock-control none --import-source on -k regex:ens_ -s 30 -c 8 -o $OUT/ens_kernels_r02a -f python scripts/ensemble_timing.py 7 32 l8
log done
