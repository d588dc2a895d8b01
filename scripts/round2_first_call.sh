#!/bin/bash
# First GPU call of the next round, one box, everything that was built after the last B200 run:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/round2_first_call.sh'
# Order: cheapest and most informative first; every stage has its own time limit and writes into gpurun_out/r02/, so a stage
# that fails or hangs costs its own limit and nothing else. Read gpurun_out/r02/SUMMARY.txt first.
set -u
OUT=gpurun_out/r02
mkdir -p $OUT
log() { echo "== $* ==" | tee -a $OUT/SUMMARY.txt; }
run() {   # run <seconds> <name> <command...>
    local limit=$1 name=$2; shift 2
    log "$name: $*"
    timeout $limit "$@" > $OUT/$name.log 2>&1
    local rc=$?
    echo "   exit $rc; tail:" >> $OUT/SUMMARY.txt
    tail -6 $OUT/$name.log | sed 's/^/   | /' >> $OUT/SUMMARY.txt
    return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1

# 1. the validated tests first, then the files of code that has never run on a B200 (one pytest process each: a sticky CUDA
#    error in one cannot fail the others)
run 900 tests_validated python -m pytest tests -m gpu -x -q -k "not surface_ and not variant_"
for f in tests/test_surface_ops_gpu.py tests/test_surface_planet_gpu.py tests/test_surface_analytical_gpu.py tests/test_surface_hybrid_gpu.py \
         tests/test_surface_sigint_gpu.py tests/test_variant_blocks_gpu.py tests/test_variant_ids16_gpu.py tests/test_variant_prefetch_gpu.py tests/test_variant_sg3_gpu.py \
         tests/test_variant_nl4_gpu.py tests/test_variant_overlap_gpu.py; do
    run 600 "$(basename $f .py)" python -m pytest $f -m gpu -q
done

# 2. timings of the opt-in selections against the default (validates each against the default's fields as well)
run 600 variants_timing python scripts/variants_timing.py 9 2
run 300 nonlinear_timing python scripts/nonlinear_timing.py 8

# 3. the bench line (with the child-process probes: opt-in selections, pipelined e2e, other configs)
run 900 bench python bench.py
grep '^{' $OUT/bench.log | tail -1 > $OUT/bench_n1.json

# 4. ncu: launch list of the bench command, then full captures of the kernels that only have launch lists so far
run 600 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --substeps 5 --no-cpu --no-variants
for spec in "edge_step_pipe_kernel:profile_step.py 9 8" "cell_step_kernel:profile_step.py 9 8" "sh_analysis_mf_kernel:profile_sh.py 9 2" \
            "sh_synthesis_mf_kernel:profile_sh.py 9 2"; do
    k=${spec%%:*}; cmd=${spec#*:}
    run 600 "ncu_$k" ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/${k}_r02 -f python scripts/$cmd
done
log done
