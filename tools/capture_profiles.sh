#!/bin/bash
# ncu captures behind profiles/r02_*: run on the GPU box (gpurun). Each report is summarised on the box with tools/ncu_summary.py
# (the reports themselves exceed what gpurun copies back) and only the text comes home.
#   bash tools/capture_profiles.sh
set -u
O=gpurun_out
mkdir -p $O /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, count, command...
    local name=$1 k=$2 s=$3 c=$4; shift 4
    $NCU -k "regex:$k" -s $s -c $c -o /tmp/ncu/$name "$@" > /dev/null 2>&1
    python tools/ncu_summary.py full /tmp/ncu/$name.ncu-rep > $O/$name.txt 2>&1
}
# launch list of a short bench run (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file /tmp/ncu/launches.csv \
    python bench.py --steps 2 --warmup 1 --candidates 262144 --fp64-candidates 65536 --e2e-candidates 65536 --no-cpu-baseline --no-configs --no-pageable --no-mode-table > /dev/null 2>&1
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > $O/r02_launches_bench.txt 2>&1
cap r02_tc_gemm tc_sweep_gemm 6 1 python tools/tc_bench.py 2048 16 3
cap r02_kstar "kstar_tc|kstar16" 6 1 python tools/tc_bench.py 2048 16 3
cap r02_finish sweep_finish 6 1 python tools/tc_bench.py 2048 16 3
cap r02_gram_2048 gram_sym 8 1 python tools/gram_bench.py --sizes 2048 --dims 16
cap r02_gram_8192 gram_sym 8 1 python tools/gram_bench.py --sizes 8192 --dims 16
cap r02_chol_step chol_step 8 3 python tools/fp64_sweep_once.py
cap r02_fp64_gemm gemm64_dmma 11 1 python tools/fp64_sweep_once.py
cap r02_map_2048 "gram_tile_kernel|lengthscale_grad|gp_scalars|btl_|sum_kernel|fbest" 0 7 python tools/map_objective_once.py 2048 16
cap r02_map_small "map_whitened_fused|small_model" 1 2 python tools/map_objective_once.py 60 6
cp /tmp/ncu/r02_tc_gemm.ncu-rep $O/ 2>/dev/null
ls -la $O
