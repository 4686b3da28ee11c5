#!/bin/bash
# ncu captures behind profiles/r02_*: run on the GPU box (gpurun), read here with tools/ncu_summary.py.
#   bash tools/capture_profiles.sh
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
# launch list of one bench step sequence (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --candidates 262144 --fp64-candidates 65536 --e2e-candidates 65536 --no-cpu-baseline --no-configs --no-pageable --no-mode-table > $O/r02_launches_bench.log 2>&1
# tensor sweep: contraction, k* generator, finish (one launch each, third shard of a warm sweep)
$NCU -k regex:tc_sweep_gemm -s 6 -c 1 -o $O/r02_tc_gemm python tools/tc_bench.py 2048 16 3 > /dev/null 2>&1
$NCU -k regex:kstar16_strip -s 6 -c 1 -o $O/r02_kstar python tools/tc_bench.py 2048 16 3 > /dev/null 2>&1
$NCU -k regex:sweep_finish -s 6 -c 1 -o $O/r02_finish python tools/tc_bench.py 2048 16 3 > /dev/null 2>&1
# Gram
$NCU -k regex:gram_sym -s 8 -c 1 -o $O/r02_gram_2048 python tools/gram_bench.py --sizes 2048 --dims 16 > /dev/null 2>&1
$NCU -k regex:gram_sym -s 8 -c 1 -o $O/r02_gram_8192 python tools/gram_bench.py --sizes 8192 --dims 16 > /dev/null 2>&1
# Cholesky steps and the inverse GEMMs, FP64 sweep GEMM
$NCU -k "regex:chol_step|gemm64_dmma" -s 0 -c 60 -o $O/r02_fp64_dense python tools/fp64_sweep_once.py > /dev/null 2>&1
# MAP objective kernels (K5 / K6) at N = 2048 and the whitened fused kernel at a small N
$NCU -k "regex:gram_tile_kernel|lengthscale_grad|gp_scalars|btl_|sum_kernel|gemv_kernel|fbest" -s 0 -c 40 -o $O/r02_map_2048 python tools/map_objective_once.py 2048 16 > /dev/null 2>&1
$NCU -k "regex:map_whitened_fused|small_model" -s 0 -c 6 -o $O/r02_map_small python tools/map_objective_once.py 60 6 > /dev/null 2>&1
ls -la $O/*.ncu-rep
