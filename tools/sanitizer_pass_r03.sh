#!/bin/bash
# compute-sanitizer over the GPU parity tests that exercise the kernels added or rewritten in the last session of round 2 (GPU box):
# chol_step_kernel<2> (pair pivots, cp.async operands, pruned diagonal update, panel form), chol_prepare_kernel, the two-stream
# two-level factorisation, map_pack_kernel / the one-synchronisation MAP objectives, the split length-scale contraction.
set -u
mkdir -p gpurun_out
O=gpurun_out/r03q_sanitizer.txt
{
echo "# compute-sanitizer (one B200, CUDA 12.9): compute-sanitizer --tool <tool> --error-exitcode 9 python -m pytest ..."
echo "# chol  = tests/test_gpu_large_parity.py -k 'pair_pivot or (two_level and 700)'   (chol_step_kernel<1|2>, chol_prepare_kernel, panel form, rank-k updates on two streams)"
echo "# map   = tests/test_gpu_parity.py -k 'map_objectives or general_map_path'        (map_pack_kernel into pinned memory, deferred pivot check, split contraction)"
echo "# small = tests/test_gpu_parity.py -k 'small_model'                               (potf2_inverse_regs_pair inside small_model_kernel)"
} > $O
run() { # tool, label, pytest args...
  local tool=$1 label=$2; shift 2
  local res
  res=$(timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -q -x 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" | tr '\n' ' ')
  printf "%-10s %-6s %s\n" "$tool" "$label" "$res" >> $O
}
run memcheck  chol  tests/test_gpu_large_parity.py -k "pair_pivot or (two_level and 700)"
run memcheck  map   tests/test_gpu_parity.py -k "map_objectives or general_map_path"
run memcheck  small tests/test_gpu_parity.py -k "small_model"
run synccheck chol  tests/test_gpu_large_parity.py -k "pair_pivot or (two_level and 700)"
run initcheck chol  tests/test_gpu_large_parity.py -k "pair_pivot or (two_level and 700)"
run initcheck map   tests/test_gpu_parity.py -k "map_objectives or general_map_path"
run racecheck chol  tests/test_gpu_large_parity.py -k "pair_pivot"
run racecheck small tests/test_gpu_parity.py -k "small_model"
cat $O
