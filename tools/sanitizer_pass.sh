#!/bin/bash
# compute-sanitizer over the GPU parity tests that exercise the kernels added or rewritten in round 2 (GPU box).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02z_sanitizer.txt
{
echo "# compute-sanitizer over GPU parity tests (one B200, CUDA 12.9): compute-sanitizer --tool <tool> --error-exitcode 9 python -m pytest ..."
echo "# tensor      = tests/test_gpu_tensor.py -k 'not full'        (kstar_tc_kernel, tc_pack_qh / xh, sweep_finish_kernel two-phase, second tier)"
echo "# incremental = tests/test_gpu_incremental.py                 (slsgp_set_data_extend: truncate_to_identity_kernel, append_*; memoised gram / factor)"
echo "# maximiser   = tests/test_gpu_parity.py -k 'maximiser or argmax' (ascent_update_kernel warp-per-start)"
} > $O
run() { # tool, label, pytest args...
  local tool=$1 label=$2; shift 2
  local res
  res=$(timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -q -x 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" | tr '\n' ' ')
  printf "%-10s %-12s %s\n" "$tool" "$label" "$res" >> $O
}
run memcheck tensor tests/test_gpu_tensor.py -k "not full"
run memcheck incremental tests/test_gpu_incremental.py
run memcheck maximiser tests/test_gpu_parity.py -k "maximiser or argmax"
run synccheck tensor tests/test_gpu_tensor.py -k "not full"
run initcheck tensor tests/test_gpu_tensor.py -k "not full"
run initcheck incremental tests/test_gpu_incremental.py
run racecheck maximiser tests/test_gpu_parity.py -k "maximiser or argmax"
cat $O
