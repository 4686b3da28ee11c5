set -u
mkdir -p gpurun_out
O=gpurun_out/r02q_finish_shard_sustained.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 >> $O
run() { echo "== $*" >> $O; env "$@" python tools/tc_bench.py 2048 16 ${SH:-8} 2>&1 | grep -E "^tensor " >> $O; }
SH=8 run X=1
SH=16 run SLSGP_TC_SHARD=18944
SH=32 run SLSGP_TC_SHARD=9472
echo "== sustained: bench.py 2^24 candidates per step, 6 steps" >> $O
B="python bench.py --steps 6 --warmup 3 --no-fp64 --no-mode-table --no-cpu-baseline --no-pageable --no-configs"
for cfg in "X=1" "SLSGP_TC_SHARD=18944" "SLSGP_TC_SHARD=18944 SLSGP_TC_SPLIT=4" "SLSGP_TC_SPLIT=4" "SLSGP_TC_SHARD=9472"; do
  echo "-- $cfg" >> $O
  env $cfg $B 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(d['value'], d['ms_per_step'], d['clocks'], {k: round(v, 1) for k, v in r['kernel_ms'].items()}, r['avg_launch_ms'], r['frac'])
" >> $O
done
cat $O
