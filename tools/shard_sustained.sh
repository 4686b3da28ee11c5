B="python bench.py --steps 8 --warmup 3 --no-fp64 --no-mode-table --no-cpu-baseline --no-pageable --no-configs"
for cfg in "X=1" "SLSGP_TC_SHARD=18944" "SLSGP_TC_SHARD=9472" "SLSGP_TC_SHARD=18944 SLSGP_TC_SPLIT=4"; do
  echo "-- $cfg"
  env $cfg $B 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(round(d['value']/1e6,2), round(d['ms_per_step'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_median'], {k: round(v, 1) for k, v in r['kernel_ms'].items()}, round(r['avg_launch_ms'],4), round(r['frac'],4))
"
done
