set -u
mkdir -p gpurun_out
O=gpurun_out/r02o_split_shard_matrix.txt
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/tc_bench.py 2048 16 ${SH:-8} 2>&1 | grep -E "^tensor " >> $O; }
SH=8 run X=1
SH=8 run SLSGP_TC_SPLIT=4
SH=8 run SLSGP_TC_SPLIT=1
SH=16 run SLSGP_TC_SHARD=18944
SH=16 run SLSGP_TC_SHARD=18944 SLSGP_TC_SPLIT=4
SH=16 run SLSGP_TC_SHARD=18944 SLSGP_TC_SPLIT=1
SH=32 run SLSGP_TC_SHARD=9472
SH=32 run SLSGP_TC_SHARD=9472 SLSGP_TC_SPLIT=1
echo "== dram bytes per tc_sweep_gemm launch (ncu)" >> $O
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for cfg in "X=1" "SLSGP_TC_SPLIT=4" "SLSGP_TC_SHARD=18944" "SLSGP_TC_SHARD=18944 SLSGP_TC_SPLIT=4" "SLSGP_TC_SHARD=9472" "SLSGP_TC_SHARD=9472 SLSGP_TC_SPLIT=1"; do
  echo "-- $cfg" >> $O
  env $cfg ncu $M --clock-control none -k regex:tc_sweep_gemm -s 6 -c 1 --csv python tools/tc_bench.py 2048 16 3 2>/dev/null | grep -E "dram__|gpu__time|lts__" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' >> $O
  env $cfg ncu $M --clock-control none -k regex:kstar16 -s 6 -c 1 --csv python tools/tc_bench.py 2048 16 3 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' '{print "kstar:", $(NF-2), $(NF-1), $NF}' >> $O
done
cat $O
