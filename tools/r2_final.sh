#!/bin/bash
# final round-2 evidence on one GPU: full bench line, launch list, cold and warm ncu captures of k_sweep_n16 at 256^3 and
# of the 128^3 half-sweep (one row per warp).  Every profiled command has exited 0 without ncu in the same call first.
cap() {  # cap <name> <kernel regex> <launch-skip> <extra ncu args or ""> <command...>
  local name=$1 rx=$2 skip=$3 extra=$4; shift 4
  ncu --set full --clock-control none $extra --import-source on -k regex:$rx --launch-skip $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  { echo "# ncu --set full --clock-control none $extra --import-source on -k regex:$rx --launch-skip $skip -c 1 $*"
    python tools/ncu_summary.py raw gpurun_out/$name.ncu-rep; echo; python tools/ncu_hot.py gpurun_out/$name.ncu-rep 25; } > gpurun_out/${name}_ncu.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err || { tail -5 gpurun_out/r2_final_bench.err; exit 1; }
tail -1 gpurun_out/r2_final_bench.log | cut -c1-400
B="python bench.py --steps 3 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity"
$B > gpurun_out/r2_final_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_final_ncu1.log 2>&1
python tools/launch_table.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launches.txt 2>&1
cap r2_n16_256 k_sweep_n16 450 "" $B
cap r2_n16_256_warm k_sweep_n16 450 "--cache-control none" $B
python tools/probe_small.py c128 > gpurun_out/r2_final_c128.log 2>&1 || exit 1
cap r2_n16_128_warm k_sweep_n16 30 "--cache-control none" python tools/probe_small.py c128
ls -la gpurun_out | tail -8
