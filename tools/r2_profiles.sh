#!/bin/bash
# round-2 ncu evidence (one GPU); every profiled command has exited 0 without ncu in the same call first.  The reports
# are summarised on the box (tools/ncu_summary.py, tools/ncu_hot.py) and deleted: only text travels back.
cap() {  # cap <name> <kernel regex> <launch-skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 $*"
    python tools/ncu_summary.py raw gpurun_out/$name.ncu-rep; echo; python tools/ncu_hot.py gpurun_out/$name.ncu-rep 25; } > gpurun_out/${name}_ncu.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
B="python bench.py --steps 3 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity"
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-parity > gpurun_out/r2p_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2p_ncu1.log 2>&1
cap r2_n16_256 k_sweep_n16 450 $B
python tools/probe_slab.py || exit 1
cap r2_n16_slab1024 k_sweep_n16 16 python tools/probe_slab.py
python tools/probe_obs.py energy || exit 1
cap r2_observe_256 k_observe_rows 0 python tools/probe_obs.py energy
cap r2_n16_meas_256 k_sweep_n16 9 python tools/probe_obs.py energy
python tools/probe_cluster.py 12 || exit 1
cap r2_cluster_12x64 k_run_cluster 0 python tools/probe_cluster.py 12
ls -la gpurun_out | tail -12
# warm-cache capture of the headline kernel (what the timed run sees: the field stays in L2)
name=r2_n16_256_warm
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_sweep_n16 --launch-skip 450 -c 1 -o gpurun_out/$name $B > gpurun_out/$name.log 2>&1
{ echo "# ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_sweep_n16 --launch-skip 450 -c 1 $B"
  python tools/ncu_summary.py raw gpurun_out/$name.ncu-rep; echo; python tools/ncu_hot.py gpurun_out/$name.ncu-rep 25; } > gpurun_out/${name}_ncu.txt 2>&1
rm -f gpurun_out/$name.ncu-rep
