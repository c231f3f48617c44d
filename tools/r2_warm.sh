#!/bin/bash
# warm-cache ncu capture of the 256^3 half-sweep (--cache-control none: L2 keeps the field as in the timed run)
B="python bench.py --steps 3 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity"
$B > gpurun_out/r2w_plain.log 2>&1 || exit 1
name=r2_n16_256_warm
ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_sweep_n16 --launch-skip 450 -c 1 -o gpurun_out/$name $B > gpurun_out/$name.log 2>&1
{ echo "# ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_sweep_n16 --launch-skip 450 -c 1 $B"
  python tools/ncu_summary.py raw gpurun_out/$name.ncu-rep; echo; python tools/ncu_hot.py gpurun_out/$name.ncu-rep 40; } > gpurun_out/${name}_ncu.txt 2>&1
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
rm -f gpurun_out/$name.ncu-rep
tail -3 gpurun_out/r2w_plain.log | cut -c1-300
