#!/bin/bash
# ncu launch lists (warm L2) of the slab path on ONE GPU (LQFT_FLAG_SELF_HALO) at several ghost depths
for d in 8 2; do
LQFT_N16_DEPTH=$d ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --cache-control none --clock-control none --launch-skip 500 -c 34 --csv \
  --log-file gpurun_out/r2f_selfhalo_d$d.csv python bench.py --steps 1 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity --flags 64 > gpurun_out/r2f_ncu_d$d.log 2>&1
echo "depth $d rc=$?"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --cache-control none --clock-control none --launch-skip 500 -c 10 --csv \
  --log-file gpurun_out/r2f_plain.csv python bench.py --steps 1 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity > gpurun_out/r2f_ncu_plain.log 2>&1
