#!/bin/bash
# the slab path on ONE GPU (LQFT_FLAG_SELF_HALO = 64): bench timing at several ghost depths, then a warm launch list
for d in 8 4 2; do
  LQFT_N16_DEPTH=$d python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-parity --flags 64 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('selfhalo depth $d value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']))"
done
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('plain value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']))"
LQFT_N16_DEPTH=8 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --cache-control none --clock-control none --launch-skip 500 -c 20 --csv \
  --log-file gpurun_out/selfhalo_d8.csv python bench.py --steps 1 --warmup 3 --sweeps 40 --no-configs --no-cpu-baseline --no-parity --flags 64 > gpurun_out/selfhalo_ncu.log 2>&1
python tools/launch_table.py gpurun_out/selfhalo_d8.csv 2>/dev/null | head -24
