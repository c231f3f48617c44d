#!/bin/bash
# usage: tools/depth_scan.sh N — bench.py (256^3 per GPU, kernel-only line) on N GPUs for ghost depths 2, 4, 8
N=$1
for d in 8 4 2; do
  LQFT_N16_DEPTH=$d python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$d bench.py --gpus $N --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/depth_${N}gpu_$d.log 2>&1
  tail -1 gpurun_out/depth_${N}gpu_$d.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('depth $d value %.4g per gpu %.4g ms/step %.3f parity %s' % (d['value'], d['value']/d['n_gpus'], d['ms_per_step'], d.get('parity_n',{}).get('ok')))"
done
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('single value %.4g' % d['value'])"
