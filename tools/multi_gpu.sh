#!/bin/bash
# usage: tools/multi_gpu.sh N   — slab parity worker + bench.py on N GPUs of one box; logs under gpurun_out/
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/_multi_worker.py > gpurun_out/r2_multi_parity_${N}gpu.log 2>&1
echo "parity rc=$?"; grep -c "^ok" gpurun_out/r2_multi_parity_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_${N}gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'per gpu',d['value']/d['n_gpus'],'parity',d.get('parity_n',{}).get('ok')); c=d.get('configs',{}).get('config5_1024x1024x128g'); print('config5',c)"
