#!/bin/bash
run() { echo "== $1 flags=${2:-64}"; env $1 timeout 200 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline --no-parity --flags ${2:-64} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['avg_launch_us'])"; }
run "A=1" 0
run "A=1"
run "LQFT_N16_DEPTH=4"
