#!/bin/bash
# Runs bench.py (kernel-only numbers) against every tuning build in build/tune/.
cd "$(dirname "$0")/.."
for lib in build/tune/*.so; do
  v=$(LQFT_LIB=$PWD/$lib timeout -k 10 120 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-configs ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.4g upd/s  %.2f us/launch frac=%.3f e2e=%.4g clk=%s %s' % (d['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['e2e']['value'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons')))" 2>&1)
  echo "$(basename $lib): $v"
done
