#!/bin/bash
# timing matrix of k_run_cluster: chains x cluster size x threads per CTA (40 sweeps, energy after every sweep)
for n in 16 64; do for c in 8 4; do for t in 768 384; do
  echo "== chains=$n ctas=$c threads=$t"
  LQFT_CLUSTER_CTAS=$c LQFT_CLUSTER_THREADS=$t LQFT_DEBUG=1 python tools/probe_cluster.py $n 2>&1 | grep -E "cluster plan|path" | sed -e "s/.*clusters resident/&/" | cut -c1-200 | awk '{print}' | tail -2 | sed -e "s/{.*budget[^}]*}//"
done; done; done
