#!/bin/bash
# per-kernel times of the bucketed path under debug switches (KG_SCATTER_DBG: 0 normal, 1 no global stores, 2 no shared atomics)
for dbg in 0 1 2; do
  KG_SCATTER_DBG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kg_owner_scatter|kg_owner_hist" -c 4 --csv python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --partitions 256 --batch-mb 1024 2>/dev/null | grep -E "kg_owner" | awk -F'","' -v d=$dbg '{print "dbg="d, $5, $(NF)}' | tail -2
done
