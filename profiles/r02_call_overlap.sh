#!/bin/bash
# single-GPU: does the next batch's parse + bucketing run BESIDE the persistent insert when the insert leaves room?
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02ov}
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > $OUT/${TAG}_$label.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$label.json').read().strip().splitlines()[-1])
    print(f"$label: {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}  verify {d['verify']['match'] if d.get('verify') else None}")
except Exception as e:
    print("$label: no bench line:", e)
PY
}
run occ1_auto KG_INSERT_OCC=1
run occ1_grid2 KG_INSERT_OCC=1 KG_INSERT_GRID=2
run occ5_auto KG_INSERT_OCC=5
run occ5_grid4 KG_INSERT_OCC=5 KG_INSERT_GRID=4
run occ5_grid3 KG_INSERT_OCC=5 KG_INSERT_GRID=3
run occ6_auto KG_INSERT_OCC=6
run occ6_grid5 KG_INSERT_OCC=6 KG_INSERT_GRID=5
run occ6_grid4 KG_INSERT_OCC=6 KG_INSERT_GRID=4
