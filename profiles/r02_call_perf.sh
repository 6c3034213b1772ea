#!/bin/bash
# quick single-GPU perf call: parity subset + bench variants
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02p1}
timeout 400 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_reference.py -m gpu -x -q -k "not C3 and not C1 and not C2" > $OUT/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log
run() { # label, env..., -- args
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA 2>/dev/null | grep '^{' > $OUT/${TAG}_$label.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_$label.json').read().strip().splitlines()[-1])
    print(f"$label: {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}  verify {d['verify']['match'] if d.get('verify') else None}")
except Exception as e:
    print("$label: no bench line:", e)
PY
}
EXTRA=""
run occ1 KG_INSERT_OCC=1
run occ5 KG_INSERT_OCC=5
run occ1_noprefetch KG_INSERT_OCC=1 KG_NO_PREFETCH=1
run occ5_noprefetch2 KG_INSERT_OCC=5 KG_NO_PREFETCH=1
EXTRA="--batch-mb 1024"
run occ5_batch1024 KG_INSERT_OCC=5
run occ1_batch1024 KG_INSERT_OCC=1
EXTRA="--batch-mb 512"
run occ5_batch512 KG_INSERT_OCC=5
EXTRA=""
KG_INSERT_OCC=5 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kg_skm_insert -s 9 -c 1 -o $OUT/${TAG}_skm_insert -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_insert.log 2>&1
