#!/bin/bash
# Round 2, GPU call 5 (one B200): first run of the minimizer-bucketed path (kg_skm_scatter / kg_skm_insert).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02c5_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/r02c5_smoke.log
tail -8 $OUT/r02c5_smoke.log
timeout 1200 python -u -m pytest tests -m gpu -x -q --durations=12 > $OUT/r02c5_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/r02c5_tests.log
tail -25 $OUT/r02c5_tests.log
timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r02c5_bench.json 2> $OUT/r02c5_bench.err; echo "bench rc=$?"
tail -3 $OUT/r02c5_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02c5_bench.json').read().strip().splitlines()[-1])
    print(f"{d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}  e2e {d['e2e']['value']/1e9 if d.get('e2e') else None}  roofline.frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print("no bench line:", e)
PY
KG_INSERT_OCC=6 timeout 240 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > $OUT/r02c5_bench_occ6.json
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02c5_bench_occ6.json').read().strip().splitlines()[-1])
    print(f"KG_INSERT_OCC=6 {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}")
except Exception as e:
    print("occ6: no bench line:", e)
PY
for k in 21 127 255; do
  timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --k $k 2>/dev/null | grep '^{' > $OUT/r02c5_bench_k$k.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r02c5_bench_k$k.json').read().strip().splitlines()[-1])
    print(f"k=$k {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}")
except Exception as e:
    print("k=$k no bench line:", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kg_ -c 400 --csv --log-file $OUT/r02c5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r02c5_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kg_skm_insert -s 9 -c 2 -o $OUT/r02c5_skm_insert -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r02c5_ncu_insert.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kg_skm_scatter -s 9 -c 1 -o $OUT/r02c5_skm_scatter -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r02c5_ncu_scatter.log 2>&1
ls -la $OUT/*.ncu-rep 2>/dev/null | tail -3
