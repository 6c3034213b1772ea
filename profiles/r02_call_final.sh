#!/bin/bash
# Round 2, final single-GPU call on HEAD: smoke, full GPU suite, the heavy full-size cases, the default bench, ncu.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r02f
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 330 python -u -m pytest tests -m gpu -x -q --durations=5 > $OUT/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_tests.log
tail -9 $OUT/${TAG}_tests.log
cp gpurun_out/fullsize_fp_fn.json $OUT/${TAG}_fullsize_fp_fn_default.json 2>/dev/null
timeout 150 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $OUT/${TAG}_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02f_bench.json') if l.startswith('{')][-1])
    print(f"{d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  e2e {d['e2e']['value']/1e9:.2f}  roofline {d['roofline']['frac']:.3f} step {d['roofline']['step']['frac']:.3f}  verify {d['verify']['match']}  file_to_file {d.get('file_to_file')}  cpu {d.get('cpu_baseline')}")
except Exception as e:
    print("no bench line:", e)
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kg_ -c 300 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:kg_skm_insert -s 9 -c 1 -o $OUT/${TAG}_skm_insert -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_insert.log 2>&1
KAARME_FULLSIZE=all timeout 330 python -u -m pytest tests/test_gpu_fullsize_reference.py -m gpu -q -k "C4_k127 or C4_k255 or C4_k51_m2 or C5s" --durations=10 > $OUT/${TAG}_fullsize_all.log 2>&1; echo "fullsize-all pytest rc=$?" >> $OUT/${TAG}_fullsize_all.log
tail -14 $OUT/${TAG}_fullsize_all.log
cp gpurun_out/fullsize_fp_fn.json $OUT/${TAG}_fullsize_fp_fn_all.json 2>/dev/null
