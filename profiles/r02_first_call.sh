#!/bin/bash
# First GPU call of the next round (one B200): everything that was written after this round's GPU budget ran out.
#   1. the GPU test suite (export buffers were re-sized, CLI reader/writer changed, error paths touched)
#   2. the rows of DESIGN.md section 9 again (profiles/io_rows_probe.sh): reader ring, GPU text dump, parallel writer
#   3. overlap experiment: blocks per SM of the persistent insert kernel
#   4. TMA-staged parse kernels: parity suite + timing under KG_PARSE_TMA=1
# Output under gpurun_out/: r02_tests.log, r02_tests_tma.log, io_probe.log, r02_insert_grid.jsonl, r02_parse_tma.jsonl.  ~8 min of box time.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -u -m pytest tests -m gpu -x -q --durations=10 > $OUT/r02_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_tests.log
timeout 120 bash profiles/io_rows_probe.sh > /dev/null 2>&1
# the TMA-staged parse kernels (cp.async.bulk + mbarrier, opt-in): the whole parity suite under KG_PARSE_TMA=1, then a timing
KG_PARSE_TMA=1 timeout 300 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q > $OUT/r02_tests_tma.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_tests_tma.log
: > $OUT/r02_parse_tma.jsonl
for t in 0 1; do
  echo "# KG_PARSE_TMA=$t" >> $OUT/r02_parse_tma.jsonl
  KG_PARSE_TMA=$t timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' >> $OUT/r02_parse_tma.jsonl
done
: > $OUT/r02_insert_grid.jsonl
for g in 8 6 5 4 3; do
  echo "# KG_INSERT_GRID=$g" >> $OUT/r02_insert_grid.jsonl
  KG_INSERT_GRID=$g timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' >> $OUT/r02_insert_grid.jsonl
done
# bit-exact Bloom emulation (KG_CFG_REFERENCE_BLOOM): its own test file, not part of the default collection yet
timeout 300 python -u -m pytest tests/experimental_refbloom_gpu.py -x -q > $OUT/r02_tests_refbloom.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_tests_refbloom.log
tail -3 $OUT/r02_tests_refbloom.log
KG_FEED_PREFETCH=1 timeout 200 python -u -m pytest tests/test_gpu_parity.py -x -q > $OUT/r02_tests_prefetch.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_tests_prefetch.log
tail -2 $OUT/r02_tests_prefetch.log
python - <<'PY'
import json
for line in open('gpurun_out/r02_insert_grid.jsonl'):
    if line.startswith('#'): print(line.strip()); continue
    d = json.loads(line)
    print(f"  {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}  roofline.frac {d['roofline']['frac']:.3f}")
PY
tail -3 $OUT/r02_tests.log $OUT/r02_tests_tma.log; cat $OUT/io_probe.log
python - <<'PY'
import json
for line in open('gpurun_out/r02_parse_tma.jsonl'):
    if line.startswith('#'): print(line.strip()); continue
    d = json.loads(line)
    print(f"  {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  stages {d.get('stage_ms_per_step')}")
PY
