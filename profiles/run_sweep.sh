#!/bin/bash
# Round-1 evidence sweep on one B200 (writes JSON lines to gpurun_out/sweep.jsonl):
#   C4 k = 21 / 51 / 127 / 255 (multi-word keys), C3 k = 31, C5 per-GPU share (1/8) k = 51 with the double Bloom filter,
#   and the direct (partitions 1) path next to the L2-blocked one at k = 51.
out=gpurun_out/sweep.jsonl; : > $out
run() { echo "# $*" >> $out; python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" 2>gpurun_out/sweep_err.log | grep '^{' >> $out || tail -3 gpurun_out/sweep_err.log >> $out; }
run --k 21
run --k 51
run --k 51 --partitions 1
run --k 127
run --k 255
run --workload C3 --k 31
run --workload C5 --scale 0.125 --k 51 --bloom
python - <<'PY'
import json
for line in open('gpurun_out/sweep.jsonl'):
    if line.startswith('#'): print(line.strip()); continue
    try: d=json.loads(line)
    except Exception: print(line[:200]); continue
    print(f"  value {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:8.2f} ms/step  e2e {d['e2e']['value']/1e9:6.2f} G/s  partitions {d['config']['partitions']}  roofline.frac {d['roofline']['frac']:.3f} ({d['roofline']['kernel']})  bloom {d['config'].get('bloom')}")
PY
