#!/bin/bash
# Round 2, N-GPU call: sharded parity (torchrun worker + CLI threads) and the weak-scaling bench.  usage: bash profiles/r02_call_multi_gpu.sh N TAG
set -u
N=${1:-2}; TAG=${2:-r02mg}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -u -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > $OUT/${TAG}_tests_n$N.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_tests_n$N.log
grep -E "^multigpu|passed|failed|rc=|Error" $OUT/${TAG}_tests_n$N.log | tail -40
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
    fi
    echo "bench n=$n rc=$?"; tail -2 $OUT/${TAG}_bench_n$n.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/${TAG}_bench_n$n.json') if l.startswith('{')][-1])
    print(f"N=$n {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  e2e {d['e2e']['value']/1e9 if d.get('e2e') else None}  stages {d.get('stage_ms_per_step')}  verify {d.get('verify', {}).get('match') if d.get('verify') else None}")
except Exception as e:
    print("N=$n no bench line:", e)
PY
  fi
done
