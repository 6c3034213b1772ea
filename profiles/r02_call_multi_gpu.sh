#!/bin/bash
# Round 2, N-GPU call: sharded parity (torchrun worker + CLI threads), C5 scale model over all GPUs, weak-scaling bench.
#   usage: bash profiles/r02_call_multi_gpu.sh N TAG "BENCH_NS"      (every step has its own tight timeout: N GPUs are charged N-fold)
set -u
N=${1:-2}; TAG=${2:-r02mg}; NS=${3:-"$N"}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
export KAARME_MULTIGPU_WORLDS=$N
KAARME_MULTIGPU_CASES=${CASES:-0,1,2,3,5,7,8,9,10,11} timeout 420 python -u -m pytest tests/test_gpu_multigpu.py -m gpu -q -s -k "sharded_counts or 51-2 or 127-0" > $OUT/${TAG}_tests_n$N.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_tests_n$N.log
grep -E "^multigpu|passed|failed|rc=|Error" $OUT/${TAG}_tests_n$N.log | tail -40
KAARME_FULLSIZE=all timeout 300 python -u -m pytest tests/test_gpu_fullsize_reference.py -m gpu -q -s -k "sharded_over_all" > $OUT/${TAG}_c5s_n$N.log 2>&1; echo "c5s pytest rc=$?" >> $OUT/${TAG}_c5s_n$N.log
tail -4 $OUT/${TAG}_c5s_n$N.log
for n in $NS; do
    if [ $n -eq 1 ]; then
      timeout 240 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
    else
      timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
    fi
    echo "bench n=$n rc=$?"; grep -E "VERIFY|Error|error" $OUT/${TAG}_bench_n$n.err | tail -3
    python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/${TAG}_bench_n$n.json') if l.startswith('{')][-1])
    print(f"N=$n {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  e2e {d['e2e']['value']/1e9 if d.get('e2e') else None}  stages {d.get('stage_ms_per_step')}  verify {d.get('verify', {}).get('match') if d.get('verify') else None}")
except Exception as e:
    print("N=$n no bench line:", e)
PY
done
cp gpurun_out/fullsize_fp_fn.json $OUT/${TAG}_fullsize_fp_fn.json 2>/dev/null
