#!/bin/bash
# 2-GPU debugging call: where does `kaarme --gpus 2` (two host threads in ONE process) stop?  And bench.py --gpus 2 again.
set -u
OUT=gpurun_out; mkdir -p $OUT
EXE=canonical-k-mer-hash-table_b200/kaarme
KG_TRACE=2 timeout 40 $EXE tests/golden/g5_long.fasta 21 -m 0 -s 400000 -a 2 -t 8 --gpus 2 --batch-mb 1 -o /dev/shm/o2.txt > $OUT/dbg_cli2.out 2> $OUT/dbg_cli2.err; echo "cli (sync after every step) rc=$?"
tail -14 $OUT/dbg_cli2.err
KG_TRACE=1 timeout 40 $EXE tests/golden/g5_long.fasta 21 -m 0 -s 400000 -a 2 -t 8 --gpus 2 --batch-mb 1 -o /dev/shm/o2.txt > $OUT/dbg_cli1.out 2> $OUT/dbg_cli1.err; echo "cli rc=$?"
tail -5 $OUT/dbg_cli1.err
NCCL_P2P_DISABLE=1 KG_TRACE=1 timeout 40 $EXE tests/golden/g5_long.fasta 21 -m 0 -s 400000 -a 2 -t 8 --gpus 2 --batch-mb 1 -o /dev/shm/o2.txt > $OUT/dbg_cli3.out 2> $OUT/dbg_cli3.err; echo "cli NCCL_P2P_DISABLE rc=$?"
tail -5 $OUT/dbg_cli3.err
CUDA_MODULE_LOADING=EAGER KG_TRACE=1 timeout 60 $EXE tests/golden/g5_long.fasta 21 -m 0 -s 400000 -a 2 -t 8 --gpus 2 --batch-mb 1 -o /dev/shm/o2.txt > $OUT/dbg_cli4.out 2> $OUT/dbg_cli4.err; echo "cli CUDA_MODULE_LOADING=EAGER rc=$?"
tail -5 $OUT/dbg_cli4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > $OUT/dbg_bench2.out 2> $OUT/dbg_bench2.err; echo "bench rc=$?"
tail -3 $OUT/dbg_bench2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/dbg_bench2.out') if l.startswith('{')][-1])
    print(f"N=2 {d['value']/1e9:6.2f} G k-mers/s  {d['ms_per_step']:7.2f} ms/step  e2e {d['e2e']['value']/1e9 if d.get('e2e') else None}  stages {d.get('stage_ms_per_step')}  verify {d.get('verify', {}).get('match') if d.get('verify') else None}")
except Exception as e:
    print("no bench line:", e)
PY
