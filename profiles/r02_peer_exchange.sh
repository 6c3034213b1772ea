#!/bin/bash
# Next round, on N >= 2 B200s (gpurun --gpus 2, then 8): validate and time the fused bucket -> peer-store exchange
#  against the ncclSend/ncclRecv path.  usage: bash profiles/r02_peer_exchange.sh N
set -u
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# parity first: the seven multi-GPU configurations (plain / uneven slices / k=127 / Bloom / partitions), both exchanges
timeout 300 $TR --master-port 29511 tests/multigpu_worker.py > $OUT/r02_multigpu_nccl.log 2>&1; echo "rc=$?" >> $OUT/r02_multigpu_nccl.log
KG_PEER=1 timeout 300 $TR --master-port 29512 tests/multigpu_worker.py > $OUT/r02_multigpu_peer.log 2>&1; echo "rc=$?" >> $OUT/r02_multigpu_peer.log
grep -E "multigpu|rc=|Error|error" $OUT/r02_multigpu_nccl.log | tail -9
grep -E "multigpu|rc=|Error|error" $OUT/r02_multigpu_peer.log | tail -9
timeout 120 bash tests/multigpu_cli_check.sh > $OUT/r02_multigpu_cli.log 2>&1; tail -8 $OUT/r02_multigpu_cli.log
# then the bench, both ways (only meaningful if the parity lines above say OK)
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > $OUT/r02_bench_n${N}_nccl.json
KAARME_PEER=1 timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > $OUT/r02_bench_n${N}_peer.json
# run length on the wire vs table region size: partitions per owner in peer mode
: > $OUT/r02_peer_partitions_n${N}.jsonl
for p in 16 32 64 128; do
  echo "# partitions=$p" >> $OUT/r02_peer_partitions_n${N}.jsonl
  KAARME_PEER=1 timeout 200 $TR --master-port $((29520 + p)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --partitions $p 2>/dev/null | grep '^{' >> $OUT/r02_peer_partitions_n${N}.jsonl
done
# end-to-end gap: pipelined H2D in kg_feed, NCCL exchange
KG_FEED_PREFETCH=1 timeout 300 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > $OUT/r02_bench_n${N}_prefetch.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n${N}_prefetch.json")); print("prefetch", f"{d['value']/1e9:.2f} G k-mers/s device  e2e {d['e2e']['value']/1e9:.2f}")
except Exception as e:
    print("prefetch: no result", e)
for line in open("gpurun_out/r02_peer_partitions_n${N}.jsonl"):
    if line.startswith("#"): print(line.strip()); continue
    d = json.loads(line); print(f"  {d['value']/1e9:.2f} G k-mers/s  {d['ms_per_step']:.2f} ms/step  stages {d.get('stage_ms_per_step')}")
for tag in ("nccl", "peer"):
    try:
        d = json.load(open("gpurun_out/r02_bench_n${N}_%s.json" % tag))
        print(tag, f"{d['value']/1e9:.2f} G k-mers/s  {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']/1e9:.2f}  {d['config']['parallelism']}")
    except Exception as e:
        print(tag, "no result:", e)
PY
