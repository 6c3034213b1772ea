#!/bin/bash
# 2-GPU debugging call: where do `kaarme --gpus 2` (threads of one process) and `bench.py --gpus 2` stop?
set -u
OUT=gpurun_out; mkdir -p $OUT
EXE=canonical-k-mer-hash-table_b200/kaarme
KG_TRACE=1 KAARME_TIMING=1 timeout 60 $EXE tests/golden/g5_long.fasta 21 -m 0 -s 400000 -a 2 -t 8 --gpus 2 --batch-mb 1 -o /dev/shm/o2.txt > $OUT/dbg_cli.out 2> $OUT/dbg_cli.err; echo "cli rc=$?"
tail -25 $OUT/dbg_cli.err
KG_TRACE=1 KAARME_BENCH_TRACE=100 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 --steps 1 --warmup 1 --scale 0.25 --no-cpu-baseline > $OUT/dbg_bench.out 2> $OUT/dbg_bench.err; echo "bench rc=$?"
grep -v "^\[kg rank" $OUT/dbg_bench.err | tail -40
echo ...; grep "^\[kg rank" $OUT/dbg_bench.err | tail -12
tail -c 600 $OUT/dbg_bench.out
