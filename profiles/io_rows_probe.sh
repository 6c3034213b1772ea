#!/bin/bash
# One GPU-box call for the rows SURVEY.md 8f adds (reader ring, GPU text dump, Kaarme file): mints the Kaarme-file
# fixture, runs smoke(), and times the CLI stages new vs previous host code.  Output: gpurun_out/{fixture,smoke,io_probe}.log
set -u
OUT=gpurun_out; mkdir -p $OUT
PKG=canonical-k-mer-hash-table_b200
timeout 20 $PKG/kaarme tests/golden/g2_reads.fa 21 -s 200000 -a 2 -t 4 -o /dev/shm/fx.txt --dump-kaarme $OUT/g2_reads_k21.kaarme > $OUT/fixture.log 2>&1
echo "fixture rc=$?" >> $OUT/fixture.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/smoke.log
timeout 30 python profiles/gen_io_probe.py /dev/shm/g.fasta /dev/shm/g30.fasta > $OUT/io_probe.log 2>&1
nproc >> $OUT/io_probe.log
run() { # label, exe, args...
    local label=$1; shift
    local t0=$(date +%s%N)
    timeout 25 "$@" > /dev/shm/cli.log 2>&1
    local rc=$? t1=$(date +%s%N)
    grep -E "Time used|GPU x|Hash table is full" /dev/shm/cli.log | sed "s/^/[$label] /" >> $OUT/io_probe.log
    echo "[$label] rc=$rc wall=$(( (t1 - t0) / 1000000 )) ms out_bytes=$(stat -c %s /dev/shm/o.txt 2>/dev/null)" >> $OUT/io_probe.log
    rm -f /dev/shm/o.txt
}
# reader-bound: 610 MB of FASTA, nothing written (-a 1000)
run "new  read+count 610MB" $PKG/kaarme /dev/shm/g30.fasta 51 -m 0 -s 50000000 -a 1000 -t 16 -o /dev/shm/o.txt
[ -x $PKG/kaarme_prev ] && run "prev read+count 610MB" $PKG/kaarme_prev /dev/shm/g30.fasta 51 -m 0 -s 50000000 -a 1000 -t 16 -o /dev/shm/o.txt
run "new  read+count 610MB (2nd)" $PKG/kaarme /dev/shm/g30.fasta 51 -m 0 -s 50000000 -a 1000 -t 16 -o /dev/shm/o.txt
# writer-bound: 20 M lines (1.1 GB of text)
run "new  gpu-format 20M lines" $PKG/kaarme /dev/shm/g.fasta 51 -m 0 -s 50000000 -a 1 -t 16 -o /dev/shm/o.txt
run "new  host-format 20M lines" $PKG/kaarme /dev/shm/g.fasta 51 -m 0 -s 50000000 -a 1 -t 16 -o /dev/shm/o.txt --host-format
[ -x $PKG/kaarme_prev ] && run "prev host-format 20M lines" $PKG/kaarme_prev /dev/shm/g.fasta 51 -m 0 -s 50000000 -a 1 -t 16 -o /dev/shm/o.txt
rm -f /dev/shm/g.fasta /dev/shm/g30.fasta /dev/shm/fx.txt
