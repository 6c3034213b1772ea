#!/bin/bash
# The CLI's host code (reader ring, writer, sharding, Kaarme files) linked against the CPU test double and run under
# AddressSanitizer+UBSan and ThreadSanitizer.  Not part of pytest (slow to build); usage: bash tests/native/sanitize_cli.sh
set -u
cd "$(dirname "$0")"
make -s -C ../../oracle oracle
G=../golden; rc_all=0
for san in address,undefined thread; do
  d=/tmp/kaarme_san_$(echo $san | tr ',' '_'); mkdir -p $d
  g++ -O1 -g -fsanitize=$san -std=c++17 -fPIC -shared -o $d/libkaarme_gpu_mock.so mock_abi.cpp -L../../oracle -loracle -Wl,-rpath,$(realpath ../../oracle) || exit 1
  g++ -O1 -g -fsanitize=$san -std=c++17 -pthread -I../../include -o $d/kaarme_mock ../../canonical-k-mer-hash-table_b200/host/kaarme_main.cpp \
      -L$d -lkaarme_gpu_mock -lz -Wl,-rpath,'$ORIGIN' || exit 1
  while read -r args; do
    $d/kaarme_mock $args > $d/log.txt 2>&1; rc=$?
    bad=$(grep -c -E 'ERROR: AddressSanitizer|WARNING: ThreadSanitizer|runtime error' $d/log.txt)
    echo "[$san] rc=$rc findings=$bad :: $args"
    [ $rc -ne 0 ] || [ $bad -ne 0 ] && rc_all=1
  done <<ARGS
$G/g5_long.fasta 51 -m 0 -a 2 -t 8 -o $d/o.txt --gpus 4 -b -u 100000 -f 0.01
$G/g2_reads.fa 21 -a 2 -t 6 -s 200000 -o $d/o.txt --dump-kaarme $d/d.kaarme
$d/d.kaarme 21 --from-kaarme -a 2 -o $d/o2.txt
$G/g1_multiline.fasta 127 -m 0 -a 1 -s 200000 -o $d/o.txt --host-format --gpus 3
$G/g3_plain.txt 21 -m 0 -a 1 -s 200000 -o $d/o.txt --gpus 8
ARGS
done
exit $rc_all
