#!/bin/bash
# Multi-GPU drop-in CLI check on a box with >= 2 B200s: `kaarme --gpus 2` must write the same sorted output as one GPU
# (and as the reference golden).  usage: tests/multigpu_cli_check.sh
set -e
EXE=canonical-k-mer-hash-table_b200/kaarme
IN=tests/golden/g5_long.fasta
T=$(mktemp -d)
for k in 31 51; do
  $EXE $IN $k -m 0 -s 400000 -a 1 -o $T/one.$k > $T/log1.$k
  $EXE $IN $k -m 0 -s 400000 -a 1 --gpus 2 --batch-mb 1 -o $T/two.$k > $T/log2.$k
  $EXE $IN $k -m 0 -b -u 100000 -a 2 --gpus 2 -o $T/twob.$k > $T/log2b.$k
  $EXE $IN $k -m 0 -s 400000 -a 2 -o $T/one2.$k > /dev/null
  # fused bucket -> peer-store exchange (kg_peer_connect) instead of ncclSend/ncclRecv
  $EXE $IN $k -m 0 -s 400000 -a 1 --gpus 2 --batch-mb 1 --peer-exchange -o $T/peer.$k > $T/logp.$k
  if cmp -s <(sort $T/one.$k) <(sort $T/peer.$k); then echo "cli --gpus 2 --peer-exchange k=$k: OK"; else echo "cli --gpus 2 --peer-exchange k=$k: MISMATCH"; fi
  if cmp -s <(sort $T/one.$k) <(sort $T/two.$k) && cmp -s <(sort $T/one2.$k) <(sort $T/twob.$k); then echo "cli --gpus 2 k=$k: OK ($(wc -l < $T/two.$k) lines; bloom $(wc -l < $T/twob.$k))"; else echo "cli --gpus 2 k=$k: MISMATCH"; fi
  grep -E "GPU x|Hash table size" $T/log2.$k | tr '\n' ' '; echo
done
python - <<PY
import hashlib, json
cases = json.load(open("tests/golden/golden.json"))
c = [x for x in cases if x["input"] == "g5_long.fasta" and x["k"] == 51 and x["mode"] == 0 and x["a"] == 2 and x["unique"] is None][0]
lines = sorted(open("$T/one2.51", "rb").read().splitlines(keepends=True))
print("golden sha k=51 a=2:", "OK" if hashlib.sha256(b"".join(lines)).hexdigest() == c["sha256"] else "MISMATCH")
PY
rm -rf $T
