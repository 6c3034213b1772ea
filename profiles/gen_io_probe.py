"""Inputs of profiles/io_rows_probe.sh: a 20 Mbp random genome as 10 kbp FASTA records wrapped at 80 columns
(20 M distinct 51-mers: a writer-bound case at -a 1), and the same text 30 times over (~610 MB: a reader-bound case)."""
import sys
import numpy as np

small, big = sys.argv[1], sys.argv[2]
rng = np.random.default_rng(11)
G, L, COLS = 20_000_000, 10_000, 80
bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, G)]
lines = bases.reshape(-1, COLS)
body = np.concatenate([lines, np.full((lines.shape[0], 1), 10, np.uint8)], axis=1).reshape(G // L, -1)   # one row per record
with open(small, "wb") as f:
    for i in range(G // L):
        f.write(b">r%d\n" % i)
        f.write(body[i].tobytes())
blob = open(small, "rb").read()
with open(big, "wb") as f:
    for _ in range(30):
        f.write(blob)
print("small", len(blob), "big", 30 * len(blob))
