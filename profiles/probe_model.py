"""Design evidence for DESIGN.md section 13, item 1a (CPU only, numpy): open addressing at load factor 0.4 (C4: 100 M
distinct k-mers in 250 M slots), uniformly hashed.  How many probe STEPS does a lookup of a stored key take
 (a) with linear probing over 16-byte slots, one slot per step (what kg_probe_step does today), and
 (b) if one step tests BOTH slots of a 32-byte sector (bucketised linear probing: same memory, same insertion order)?
Every step after the first goes through the warp's probe queue (another shared-memory round trip and another L2
request), so the figure of merit is the mean number of EXTRA steps per k-mer."""
import numpy as np

rng = np.random.default_rng(11)
S = 1 << 22                                   # slots (scaled down; the statistics depend on the load factor only)
N = int(0.4 * S)
home = rng.integers(0, S, N)


def build(home, S, bucket):
    """insert keys in order; returns the number of probe steps a later lookup of each key needs"""
    nb = S // bucket
    fill = np.zeros(nb, np.int64)             # occupied slots per bucket
    steps = np.empty(len(home), np.int64)
    hb = home // bucket
    for i, b in enumerate(hb):                # plain loop: 1.7 M keys, a few seconds
        s = 1
        while fill[b] == bucket:
            b = b + 1 if b + 1 < nb else 0
            s += 1
        fill[b] += 1
        steps[i] = s
    return steps


for bucket, name in ((1, "one 16-byte slot per step (today)"), (2, "both slots of a 32-byte sector per step")):
    st = build(home, S, bucket)
    extra = st - 1
    # a warp probes 32 keys together: steps until its slowest lane is done (what a per-window loop would cost) vs the sum
    w = st[: len(st) // 32 * 32].reshape(-1, 32)
    print(f"{name}: mean steps {st.mean():.3f}, keys needing more than one step {100 * (st > 1).mean():.1f} %, "
          f"extra steps per key {extra.mean():.3f}, slowest lane of a warp {w.max(axis=1).mean():.2f} steps")
