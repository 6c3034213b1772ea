"""Design evidence for DESIGN.md section 11 item 1 (minimizer / super-k-mer bucketing), CPU only: on a random 2 Mbp genome, k=51,
canonical m-minimizers under a random order -> super-k-mer length, bytes per k-mer if a batch ships super-k-mers (2-bit
packed bases + an 8-byte descriptor) instead of 16-byte keys, and the load imbalance of 2048 buckets chosen by an
INDEPENDENT hash of the minimizer (the ordering hash itself must not be reused: minimizers are its smallest values).
Result (numpy, seed 3): m=21: mean 16.0 k-mers per super-k-mer (theory (k-m+2)/2), 1.53 B/k-mer vs 16; bucket max/mean 1.54
at only 61 super-k-mers per bucket (Poisson noise; C4 has ~3000 per bucket and batch)."""
import numpy as np, time
rng=np.random.default_rng(3)
G=2_000_000; k=51
g=rng.integers(0,4,G).astype(np.uint64)
def mix(x):
    x=(x^(x>>np.uint64(33)))*np.uint64(0xff51afd7ed558ccd); x=(x^(x>>np.uint64(33)))*np.uint64(0xc4ceb9fe1a85ec53); return x^(x>>np.uint64(33))
for m in (11,15,21,27):
    # forward and rc m-mer codes
    f=np.zeros(G-m+1,np.uint64); r=np.zeros(G-m+1,np.uint64)
    for i in range(m):
        f=(f<<np.uint64(2))|g[i:G-m+1+i]
        r=r|((np.uint64(3)-g[i:G-m+1+i])<<np.uint64(2*i))
    can=np.minimum(f,r)
    h=mix(can)                      # random order on canonical m-mers
    w=k-m+1                         # m-mers per k-mer
    # sliding window minimum of h over w positions
    from numpy.lib.stride_tricks import sliding_window_view
    nk=G-k+1
    mins=np.empty(nk,np.uint64)
    B=200000
    for s in range(0,nk,B):
        e=min(nk,s+B)
        mins[s:e]=sliding_window_view(h[s:e+w-1],w).min(axis=1)
    change=np.flatnonzero(mins[1:]!=mins[:-1])
    runs=np.diff(np.concatenate([[0],change+1,[nk]]))
    # bytes per k-mer: super-k-mer of L k-mers = (L+k-1) bases packed 2 bits + 8 B descriptor
    bytes_per_kmer=((runs+k-1)/4+8).sum()/nk
    # bucket imbalance: k-mers per bucket for 2048 buckets by minimizer hash
    nb=2048
    b=(mix(mins^np.uint64(0x9E3779B97F4A7C15))>>np.uint64(53)).astype(np.int64)   # independent hash of the minimizer, top 11 bits
    load=np.bincount(b,minlength=nb)
    print(f"m={m}: super-k-mers {len(runs)}, mean length {runs.mean():.2f} k-mers (theory (w+1)/2={(w+1)/2:.1f}), bytes/k-mer {bytes_per_kmer:.2f} vs 16; bucket load max/mean {load.max()/load.mean():.2f} min/mean {load.min()/load.mean():.2f}")
