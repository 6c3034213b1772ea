"""Synthetic read generators for the BASELINE.json configurations.

Generated ON THE GPU with torch (plumbing only; 2 GB of FASTA in about a second) and returned as one uint8
CUDA tensor holding the FASTA text.  Shapes:
  C3  5 Mbp genome, 50x, 150 bp reads, 1 % substitutions, single-line records          k=31
  C4  100 Mbp genome, 20x, 10 kbp reads, no errors, records wrapped at 80 columns       k=21/51/127/255
  C5  1 Gbp genome, 30x, 150 bp reads                                                    k=51 + Bloom
Headers have a fixed width (">r%0*d") so records are rectangular and the layout vectorises.
Each read is reverse-complemented with probability 0.5.
"""
import math

import torch

ASCII = (65, 67, 71, 84)  # A C G T


def make_genome(G, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 4, (G,), dtype=torch.uint8, device=device, generator=g)


def record_bytes(L, wrap, id_digits):
    lines = math.ceil(L / wrap) if wrap else 1
    return 2 + id_digits + 1 + L + lines


def make_reads_fasta(genome, n_reads, L, seed, err=0.0, wrap=0, first_id=0, id_digits=9, chunk_reads=None):
    """FASTA text of n_reads reads of length L sampled uniformly from `genome` (uint8 codes 0..3, CUDA)."""
    dev = genome.device
    G = genome.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(ASCII, dtype=torch.uint8, device=dev)
    lines = math.ceil(L / wrap) if wrap else 1
    Lpad = lines * wrap if wrap else L
    rec = 2 + id_digits + 1 + (Lpad + lines if wrap else L + 1)
    if wrap and Lpad != L:
        raise ValueError("L must be a multiple of wrap (keeps records rectangular)")
    out = torch.empty((n_reads, rec), dtype=torch.uint8, device=dev)
    chunk = chunk_reads or max(1, (64 << 20) // L)
    ar = torch.arange(L, device=dev, dtype=torch.int64)
    pow10 = torch.tensor([10 ** (id_digits - 1 - d) for d in range(id_digits)], device=dev, dtype=torch.int64)
    for b in range(0, n_reads, chunk):
        m = min(chunk, n_reads - b)
        pos = torch.randint(0, G - L + 1, (m,), device=dev, generator=g, dtype=torch.int64)
        r = genome[pos[:, None] + ar[None, :]]
        if err > 0:
            e = torch.rand((m, L), device=dev, generator=g) < err
            sub = torch.randint(1, 4, (m, L), device=dev, generator=g, dtype=torch.uint8)
            r = torch.where(e, (r + sub) & 3, r)
        rev = torch.rand((m,), device=dev, generator=g) < 0.5
        rc = (3 - r).flip(1)
        r = torch.where(rev[:, None], rc, r)
        o = out[b:b + m]
        o[:, 0] = 62   # '>'
        o[:, 1] = 114  # 'r'
        ids = torch.arange(first_id + b, first_id + b + m, device=dev, dtype=torch.int64)
        o[:, 2:2 + id_digits] = ((ids[:, None] // pow10[None, :]) % 10 + 48).to(torch.uint8)
        o[:, 2 + id_digits] = 10
        body = o[:, 3 + id_digits:]
        a = lut[r.long()]
        if wrap:
            bv = body.view(m, lines, wrap + 1)
            bv[:, :, :wrap] = a.view(m, lines, wrap)
            bv[:, :, wrap] = 10
        else:
            body[:, :L] = a
            body[:, L] = 10
    return out.view(-1)


CONFIGS = {
    # name: (genome bp, coverage, read length, error rate, wrap, k, -s slots)
    "C3": dict(G=5_000_000, cov=50, L=150, err=0.01, wrap=0, k=31, slots=160_000_000, seed=42),
    "C4": dict(G=100_000_000, cov=20, L=10_000, err=0.0, wrap=80, k=51, slots=250_000_000, seed=43),
    "C5": dict(G=1_000_000_000, cov=30, L=150, err=0.0, wrap=0, k=51, slots=0, seed=44),
}


def make_config(name, device, scale=1.0, rank=0, world=1, genome_scale=None):
    """-> (fasta uint8 CUDA tensor, meta dict).  With world > 1 every rank draws its own reads (same genome,
    genome length scaled by `genome_scale` (default world) so that per-GPU distinct k-mers stay fixed)."""
    c = dict(CONFIGS[name])
    gs = world if genome_scale is None else genome_scale
    G = int(c["G"] * scale * gs)
    n_reads = int(c["G"] * scale * c["cov"] / c["L"])
    genome = make_genome(G, c["seed"], device)
    fasta = make_reads_fasta(genome, n_reads, c["L"], seed=c["seed"] * 1000 + rank, err=c["err"], wrap=c["wrap"],
                             first_id=rank * n_reads)
    del genome
    meta = dict(c)
    meta.update(G=G, n_reads=n_reads, bytes=fasta.numel(), input_kmers=n_reads * (c["L"] - c["k"] + 1),
                record_bytes=fasta.numel() // n_reads, slots=int(c["slots"] * scale) if c["slots"] else 0)
    return fasta, meta
