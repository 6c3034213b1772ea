"""SURVEY.md section 8f-4 (bit-exact emulation of the reference's double Bloom filter): the design, proven on the CPU.

The reference's pass 1 (insertion_process, double_bloomfilter.hpp:371-413) looks inherently sequential: whether a window
lands in filter 1 or is promoted to filter 2 depends on every insertion before it.  It is not.  After window s has been
processed all of its bit positions B_s are set in filter 1 whichever branch it took (a filter-2 bit is always a filter-1
bit), so filter 1 just before window t is the union of B_s over s < t, and with
    T1[b] = min { t : b in B_t }                      (one atomicMin of the window ordinal per bit position)
"all of my filter-1 bits were set before me" is  max_{b in B_t} T1[b] < t.  The duplicate-position quirk (two of the nh
hashes coincide on a still-unset bit => the k-mer is ALSO put into filter 2, :401-411) is "some duplicated b has
T1[b] == t".  The promoted windows define T2[b] the same way, and both counters follow.  Two min-arrays and three order-
free sweeps reproduce the single-worker reference bit for bit -- which is what a GPU needs to emulate it.

This test checks that formulation against the oracle's sequential restatement (itself pinned to the reference binary's
new_in_first / new_in_second, tests/golden): filter 2 bit for bit and both counters, including saturated filters and
duplicate positions.  The GPU kernels for it are next round's work."""
import ctypes as C

import numpy as np
import pytest

M54 = (1 << 54) - 1
INF = 1 << 62


def roots_of(seq, k):
    """min(Hf, Hb) of the base-5 polynomial hashes mod 2^54 (hash_functions.cpp:102-232)"""
    c = ["ACGT".index(x) for x in seq]
    p5 = [pow(5, i, 1 << 64) for i in range(k)]
    out = []
    for j in range(len(c) - k + 1):
        w = c[j:j + k]
        hf = sum(w[i] * p5[k - 1 - i] for i in range(k)) & M54
        hb = sum((3 - w[i]) * p5[i] for i in range(k)) & M54
        out.append(min(hf, hb))
    return out


def order_free_pass1(L, roots, m, nh):
    pos = [[L.ko_xxh64_u64(r, L.ko_bloom_seed(i)) & (m - 1) for i in range(nh)] for r in roots]

    def first_times(windows):                 # "atomicMin": any evaluation order gives the same array
        T = {}
        for t in windows:
            for p in pos[t]:
                if T.get(p, INF) > t:
                    T[p] = t
        return T

    def dup_on_fresh_bit(P, T, t):            # a position hit twice by this window that nobody set before it
        seen = set()
        for p in P:
            if p in seen and T.get(p, INF) == t:
                return True
            seen.add(p)
        return False

    order = list(range(len(pos)))
    np.random.default_rng(0).shuffle(order)   # sweeps run in a scrambled order on purpose
    T1 = first_times(order)
    promoted = [t for t in order if all(T1[p] < t for p in pos[t]) or dup_on_fresh_bit(pos[t], T1, t)]
    T2 = first_times(promoted)
    pset = set(promoted)
    new1 = new2 = 0
    for t in order:
        if all(T2.get(p, INF) < t for p in pos[t]):
            continue                          # already in filter 2 when it arrives: nothing happens
        if t in pset:
            new2 += 0 if dup_on_fresh_bit(pos[t], T2, t) else 1
        else:
            new1 += 1
    f2 = np.zeros(m // 8 + 1, np.uint8)
    for p in T2:
        f2[p >> 3] |= 0x80 >> (p & 7)
    return new1, new2, f2


@pytest.mark.parametrize("n,k,U,fpr", [(3000, 21, 50, 0.01), (5000, 21, 300, 0.01), (4000, 31, 2000, 0.05),
                                       (6000, 21, 100, 0.2), (3000, 51, 4000, 0.01)])
def test_order_free_formulation_equals_sequential_reference_semantics(oracle, n, k, U, fpr):
    L = oracle.lib()
    assert roots_of("ACGTTGCAAGGCTTAACCGGT", 21) == [24829212378163]          # SURVEY.md 8c known answer
    rng = np.random.default_rng(n + k)
    g = "".join("ACGT"[x] for x in rng.integers(0, 4, n // 2))
    seq = g + g[: n // 2]                                                     # second half repeats: true duplicates
    data = np.frombuffer((">r\n" + seq + "\n").encode(), np.uint8)
    m, nh, _ = oracle.bloom_params(U, fpr)
    st = oracle.BloomStats()
    f2_seq = np.zeros(m // 8 + 1, np.uint8)
    assert L.ko_bloom_pass1(data.ctypes.data, data.size, k, oracle.FASTA, U, C.c_double(fpr), f2_seq.ctypes.data, C.byref(st)) == 0
    new1, new2, f2 = order_free_pass1(L, roots_of(seq, k), m, nh)
    assert (new1, new2) == (st.new_in_first, st.new_in_second)
    assert (f2 == f2_seq).all()


@pytest.fixture(scope="module")
def refbloom_exe():
    """tests/native/refbloom_host.cu: the product's own per-window functions (csrc/kg_refbloom.cuh, __host__ __device__)
    compiled for the host -- nvcc, sm_100a code generation for the device half of the headers, no GPU touched"""
    import os
    import subprocess
    from conftest import ROOT
    native = os.path.join(ROOT, "tests", "native")
    exe = os.path.join(native, "_build", "refbloom_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(native, "refbloom_host.cu")
    deps = [src] + [os.path.join(ROOT, "canonical-k-mer-hash-table_b200", "csrc", f) for f in ("kg_refbloom.cuh", "kg_refhash.cuh", "kg_count.cuh", "kg_device.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return exe


@pytest.mark.parametrize("n,k,U,fpr,first", [(3000, 21, 50, 0.01, 0), (5000, 21, 300, 0.01, 12345), (4000, 31, 2000, 0.05, 7),
                                             (6000, 21, 100, 0.2, 4_000_000_000), (3000, 51, 4000, 0.01, 1), (2500, 127, 3000, 0.01, 99)])
def test_product_per_window_code_equals_sequential_reference_semantics(oracle, refbloom_exe, n, k, U, fpr, first):
    """the C++ the kernels call per window (kg_rb_window<1|2|3>, kg_rb_admits), run on the CPU in a scrambled order:
    counters, filter 2 and the pass-2 admission of every window == the oracle's sequential execution"""
    import struct
    import subprocess
    L = oracle.lib()
    rng = np.random.default_rng(n * 7 + k)
    g = rng.integers(0, 4, n // 2)
    codes = np.concatenate([g, g[: n // 2]]).astype(np.uint8)
    seq = "".join("ACGT"[x] for x in codes)
    data = np.frombuffer((">r\n" + seq + "\n").encode(), np.uint8)
    m, nh, nh2 = oracle.bloom_params(U, fpr)
    st = oracle.BloomStats()
    f2_seq = np.zeros(m // 8 + 1, np.uint8)
    assert L.ko_bloom_pass1(data.ctypes.data, data.size, k, oracle.FASTA, U, C.c_double(fpr), f2_seq.ctypes.data, C.byref(st)) == 0
    payload = struct.pack("<IIIIII", k, len(codes), m.bit_length() - 1, nh, nh2, first) + codes.tobytes()
    out = subprocess.run([refbloom_exe], input=payload, stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    assert tuple(map(int, out[0].split())) == (st.new_in_first, st.new_in_second)
    bits = np.unpackbits(f2_seq)[:m]
    assert list(map(int, out[1].split())) == [int(b) for b in np.flatnonzero(bits)]
    roots = roots_of(seq, k)
    L.ko_bloom_admits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]     # (a 54-bit root does not fit the default int)
    L.ko_bloom_admits.restype = C.c_int
    want = "".join("1" if L.ko_bloom_admits(f2_seq.ctypes.data, C.byref(st), r) else "0" for r in roots)
    assert out[2] == want
