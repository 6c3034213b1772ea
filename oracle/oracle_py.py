"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product path (canonical-k-mer-hash-table_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "kaarme")
REF_XXH = os.path.join(HERE, "_ref", "xxh64_kat")

FASTA, PLAIN = 0, 2
TABLE_PLAIN, TABLE_KAARME, TABLE_EXACT = 0, 2, -1


def build(ref: bool = False) -> None:
    """(re)build liboracle.so; with ref=True also oracle/_ref/kaarme when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(os.environ.get("KAARME_REF", "/root/reference")):
        subprocess.run(["make", "-s", "-C", HERE, "ref", "-j8"], check=True)


class _Counts(C.Structure):
    _fields_ = [("n", C.c_uint64), ("W", C.c_uint32), ("k", C.c_uint32),
                ("keys", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint64)),
                ("total_windows", C.c_uint64), ("invalid_bytes", C.c_uint64)]


class _Roller(C.Structure):
    _fields_ = [("q", C.c_uint64), ("d", C.c_uint64), ("di", C.c_uint64), ("h", C.c_uint64),
                ("m", C.c_uint64), ("tbm", C.c_int), ("hf", C.c_uint64), ("hb", C.c_uint64),
                ("hashed", C.c_uint64)]


class BloomStats(C.Structure):
    _fields_ = [("m", C.c_uint64), ("nh_ceil", C.c_uint32), ("nh_floor", C.c_uint32),
                ("new_in_first", C.c_uint64), ("new_in_second", C.c_uint64),
                ("table_slots", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.ko_char2int.restype = C.c_uint32
        L.ko_char2int.argtypes = [C.c_uint8]
        L.ko_next_prime3mod4.restype = C.c_uint64
        L.ko_next_prime3mod4.argtypes = [C.c_uint64]
        L.ko_modinv.restype = C.c_uint64
        L.ko_modinv.argtypes = [C.c_int64, C.c_int64]
        L.ko_xxh64_u64.restype = C.c_uint64
        L.ko_xxh64_u64.argtypes = [C.c_uint64, C.c_uint64]
        L.ko_bloom_seed.restype = C.c_uint64
        L.ko_bloom_seed.argtypes = [C.c_uint32]
        L.ko_bloom_params.restype = None
        L.ko_bloom_params.argtypes = [C.c_uint64, C.c_double, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.ko_roller_init.argtypes = [C.POINTER(_Roller), C.c_uint64, C.c_uint64, C.c_int]
        L.ko_roller_update.argtypes = [C.POINTER(_Roller), C.c_uint64, C.c_uint64]
        L.ko_count.restype = C.c_int
        L.ko_count.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_int, C.POINTER(_Counts)]
        L.ko_counts_free.argtypes = [C.POINTER(_Counts)]
        L.ko_reported_count.restype = C.c_uint64
        L.ko_reported_count.argtypes = [C.c_uint64, C.c_int]
        L.ko_format.restype = C.c_size_t
        L.ko_format.argtypes = [C.POINTER(_Counts), C.c_uint64, C.c_int, C.c_void_p, C.c_size_t]
        L.ko_bloom_pass1.restype = C.c_int
        L.ko_bloom_pass1.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint64, C.c_double,
                                     C.c_void_p, C.POINTER(BloomStats)]
        L.ko_count_bloom.restype = C.c_int
        L.ko_count_bloom.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint64, C.c_double,
                                     C.POINTER(_Counts), C.POINTER(BloomStats)]
        L.ko_kaarme_decode.restype = C.c_int64
        L.ko_kaarme_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64,
                                       C.c_void_p]
        _lib = L
    return _lib


class Counts:
    """Sorted (keys[n,W] uint64, counts[n] uint64) + totals."""

    def __init__(self, keys, counts, k, total_windows, invalid_bytes=0):
        self.keys, self.counts, self.k = keys, counts, k
        self.W = (k + 31) // 32
        self.total_windows, self.invalid_bytes = total_windows, invalid_bytes

    @property
    def n(self):
        return len(self.counts)

    def reported(self, table_mode):
        if table_mode == TABLE_PLAIN:
            return self.counts & np.uint64(0xFFFF)
        if table_mode == TABLE_KAARME:
            return np.minimum(self.counts, np.uint64(16383))
        return self.counts

    def filtered(self, min_abundance, table_mode=TABLE_EXACT):
        rep = self.reported(table_mode)
        sel = rep >= np.uint64(min_abundance) if min_abundance > 0 else np.zeros(len(rep), bool)
        return Counts(self.keys[sel], rep[sel], self.k, self.total_windows, self.invalid_bytes)

    def text(self, min_abundance=1, table_mode=TABLE_EXACT) -> bytes:
        c = _Counts()
        keys = np.ascontiguousarray(self.keys, dtype=np.uint64)
        cnts = np.ascontiguousarray(self.counts, dtype=np.uint64)
        c.n, c.W, c.k = self.n, self.W, self.k
        c.keys = keys.ctypes.data_as(C.POINTER(C.c_uint64))
        c.counts = cnts.ctypes.data_as(C.POINTER(C.c_uint64))
        need = lib().ko_format(C.byref(c), min_abundance, table_mode, None, 0)
        buf = C.create_string_buffer(need + 1)
        lib().ko_format(C.byref(c), min_abundance, table_mode, buf, need)
        return buf.raw[:need]


def _as_bytes(data):
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(data, dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


def _take(c: _Counts) -> Counts:
    n, W = int(c.n), int(c.W)
    keys = np.ctypeslib.as_array(c.keys, shape=(max(n, 1) * W,))[: n * W].copy().reshape(n, W)
    counts = np.ctypeslib.as_array(c.counts, shape=(max(n, 1),))[:n].copy()
    out = Counts(keys, counts, int(c.k), int(c.total_windows), int(c.invalid_bytes))
    lib().ko_counts_free(C.byref(c))
    return out


def count(data, k, input_mode=FASTA, starts_in_header=False) -> Counts:
    a = _as_bytes(data)
    c = _Counts()
    rc = lib().ko_count(a.ctypes.data, a.size, k, input_mode, int(starts_in_header), C.byref(c))
    if rc != 0:
        raise RuntimeError(f"ko_count rc={rc}")
    return _take(c)


def count_bloom(data, k, expected_unique, fpr=0.01, input_mode=FASTA):
    a = _as_bytes(data)
    c, st = _Counts(), BloomStats()
    rc = lib().ko_count_bloom(a.ctypes.data, a.size, k, input_mode, expected_unique, fpr, C.byref(c),
                              C.byref(st))
    if rc != 0:
        raise RuntimeError(f"ko_count_bloom rc={rc}")
    return _take(c), st


def bloom_pass1(data, k, expected_unique, fpr=0.01, input_mode=FASTA) -> BloomStats:
    a = _as_bytes(data)
    st = BloomStats()
    rc = lib().ko_bloom_pass1(a.ctypes.data, a.size, k, input_mode, expected_unique, fpr, None, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"ko_bloom_pass1 rc={rc}")
    return st


def bloom_params(expected_unique, fpr):
    m, a, b = C.c_uint64(), C.c_uint32(), C.c_uint32()
    lib().ko_bloom_params(expected_unique, fpr, C.byref(m), C.byref(a), C.byref(b))
    return m.value, a.value, b.value


def rolling_hashes(codes, k, q, tbm=False):
    """(Hf, Hb) after pushing the 2-bit codes of one window prefix, as the functors drive the hasher."""
    r = _Roller()
    lib().ko_roller_init(C.byref(r), q, k, int(tbm))
    window = []
    for c in codes:
        out = window[0] if len(window) == k else 0
        lib().ko_roller_update(C.byref(r), int(c), int(out))
        window.append(int(c))
        if len(window) > k:
            window.pop(0)
    return r.hf, r.hb


def kaarme_decode(table, roots, k, slot):
    table = np.ascontiguousarray(table, dtype=np.uint64)
    roots = np.ascontiguousarray(roots, dtype=np.uint64)
    out = np.zeros(k, dtype=np.uint8)
    hops = lib().ko_kaarme_decode(table.ctypes.data, table.size, roots.ctypes.data, k, slot, out.ctypes.data)
    return hops, out


def key_strings(keys, k):
    """[n,W] uint64 keys -> list of str"""
    W = (k + 31) // 32
    out = []
    for row in np.asarray(keys, dtype=np.uint64).reshape(-1, W):
        v = 0
        for w in row:
            v = (v << 64) | int(w)
        out.append("".join("ACGT"[(v >> (2 * (k - 1 - j))) & 3] for j in range(k)))
    return out


# ---- the reference binary (oracle/_ref/kaarme), when it was built in this container -----------------

def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def run_ref(path, k, mode=0, slots=None, unique=None, fpr=None, min_abundance=2, threads=3, out=None,
            timeout=600):
    """Run the unmodified reference CLI; returns (sorted output bytes, stdout text)."""
    out = out or (str(path) + f".ref.k{k}.m{mode}.out")
    cmd = [REF_BIN, str(path), str(k), "-m", str(mode), "-a", str(min_abundance), "-t", str(threads),
           "-o", out]
    if unique is not None:
        cmd += ["-b", "-u", str(unique)]
        if fpr is not None:
            cmd += ["-f", str(fpr)]
    else:
        cmd += ["-s", str(slots)]
    if os.path.exists(out):
        os.remove(out)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference rc={p.returncode}: {' '.join(cmd)}\n{p.stdout.decode()[-2000:]}")
    lines = []
    if os.path.exists(out):
        with open(out, "rb") as f:
            lines = f.read().splitlines(keepends=True)
        os.remove(out)
    lines.sort()
    return b"".join(lines), p.stdout.decode()
