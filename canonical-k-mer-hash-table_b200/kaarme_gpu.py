"""ctypes binding of libkaarme_gpu.so (include/kaarme_gpu.h) -- the product path as seen from Python.

Used by tests/ (the parity tests), bench.py and __graft_entry__.smoke().  PyTorch is plumbing
only (device buffers, torch.distributed); every data-path step is a kernel of the in-tree library.
There is NO fallback: if the library is missing or no B200 is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkaarme_gpu.so")

ABI_VERSION = 2
INPUT_FASTA, INPUT_PLAIN = 0, 2
TABLE_PLAIN, TABLE_KAARME = 0, 2
PASS_BLOOM, PASS_COUNT = 1, 2
FEED_CONTEXT = 1
COUNT_EXACT, COUNT_REFERENCE = 0, 1
CFG_REFERENCE_BLOOM = 1

STATUS = {0: "KG_OK", 1: "KG_EBADARG", 2: "KG_ECUDA", 3: "KG_ETABLE_FULL", 4: "KG_ENCCL", 5: "KG_ENOMEM",
          6: "KG_ESINK"}


class KaarmeError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: {STATUS.get(status, status)} {detail}".strip())


class TableFull(KaarmeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("k", C.c_uint32), ("table_mode", C.c_int32),
                ("input_mode", C.c_int32), ("min_slots", C.c_uint64), ("use_bloom", C.c_int32),
                ("device", C.c_int32), ("fpr", C.c_double), ("expected_unique", C.c_uint64),
                ("batch_bytes", C.c_uint64), ("rank", C.c_int32), ("world", C.c_int32),
                ("partitions", C.c_uint32), ("reserved", C.c_uint32)]


class PassStats(C.Structure):
    _fields_ = [("input_kmers", C.c_uint64), ("inserted_kmers", C.c_uint64), ("distinct", C.c_uint64),
                ("table_slots", C.c_uint64), ("new_in_first", C.c_uint64), ("new_in_second", C.c_uint64),
                ("bloom_bits", C.c_uint64), ("bloom_hashes", C.c_uint32), ("partitions", C.c_uint32),
                ("raw_bytes", C.c_uint64), ("bases", C.c_uint64), ("device_ms", C.c_double),
                ("parse_ms", C.c_double), ("count_ms", C.c_double), ("exchange_ms", C.c_double),
                ("insert_ms", C.c_double), ("insert_launches", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class CompactStats(C.Structure):
    _fields_ = [("kmers", C.c_uint64), ("roots", C.c_uint64), ("bytes", C.c_uint64),
                ("reference_bytes", C.c_uint64), ("max_chain", C.c_uint64), ("device_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_size_t)
TEXT_SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_char), C.c_size_t, C.c_size_t)

EXPORTS = ["kg_abi_version", "kg_strerror", "kg_last_error", "kg_device_count", "kg_create", "kg_destroy",
           "kg_host_alloc", "kg_host_free", "kg_comm_unique_id", "kg_comm_init", "kg_pass_begin",
           "kg_stream_begin", "kg_feed", "kg_feed_device", "kg_pass_end", "kg_compact", "kg_export",
           "kg_table_info", "kg_atomic_ceiling", "kg_launch_count", "kg_kaarme_download", "kg_export_text",
           "kg_kaarme_upload", "kg_checksum"]

_lib = None


def lib():
    """Load the in-tree library; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KaarmeError(2, "libkaarme_gpu.so missing",
                              f"build it with `make -C {HERE} lib` (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.kg_strerror.restype = C.c_char_p
        L.kg_last_error.restype = C.c_char_p
        L.kg_last_error.argtypes = [C.c_void_p]
        L.kg_device_count.argtypes = [C.POINTER(C.c_int)]
        L.kg_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.kg_destroy.argtypes = [C.c_void_p]
        L.kg_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.kg_host_free.argtypes = [C.c_void_p]
        L.kg_comm_unique_id.argtypes = [C.c_void_p]
        L.kg_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.kg_pass_begin.argtypes = [C.c_void_p, C.c_int]
        L.kg_stream_begin.argtypes = [C.c_void_p, C.c_int]
        L.kg_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
        L.kg_feed_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
        L.kg_pass_end.argtypes = [C.c_void_p, C.POINTER(PassStats)]
        L.kg_compact.argtypes = [C.c_void_p, C.POINTER(CompactStats)]
        L.kg_export.argtypes = [C.c_void_p, C.c_uint64, C.c_int, SINK_FN, C.c_void_p]
        L.kg_table_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.kg_atomic_ceiling.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
        L.kg_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.kg_kaarme_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.kg_export_text.argtypes = [C.c_void_p, C.c_uint64, C.c_int, TEXT_SINK_FN, C.c_void_p]
        L.kg_kaarme_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.kg_checksum.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def comm_unique_id() -> bytes:
    """rank 0 calls this and broadcasts the 256 bytes (e.g. torch.distributed.broadcast_object_list)."""
    buf = C.create_string_buffer(256)
    rc = lib().kg_comm_unique_id(buf)
    if rc:
        raise KaarmeError(rc, "kg_comm_unique_id", lib().kg_last_error(None).decode())
    return buf.raw


def device_count() -> int:
    n = C.c_int(0)
    lib().kg_device_count(C.byref(n))
    return n.value


def atomic_ceiling(device=0, region_bytes=8 << 30, n_ops=1 << 30, reps=3) -> float:
    v = C.c_double(0)
    rc = lib().kg_atomic_ceiling(device, region_bytes, n_ops, reps, C.byref(v))
    if rc:
        raise KaarmeError(rc, "kg_atomic_ceiling", lib().kg_last_error(None).decode())
    return v.value


def _host_view(data):
    """-> (address, nbytes, keepalive) for bytes / bytearray / numpy / CPU torch tensors."""
    if isinstance(data, (bytes, bytearray, memoryview)):
        a = np.frombuffer(data, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        return a.ctypes.data, a.size, a
    if hasattr(data, "data_ptr"):  # torch tensor
        t = data.contiguous()
        return t.data_ptr(), t.numel() * t.element_size(), t
    raise TypeError(type(data))


class Counter:
    """One kg_ctx (one GPU / one hash shard)."""

    def __init__(self, k, table_mode=TABLE_PLAIN, input_mode=INPUT_FASTA, min_slots=0, use_bloom=False,
                 fpr=0.01, expected_unique=0, device=0, batch_bytes=0, rank=0, world=1, partitions=0,
                 reference_bloom=False):
        # reference_bloom: EXPERIMENTAL KG_CFG_REFERENCE_BLOOM -- the reference's own filter, bit for bit
        self.cfg = Config(ABI_VERSION, k, table_mode, input_mode, min_slots, int(bool(use_bloom)), device,
                          fpr, expected_unique, batch_bytes, rank, world, partitions,
                          CFG_REFERENCE_BLOOM if reference_bloom else 0)
        self.k, self.W = k, (k + 31) // 32
        self._h = C.c_void_p()
        rc = lib().kg_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            raise KaarmeError(rc, "kg_create", lib().kg_last_error(None).decode())

    # -- plumbing ---------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            detail = lib().kg_last_error(self._h).decode()
            raise (TableFull if rc == 3 else KaarmeError)(rc, what, detail)

    def close(self):
        if self._h:
            lib().kg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the ABI ------------------------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 256)
        self._check(lib().kg_comm_init(self._h, buf, rank, world), "kg_comm_init")

    def pass_begin(self, which):
        self._check(lib().kg_pass_begin(self._h, which), "kg_pass_begin")

    def stream_begin(self, starts_in_header=False):
        self._check(lib().kg_stream_begin(self._h, int(starts_in_header)), "kg_stream_begin")

    def feed(self, data, flags=0):
        if hasattr(data, "is_cuda") and data.is_cuda:
            return self.feed_device(data.data_ptr(), data.numel() * data.element_size(), flags)
        addr, n, keep = _host_view(data)
        self._check(lib().kg_feed(self._h, addr, n, flags), "kg_feed")
        del keep

    def feed_device(self, ptr, nbytes, flags=0):
        self._check(lib().kg_feed_device(self._h, ptr, nbytes, flags), "kg_feed_device")

    def pass_end(self) -> dict:
        st = PassStats()
        self._check(lib().kg_pass_end(self._h, C.byref(st)), "kg_pass_end")
        return st.as_dict()

    def checksum(self, min_abundance=1, count_mode=COUNT_EXACT):
        """-> (k-mers, sum of counts, sum of g(k-mer), sum of g(k-mer)*count): order-independent, additive over shards."""
        out = (C.c_uint64 * 4)()
        self._check(lib().kg_checksum(self._h, min_abundance, count_mode, out), "kg_checksum")
        return tuple(int(x) for x in out)

    def compact(self) -> dict:
        st = CompactStats()
        self._check(lib().kg_compact(self._h, C.byref(st)), "kg_compact")
        self._compact = st.as_dict()
        return self._compact

    def kaarme_download(self):
        """-> (slots [kmers] uint64, roots [roots, W] uint64) of the compact structure built by compact()."""
        st = self._compact
        slots = np.zeros(max(1, st["kmers"]), np.uint64)
        roots = np.zeros(max(1, st["roots"]) * self.W, np.uint64)
        self._check(lib().kg_kaarme_download(self._h, slots.ctypes.data, roots.ctypes.data), "kg_kaarme_download")
        return slots[:st["kmers"]], roots[:st["roots"] * self.W].reshape(-1, self.W)

    def kaarme_upload(self, slots, roots):
        """Load a compact structure (as returned by kaarme_download) into this (-m 2, not yet counted) context;
        export() / export_text() then decode it on the GPU."""
        slots = np.ascontiguousarray(slots, dtype=np.uint64).reshape(-1)
        roots = np.ascontiguousarray(roots, dtype=np.uint64).reshape(-1)
        assert roots.size % self.W == 0
        self._check(lib().kg_kaarme_upload(self._h, slots.ctypes.data, slots.size, roots.ctypes.data,
                                           roots.size // self.W), "kg_kaarme_upload")
        self._compact = {"kmers": int(slots.size), "roots": int(roots.size // self.W)}

    def export_text(self, min_abundance=1, count_mode=COUNT_EXACT):
        """-> (text bytes, records): the output lines '<KMER> <COUNT>\\n' formatted on the GPU (kg_format_text),
        in the unspecified order the device produced them."""
        parts, total = [], [0]

        def sink(user, text, nbytes, records):
            parts.append(C.string_at(text, nbytes))
            total[0] += records
            return 0

        cb = TEXT_SINK_FN(sink)
        self._check(lib().kg_export_text(self._h, min_abundance, count_mode, cb, None), "kg_export_text")
        return b"".join(parts), total[0]

    def export(self, min_abundance=1, count_mode=COUNT_EXACT, sort=True):
        """-> (keys [n, W] uint64, counts [n] uint32), sorted by key when sort=True."""
        W = self.W
        keys_parts, count_parts = [], []

        def sink(user, keys, counts, n):
            keys_parts.append(np.ctypeslib.as_array(keys, shape=(n * W,)).copy())
            count_parts.append(np.ctypeslib.as_array(counts, shape=(n,)).copy())
            return 0

        cb = SINK_FN(sink)
        self._check(lib().kg_export(self._h, min_abundance, count_mode, cb, None), "kg_export")
        if keys_parts:
            keys = np.concatenate(keys_parts).reshape(-1, W)
            counts = np.concatenate(count_parts)
        else:
            keys = np.zeros((0, W), np.uint64)
            counts = np.zeros((0,), np.uint32)
        if sort and len(counts):
            order = np.lexsort([keys[:, j] for j in range(W - 1, -1, -1)])
            keys, counts = keys[order], counts[order]
        return keys, counts

    def table_info(self):
        s, b, w = C.c_uint64(), C.c_uint32(), C.c_uint32()
        lib().kg_table_info(self._h, C.byref(s), C.byref(b), C.byref(w))
        return {"slots": s.value, "slot_bytes": b.value, "key_words": w.value}

    def launch_count(self) -> int:
        v = C.c_uint64()
        lib().kg_launch_count(self._h, C.byref(v))
        return v.value

    # -- whole-input convenience (what the CLI does per file) -------------------------------------------------
    def run_pass(self, which, data, starts_in_header=False):
        self.pass_begin(which)
        self.stream_begin(starts_in_header)
        self.feed(data)
        return self.pass_end()


def slice_context(data, lo, k, input_mode=INPUT_FASTA):
    """Host-side sharding of one input across ranks (the role of text_reader.h:141-184 in the reference).

    A rank that owns bytes [lo, hi) of the file must also see the k-1 bases before `lo` (fed with
    FEED_CONTEXT so they are not counted twice) and must know whether its first byte lies inside a FASTA
    header.  Returns (ctx_lo, starts_in_header): feed data[ctx_lo:lo] as context after
    stream_begin(starts_in_header), then data[lo:hi].
    """
    buf = memoryview(data) if not isinstance(data, memoryview) else data
    need = k - 1
    i = lo
    while i > 0 and need > 0:       # walk back over k-1 non-newline bytes (newlines inside records are skipped)
        i -= 1
        if buf[i] != 10:
            need -= 1
    ctx_lo = i
    in_header = False
    if input_mode == INPUT_FASTA:   # header state at ctx_lo: a '>' since the last newline (parallel_parser.hpp:602-620)
        j = ctx_lo
        while j > 0:
            j -= 1
            if buf[j] == 10:
                break
            if buf[j] == 62:
                in_header = True
                break
    return ctx_lo, in_header


def shard_ranges(nbytes, world):
    """Even byte ranges [lo, hi) per rank; any cut is legal because slice_context repairs the boundary."""
    return [(nbytes * r // world, nbytes * (r + 1) // world) for r in range(world)]


def keys_to_text(keys, counts, k) -> bytes:
    """Format (keys, counts) as the reference writer does: '<KMER> <COUNT>\\n' (kmer_hash_table.cpp:2022-2043)."""
    keys = np.asarray(keys, dtype=np.uint64)
    n, W = keys.shape
    if n == 0:
        return b""
    # unpack 2-bit characters, most significant first
    shifts = np.arange(62, -2, -2, dtype=np.uint64)
    codes = ((keys[:, :, None] >> shifts[None, None, :]) & np.uint64(3)).astype(np.uint8).reshape(n, W * 32)
    codes = codes[:, W * 32 - k:]
    chars = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    lines = [chars[i].tobytes() + b" " + str(int(counts[i])).encode() + b"\n" for i in range(n)]
    return b"".join(lines)


# ---- mirrors of the reference functors (parallel_parser.hpp), file in -> file out -----------------------------
def _sniff(path):
    """main.cpp:27-68: format by extension + first byte. Returns input_mode or raises like the reference exits."""
    ext = os.path.splitext(path)[1]
    with open(path, "rb") as f:
        first = f.read(1)
    if ext in (".fasta", ".fa"):
        if first != b">":
            raise ValueError(f"Input file {path} is ill-formed")
        return INPUT_FASTA
    if ext in (".fastq", ".fq"):
        raise NotImplementedError("Not implemented yet")  # parallel_parser.hpp:797-800
    if first not in b"actgACGT" or first == b"":
        raise ValueError(f"Input file {path} is ill-formed")
    return INPUT_PLAIN


def _run_file(input_file, output_file, k, table_mode, min_abundance, min_slots=0, expected_unique=0, fpr=0.01,
              use_bloom=False, device=0, batch_bytes=0):
    input_mode = _sniff(input_file)
    with open(input_file, "rb") as f:
        data = f.read()
    with Counter(k, table_mode, input_mode, min_slots, use_bloom, fpr, expected_unique, device, batch_bytes) as c:
        stats = {}
        if use_bloom:
            stats["bloom"] = c.run_pass(PASS_BLOOM, data)
        stats["count"] = c.run_pass(PASS_COUNT, data)
        if table_mode == TABLE_KAARME:
            stats["compact"] = c.compact()
        text, stats["written"] = c.export_text(min_abundance, COUNT_REFERENCE)
    with open(output_file, "wb") as f:
        f.write(text)
    return stats


def parse_input_atomic_flag(input_file, output_file, k, min_slots, min_abundance, **kw):
    """parallel_parser.hpp:223-871 (-m 0, no Bloom)"""
    return _run_file(input_file, output_file, k, TABLE_PLAIN, min_abundance, min_slots=min_slots, **kw)


def parse_input_atomic_flag_BF(input_file, output_file, k, expected_unique, fpr, min_abundance, **kw):
    """main.cpp:395-480 + parallel_parser.hpp:2678-2974 (pass 1) + :1573-2242 (pass 2), -m 0 -b"""
    return _run_file(input_file, output_file, k, TABLE_PLAIN, min_abundance, expected_unique=expected_unique,
                     fpr=fpr, use_bloom=True, **kw)


def parse_input_pointer_atomic_variable(input_file, output_file, k, min_slots, min_abundance, **kw):
    """parallel_parser.hpp:1175-1565 (-m 2, no Bloom)"""
    return _run_file(input_file, output_file, k, TABLE_KAARME, min_abundance, min_slots=min_slots, **kw)


def parse_input_pointer_atomic_variable_BF(input_file, output_file, k, expected_unique, fpr, min_abundance, **kw):
    """parallel_parser.hpp:2249-2672 (-m 2 -b)"""
    return _run_file(input_file, output_file, k, TABLE_KAARME, min_abundance, expected_unique=expected_unique,
                     fpr=fpr, use_bloom=True, **kw)
