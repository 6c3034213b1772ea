"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/kaarme_gpu.h declares; argument validation and the no-fallback rule (no compute without a GPU)."""
import ctypes as C
import importlib
import os
import re

import pytest

from conftest import ROOT

kg = importlib.import_module("canonical-k-mer-hash-table_b200")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(kg.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "kaarme_gpu.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(kg_[a-z_0-9]+)\s*\(", hdr, flags=re.M))
    assert declared == set(kg.EXPORTS), declared ^ set(kg.EXPORTS)
    L = C.CDLL(kg.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.kg_abi_version() == kg.kaarme_gpu.ABI_VERSION


def test_struct_layout_matches_header():
    # sizes as laid out by the C compiler for include/kaarme_gpu.h (checked with a tiny C program)
    import subprocess, tempfile
    src = '#include "kaarme_gpu.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(kg_config), sizeof(kg_pass_stats), sizeof(kg_compact_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
        out = subprocess.run([os.path.join(d, "s")], stdout=subprocess.PIPE, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(kg.kaarme_gpu.Config), C.sizeof(kg.kaarme_gpu.PassStats),
                                     C.sizeof(kg.kaarme_gpu.CompactStats)]


def test_bad_arguments_rejected():
    for kw in (dict(k=0, min_slots=10), dict(k=257, min_slots=10), dict(k=21, min_slots=0),
               dict(k=21, min_slots=10, table_mode=1), dict(k=21, use_bloom=True, expected_unique=0),
               dict(k=21, min_slots=10, rank=2, world=2)):
        with pytest.raises(kg.KaarmeError) as e:
            kg.Counter(**kw)
        assert e.value.status == 1


def test_no_cpu_fallback():
    """Without a B200 the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert kg.device_count() == 0
    with pytest.raises(kg.KaarmeError) as e:
        kg.Counter(k=21, min_slots=1000)
    assert e.value.status == 2


def test_product_does_not_touch_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, link or execute it."""
    pkg = os.path.join(ROOT, "canonical-k-mer-hash-table_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)
