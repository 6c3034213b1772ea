"""kaarme-b200: B200-native canonical k-mer counting path (drop-in for the reference's counting functors).

Layout: csrc/ (sm_100a kernels + C ABI, built in-tree as libkaarme_gpu.so), host/ (the `kaarme` CLI, C++),
kaarme_gpu.py (ctypes binding used by tests and bench).  Import with
    importlib.import_module("canonical-k-mer-hash-table_b200")
"""
from .kaarme_gpu import *  # noqa: F401,F403
from .kaarme_gpu import (Counter, KaarmeError, TableFull, lib, device_count, atomic_ceiling, keys_to_text, comm_unique_id,
                         parse_input_atomic_flag, parse_input_atomic_flag_BF, parse_input_pointer_atomic_variable,
                         parse_input_pointer_atomic_variable_BF, EXPORTS, LIB_PATH)
