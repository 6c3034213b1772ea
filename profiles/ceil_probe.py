import importlib, sys
sys.path.insert(0, '/root/repo')
kg = importlib.import_module("canonical-k-mer-hash-table_b200")
for mb in (8, 32, 64, 128, 256, 1024, 8192):
    v = kg.atomic_ceiling(0, region_bytes=mb << 20, n_ops=1 << 29, reps=2)
    print(f"region {mb:6d} MiB: {v/1e9:8.2f} G RED/s")
