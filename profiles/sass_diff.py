#!/usr/bin/env python3
"""Which kernels of two builds of libkaarme_gpu.so differ?  (per-function hash of `cuobjdump -sass`, addresses and encodings
stripped).  Used to show what changed after the last build that ran on hardware.
    python profiles/sass_diff.py OLD.so NEW.so"""
import hashlib
import re
import subprocess
import sys


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return {k: hashlib.sha1("\n".join(v).encode()).hexdigest() for k, v in funcs.items()}


a, b = sass(sys.argv[1]), sass(sys.argv[2])
diff = [k for k in sorted(set(a) | set(b)) if a.get(k) != b.get(k)]
print(f"{len(a)} / {len(b)} functions; differing: {len(diff)}")
for k in diff:
    print("  ", k)
