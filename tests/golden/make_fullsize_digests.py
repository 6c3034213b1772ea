#!/usr/bin/env python3
"""Mints tests/golden/fullsize_digests.json: the UNMODIFIED reference binary (oracle/_ref/kaarme, built by
oracle/Makefile from /root/reference) run on the BASELINE.json configurations at full size (C5 at the 1/10 scale
SURVEY.md section 8d allows), its output reduced to an order-independent digest (tests/native/linedigest.c).

Runs in the build container on the host CPU (no GPU involved; the reference has no GPU code).  Resumable: cases already
in the JSON are skipped unless named on the command line.  Takes about an hour on 8 cores.

    python tests/golden/make_fullsize_digests.py [CASE ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fullsize_util as fu  # noqa: E402


def main():
    want = sys.argv[1:]
    done = json.load(open(fu.GOLDEN)) if os.path.exists(fu.GOLDEN) else {}
    done.setdefault("_inputs", fu.INPUTS)
    done.setdefault("_about", "digests of the unmodified reference binary's output; see tests/golden/make_fullsize_digests.py")
    threads = max(3, min(64, os.cpu_count() or 3))
    last_input = None
    for name, (inp, k, args) in fu.CASES.items():
        if (want and name not in want) or (not want and name in done):
            continue
        if last_input and last_input != inp:
            fu.remove_input(last_input)
        last_input = inp
        path = fu.generate(inp)
        dig, log, wall = fu.run_digest(fu.REF, path, k, args, threads)
        rec = {"input": inp, "k": k, "args": args, "digest": dig, "reference_threads": threads,
               "reference_wall_s": round(wall, 2), "reference_timers_s": fu.timers(log),
               "table_slots": fu.log_value(log, "Hash table size is:"),
               "input_bytes": os.path.getsize(path)}
        for tag, key in (("New k-mers in first bloom filter", "new_in_first"), ("New k-mers in second bloom filter", "new_in_second"),
                         ("Main array slots used", "main_slots_used"), ("Max secondary array slots used", "max_secondary_used")):
            v = fu.log_value(log, tag)
            if v is not None:
                rec[key] = v
        done[name] = rec
        with open(fu.GOLDEN + ".tmp", "w") as f:
            json.dump(done, f, indent=1, sort_keys=True)
        os.replace(fu.GOLDEN + ".tmp", fu.GOLDEN)
        print(name, json.dumps(rec), flush=True)
    if last_input:
        fu.remove_input(last_input)


if __name__ == "__main__":
    main()
