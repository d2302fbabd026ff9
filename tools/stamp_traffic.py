#!/usr/bin/env python
"""Stamp profiles/update_kernel_traffic.json with the sha256 of the kernel sources the ncu capture was taken from.
bench.py reports roofline.traffic only while that stamp matches the tree (otherwise null: the capture is stale).

    python tools/stamp_traffic.py <dram_read_bytes_per_launch> <dram_write_bytes_per_launch> <source-note> [--no-skip]
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["simplexoncuda_b200/csrc/b2s_kernels.cuh", "simplexoncuda_b200/csrc/b2s_device.cuh",
           "simplexoncuda_b200/csrc/b2s_lookahead.cuh"]


def main():
    rd, wr, note = float(sys.argv[1]), float(sys.argv[2]), sys.argv[3]
    skip = "--no-skip" not in sys.argv
    srcs = [f for f in SOURCES if os.path.exists(os.path.join(ROOT, f))]
    h = hashlib.sha256()
    for f in srcs:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    out = {"kernel": "dominant update kernel of bench.py's default run", "source": note, "vars_constraints": [8192, 8192],
           "skip_zero_rows": skip, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
           "dram_bytes_per_launch": rd + wr, "kernel_sources": srcs, "kernel_sources_sha256": h.hexdigest()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
