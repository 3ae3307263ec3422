#!/usr/bin/env python
"""Derive the lookup tables used by both the oracle and the CUDA library.

Round-1 provenance: the three tables are *data* (tomohxx shanten tables, 山岡 agari
table) that libriichi ships gzipped under libriichi/src/algo/data/. This script only
gunzips them into mortal_b200/data/ (git-ignored, travels to the GPU box like a built
.so). It runs in the dev container where /root/reference exists; on the GPU box the
prebuilt files are used. Formats: SURVEY.md Appendix A.
"""
import gzip
import os
import sys

REF = os.environ.get("MORTAL_REF_DATA", "/root/reference/libriichi/src/algo/data")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mortal_b200", "data")
FILES = {
    "shanten_suhai.bin": ("shanten_suhai.bin.gz", 9_703_885),
    "shanten_jihai.bin": ("shanten_jihai.bin.gz", 390_160),
    "agari.bin": ("agari.bin.gz", 86_058),
}


def main() -> int:
    os.makedirs(OUT, exist_ok=True)
    for out_name, (src, size) in FILES.items():
        dst = os.path.join(OUT, out_name)
        if os.path.exists(dst) and os.path.getsize(dst) == size:
            continue
        src_path = os.path.join(REF, src)
        if not os.path.exists(src_path):
            print(f"build_tables: {src_path} missing and {dst} not prebuilt", file=sys.stderr)
            return 1
        with gzip.open(src_path, "rb") as f:
            raw = f.read()
        assert len(raw) == size, (out_name, len(raw), size)
        with open(dst, "wb") as f:
            f.write(raw)
        print(f"build_tables: wrote {dst} ({len(raw)} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
