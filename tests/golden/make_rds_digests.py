"""Regenerates tests/golden/rds_digests.json: SHA-256 of the DECOMPRESSED serialisation of every `.rds` file the reference
ships, plus its header. Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_rds_digests.py
tests/test_rds_io.py re-encodes each file with `magmaclust_b200.rds_io` and compares against these digests when the
reference tree is present, and always checks the digest of its own encoding of the committed known-answer objects."""
import glob
import gzip
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from magmaclust_b200 import rds_io  # noqa: E402

REF = "/root/reference"


def main():
    out = {}
    for p in sorted(glob.glob(f"{REF}/**/*.rds", recursive=True)):
        raw = gzip.decompress(open(p, "rb").read())
        _, hdr = rds_io.decode(raw)
        out[os.path.relpath(p, REF)] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw), "header": hdr}
    path = os.path.join(os.path.dirname(__file__), "rds_digests.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out), "files")


if __name__ == "__main__":
    main()
