"""Re-derives the state_dict layout from the UNMODIFIED reference (needs /root/reference; run here only) and records its
digest in oracle/state_manifest.sha256.json.  The list itself (names / shapes / dtypes, 2053 entries) is kept ONCE, as
the product's schema iggt_official_b200/state_layout.json; tests/test_layout.py holds that file to the digest, so the
product's layout is pinned to the reference without a second 150 KB copy of the list.  `--write-layout` also rewrites
the product's file from the reference (only needed if the reference's modules change)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402


def canonical(man):
    return json.dumps(man, separators=(",", ":")).encode()


if __name__ == "__main__":
    shims.install()
    from iggt.models.vggt import IGGT
    sd = IGGT().state_dict()
    man = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
    rec = {"entries": len(man), "elements": sum(v.numel() for v in sd.values()),
           "sha256_canonical_json": hashlib.sha256(canonical(man)).hexdigest(),
           "what": "names / shapes / dtypes of the unmodified reference's IGGT().state_dict(), written by oracle/make_manifest.py; "
                   "the list itself is iggt_official_b200/state_layout.json (one copy), held to this digest by tests/test_layout.py"}
    json.dump(rec, open(os.path.join(ROOT, "oracle", "state_manifest.sha256.json"), "w"), indent=1)
    layout = os.path.join(ROOT, "iggt_official_b200", "state_layout.json")
    if "--write-layout" in sys.argv:
        json.dump(man, open(layout, "w"))
    same = json.load(open(layout)) == man
    print(rec["entries"], "entries,", rec["elements"], "elements; product layout file matches the reference:", same)
