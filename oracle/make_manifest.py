"""Regenerates oracle/state_manifest.json from the UNMODIFIED reference (needs /root/reference; run here only)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import shims  # noqa: E402

if __name__ == "__main__":
    shims.install()
    from iggt.models.vggt import IGGT
    sd = IGGT().state_dict()
    man = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_manifest.json")
    json.dump(man, open(out, "w"))
    print(len(man), "entries,", sum(v.numel() for v in sd.values()), "elements ->", out)
