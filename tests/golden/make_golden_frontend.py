"""Generates tests/golden/frontend_hashes.json from the reference's own code run in the build container (oracle/_ref:
unmodified ORBmatcher.cc, MapPoint.cc, DBoW2, and the ComputeStereo* functions of Frame.cc):
    python tests/golden/make_golden_frontend.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from golden_frontend import cases, digest  # noqa: E402

out = {}
for name, (ref, _) in cases().items():
    arrays = ref()
    out[name] = {"sha256": digest(*arrays), "shapes": [list(a.shape) for a in arrays]}
    print(name, out[name]["shapes"])
with open(os.path.join(ROOT, "tests", "golden", "frontend_hashes.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
