"""Generates tests/golden/extract_hashes.json from oracle/_ref (the unmodified reference ORBextractor.cc).
Run in the build container, where /root/reference exists:  python tests/golden/make_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from oracle import pyoracle as po  # noqa: E402
from test_oracle_vs_ref import _digest, golden_cases  # noqa: E402

out = {}
for name, (img, kw) in golden_cases().items():
    k, d = po.RefExtractor(**kw).extract(img)
    out[name] = {"n": int(len(k)), "sha256": _digest(k, d), "shape": list(img.shape), **kw}
with open(os.path.join(ROOT, "tests", "golden", "extract_hashes.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps({k: v["n"] for k, v in out.items()}))
