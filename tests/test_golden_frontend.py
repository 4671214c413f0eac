"""The oracle restatements of everything beside the extractor against known answers frozen from the reference's own code
(tests/golden/frontend_hashes.json, made by tests/golden/make_golden_frontend.py where /root/reference exists).  Unlike
tests/test_oracle_*_vs_ref.py this needs no reference binary, so it also runs where oracle/_ref was not shipped."""
import json
import os

import pytest

from golden_frontend import cases, digest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "frontend_hashes.json")
_CASES = None


def _cases():
    global _CASES
    if _CASES is None:
        _CASES = cases()
    return _CASES


with open(GOLDEN) as _f:
    _GOLD = json.load(_f)


def test_golden_file_covers_every_case():
    assert set(_GOLD) == set(_cases())


@pytest.mark.parametrize("name", sorted(_GOLD))
def test_oracle_reproduces_the_reference_answers(name):
    if name in ("search_by_projection_kf",):
        from oracle import pyoracle as po
        if not os.path.exists(po.MATCH_REF_SO):
            pytest.skip("this case derives its query arrays with helpers of the reference binary")
    arrays = _cases()[name][1]()
    assert [list(a.shape) for a in arrays] == _GOLD[name]["shapes"], name
    assert digest(*arrays) == _GOLD[name]["sha256"], name
