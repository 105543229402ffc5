"""The committed digests of the reference's output for every filter on the path (tests/golden/filters_golden.json, made
by tests/golden/make_filters_golden.py from the compiled reference) still describe what the reference produces here."""
import json
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).parent / "golden"))
from filter_cases import FMT, cases, digest  # noqa: E402

GOLDEN = json.loads((Path(__file__).parent / "golden" / "filters_golden.json").read_text())
CASES = cases()


def test_golden_file_covers_every_case():
    assert sorted(GOLDEN) == sorted(CASES)
    for name, g in GOLDEN.items():
        assert len(g["sha256"]) == len(g["start"]) == len(g["combed"]) > 0


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_matches_golden(ref, name):
    c, g = CASES[name], GOLDEN[name]
    r = ref.run(c["ref"], c["settings"], c["clip"], FMT[c["depth"]], c["w"], c["h"], flags=c["flags"], combed=c["combed"])
    d = digest(r)
    assert d["start"] == g["start"] and d["combed"] == g["combed"] and d["sha256"] == g["sha256"]


@pytest.mark.parametrize("name", sorted(n for n in CASES if n.startswith("detelecine")))
def test_detelecine_hostlogic_matches_golden(name):
    """the product's pullup state machine over the plain-C metric restatement (see tests/test_detelecine.py)"""
    from handbrake_b200.hblib import FilterLib
    from test_detelecine import HOSTLOGIC_SO
    if not HOSTLOGIC_SO.exists():
        pytest.skip("oracle/_ref/libhostlogic.so not built")
    c, g = CASES[name], GOLDEN[name]
    r = FilterLib(HOSTLOGIC_SO).run(c["cuda"], c["settings"], c["clip"], FMT[c["depth"]], c["w"], c["h"], flags=c["flags"])
    d = digest(r)
    assert d["start"] == g["start"] and d["sha256"] == g["sha256"]
