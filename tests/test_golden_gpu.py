"""GPU parity against the COMMITTED digests of the reference output (tests/golden/filters_golden.json): every CUDA filter
object on the cases of tests/golden/filter_cases.py -- no reference build needed at run time."""
import json
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).parent / "golden"))
from filter_cases import FMT, cases, digest  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "filters_golden.json").read_text())
CASES = cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(cuda_filters, name):
    c, g = CASES[name], GOLDEN[name]
    r = cuda_filters.run(c["cuda"], c["settings"], c["clip"], FMT[c["depth"]], c["w"], c["h"], flags=c["flags"], combed=c["combed"])
    assert not r.init_failed and r.saw_eof
    d = digest(r)
    assert d["start"] == g["start"], name
    assert d["combed"] == g["combed"], name
    bad = [i for i, (a, b) in enumerate(zip(d["sha256"], g["sha256"])) if a != b]
    assert len(d["sha256"]) == len(g["sha256"]) and not bad, f"{name}: frames {bad[:8]} differ from the reference's digests"
    assert cuda_filters.buffers_alive() == 0
