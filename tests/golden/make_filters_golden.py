"""Generates tests/golden/filters_golden.json from the reference itself (oracle/_ref/libhbref.so, the reference's own
sources compiled unmodified by oracle/Makefile).  The reference tree holds no fixtures for these filters.
Run in the build container:   python tests/golden/make_filters_golden.py"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(Path(__file__).parent))
from handbrake_b200.hblib import FilterLib  # noqa: E402
from filter_cases import FMT, cases, digest  # noqa: E402


def main():
    ref = FilterLib(REPO / "oracle" / "_ref" / "libhbref.so")
    out = {}
    for name, c in cases().items():
        r = ref.run(c["ref"], c["settings"], c["clip"], FMT[c["depth"]], c["w"], c["h"], flags=c["flags"], combed=c["combed"])
        assert r.saw_eof and not r.init_failed
        out[name] = dict(ref=c["ref"], settings=c["settings"], depth=c["depth"], w=c["w"], h=c["h"], frames_in=int(c["clip"].shape[0]), **digest(r))
        print(name, len(out[name]["sha256"]), out[name]["sha256"][0][:16])
    (Path(__file__).parent / "filters_golden.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
