"""Extracts the CIE 1931 standard-observer table (measured DATA the reference's SpectrumUtils unit test compares its
analytic fit against, Tests/Utils/Color/SpectrumUtilsTests.cpp:40-80) into tests/golden/cie1931_1nm.json.
Run in the build container:  python tools/make_golden_cie.py"""
import json
import os
import re
import sys

SRC = "/root/reference/external/include/xyzcurves/ciexyzCurves1931_1nm.h"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cie1931_1nm.json")


def main():
    if not os.path.exists(SRC):
        sys.exit("reference tables not available here")
    rows = []
    for line in open(SRC):
        m = re.match(r"\s*\{\s*([-0-9.e]+)f,\s*([-0-9.e]+)f,\s*([-0-9.e]+)f\s*\}", line)
        if m:
            rows.append([float(m.group(i)) for i in (1, 2, 3)])
    assert len(rows) == 471
    with open(OUT, "w") as f:
        json.dump({"start_nm": 360.0, "end_nm": 830.0, "xyz": rows}, f, separators=(",", ":"))
    print("wrote", OUT, len(rows))


if __name__ == "__main__":
    main()
