"""Extracts a handful of spectral tables (raw `lambda value` rows) from the reference's Source/Tables into a
single JSON fixture, so that tests / bench can run on the GPU box where /root/reference does not exist.

The .spd files are measured DATA tables (IOR: luxpop database; emission: LSPDD), not reference source code.
Run in the build container:  python tools/make_golden_tables.py
"""
import json
import os
import sys

REF = "/root/reference/Source/Tables"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pltfalcor_b200", "data", "spd_tables.json")

IOR = ["Au", "Ag", "Cu", "Al", "Cr"]
EMISSION = [
    "2475_INC_A19_domestic-use_softwhite-color_Rona-brand_100W_130V_2450K",
    "2499_LED_Encaster_Calculite_domestic-use_softwhite-color_Philips_brand_22W_1500lm_7inch_3000K",
]


def parse(path):
    """Same rules as PiecewiseLinearSpectrum::fromFile (Source/Falcor/Utils/Color/Spectrum.cpp:89-112)."""
    w, v = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            parts = line.split()
            try:
                a, b = float(parts[0]), float(parts[1])
            except (ValueError, IndexError):
                break
            w.append(a)
            v.append(b)
    return {"wavelengths": w, "values": v}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tables not available here")
    out = {"ior": {}, "emission": {}}
    for name in IOR:
        out["ior"][name] = {"eta": parse(f"{REF}/ior/{name}.eta.spd"), "k": parse(f"{REF}/ior/{name}.k.spd")}
    for name in EMISSION:
        out["emission"][name] = parse(f"{REF}/emission/{name}.spd")
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
