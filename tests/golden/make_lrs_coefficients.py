"""Extracts the checked-in LRS coefficient sets of the reference into a small
fixture. Run in the build container (needs /root/reference); the GPU box only
reads the committed JSON.

source: /root/reference/bin/boundary_test/output.soft/coefficients.txt
        (cereal JSON written by bin/boundary_test/boundary_test.cpp; nine
        entries = 3 materials x 3 angles, the filters do not depend on angle)
"""
import json
import os

SRC = "/root/reference/bin/boundary_test/output.soft/coefficients.txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lrs_coefficients.json")


def vec(d):
    return [d["value%d" % i] for i in range(7)]


def main():
    raw = json.load(open(SRC))
    out = {"source": "bin/boundary_test/output.soft/coefficients.txt", "sets": []}
    for k in sorted(raw):
        e = raw[k]
        out["sets"].append({
            "test": e["test"], "material": e["material"],
            "reflectance": {"b": vec(e["reflectance"]["b"]), "a": vec(e["reflectance"]["a"])},
            "impedance": {"b": vec(e["impedance"]["b"]), "a": vec(e["impedance"]["a"])},
        })
    json.dump(out, open(DST, "w"), indent=1)
    print("wrote", DST, len(out["sets"]), "sets")


if __name__ == "__main__":
    main()
