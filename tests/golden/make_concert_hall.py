"""Writes tests/golden/concert_hall.obj: the GEOMETRY of the reference's demo concert hall
(/root/reference/demo/evaluation/models/object/concert.obj, BASELINE config 5) as a fixture,
because /root/reference does not exist on the GPU box. Only the `v`, `f`, `g` and `usemtl`
records are kept (texture coordinates and normals are dropped, face corners reduced to their
vertex index); coordinates are copied as written. Run here, in the build container:

    python tests/golden/make_concert_hall.py
"""
import os

SRC = "/root/reference/demo/evaluation/models/object/concert.obj"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "concert_hall.obj")


def main():
    out = ["# geometry of the wayverb demo concert hall (see make_concert_hall.py); units: metres"]
    for line in open(SRC):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            out.append("v " + " ".join(p[1:4]))
        elif p[0] == "f":
            out.append("f " + " ".join(c.split("/")[0] for c in p[1:]))
        elif p[0] in ("g", "usemtl"):
            out.append(" ".join(p))
    with open(DST, "w") as f:
        f.write("\n".join(out) + "\n")
    print(DST, sum(1 for l in out if l.startswith("f ")), "faces")


if __name__ == "__main__":
    main()
