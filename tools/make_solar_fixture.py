"""Generates tests/golden/solar_178.csv: the input of BASELINE config 1 (small solar-system CSV, 177 bodies + Sun).

Deterministic Python restatement of the reference's offline `dataset_converter` (Keplerian elements -> Cartesian state
vectors; reference dataset_converter/src/conversion.cpp:13-128, io.cpp:38-155) applied to the reference's own data file
dataset_converter/data/planets_and_moons.csv.  Run in the build container only (it reads /root/reference); the CSV it
writes is committed so nothing on the GPU box needs the reference tree.  Every body in that file carries a mass, so the
converter's random albedo-based mass estimate (approximation.cpp) is never exercised.
"""
import csv
import math
import os
import sys

SRC = "/root/reference/dataset_converter/data/planets_and_moons.csv"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "solar_178.csv")

G = (86400.0 * 86400.0) * (6.67428e-11 / (1.49597870691e11 * 1.49597870691e11 * 1.49597870691e11))  # constants.hpp
EPOCH = 2451544.5  # midnight January 1, 2000 (conversion.cpp:98)


def num(s):
    return float(s) if s.strip() else 0.0


def to_state(k, c):
    mu = G * c["mass"]
    M_t = k["ma"] + (EPOCH - k["epoch"]) * math.sqrt(mu / k["a"] ** 3)
    E = M_t
    for _ in range(30):  # Newton-Raphson on Kepler's equation
        E = E - (E - k["e"] * math.sin(E) - M_t) / (1.0 - k["e"] * math.cos(E))
    v_t = 2.0 * math.atan2(math.sqrt(1.0 + k["e"]) * math.sin(E / 2.0), math.sqrt(1.0 - k["e"]) * math.cos(E / 2.0))
    r_c = k["a"] * (1 - k["e"] * math.cos(E))
    ox, oy = r_c * math.cos(v_t), r_c * math.sin(v_t)
    s = math.sqrt(mu * k["a"]) / r_c
    ovx, ovy = s * -math.sin(E), s * (math.sqrt(1 - k["e"] ** 2) * math.cos(E))
    cw, sw, co, so, ci, si = (math.cos(k["w"]), math.sin(k["w"]), math.cos(k["om"]), math.sin(k["om"]),
                              math.cos(k["i"]), math.sin(k["i"]))

    def rot(a, b):
        return (a * (cw * co - sw * ci * so) - b * (sw * co + cw * ci * so),
                a * (cw * so + sw * ci * co) + b * (cw * ci * co - sw * so),
                a * (sw * si) + b * (cw * si))
    x, y, z = rot(ox, oy)
    vx, vy, vz = rot(ovx, ovy)
    return dict(name=k["name"], cls=k["cls"], mass=k["mass"], x=x + c["x"], y=y + c["y"], z=z + c["z"],
                vx=vx + c["vx"], vy=vy + c["vy"], vz=vz + c["vz"])


def main():
    rows = []
    with open(SRC) as f:
        for r in csv.DictReader(f):
            rows.append(dict(a=num(r["a"]), e=num(r["e"]), w=math.radians(num(r["w"])), om=math.radians(num(r["om"])),
                             i=math.radians(num(r["i"])), ma=math.radians(num(r["ma"])), epoch=num(r["epoch"]),
                             mass=num(r["mass"]), name=r["name"], cls=r["class"], central=r["central_body"]))
    rows.sort(key=lambda k: k["name"])  # io.cpp:158-172 sorts by name before conversion
    bodies = [dict(name="Sun", cls="STA", mass=1.98847e30, x=0.0, y=0.0, z=0.0, vx=0.0, vy=0.0, vz=0.0)]
    for k in rows:
        if k["central"] == "Sun":
            bodies.append(to_state(k, bodies[0]))
    for k in rows:
        if k["central"] != "Sun":
            c = next(b for b in bodies if b["name"] == k["central"])
            bodies.append(to_state(k, c))
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    with open(DST, "w") as out:
        out.write("id,name,class,mass,pos_x,pos_y,pos_z,vel_x,vel_y,vel_z\n")
        for i, b in enumerate(bodies):  # default ostream precision = %g (io.cpp:146-152)
            out.write("%d,%s,%s,%s\n" % (i, b["name"], b["cls"], ",".join("%g" % b[k] for k in ("mass", "x", "y", "z", "vx", "vy", "vz"))))
    print("wrote", DST, len(bodies), "bodies")


if __name__ == "__main__":
    sys.exit(main())
