"""End-to-end through the reference-facing executable (same flags as the reference's main.cpp): CSV in, ParaView /
lastState.csv / times.json out, compared with the oracle's restatement of startSimulation."""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "n-body-simulation_b200", "N_Body_Simulation")


def load_csv(path):
    rows = [l.rstrip("\n").split(",") for l in open(path)][1:]
    cols = list(zip(*rows))
    f = lambda k: np.array(cols[k], dtype=np.float64)
    return cols[1], cols[2], f(3), f(4), f(5), f(6), f(7), f(8), f(9)


def run_exe(tmp_path, fixture, *flags):
    out = tmp_path / "out"
    cmd = [EXE, "--file=" + fixture, "--vs_dir=" + str(out)] + list(flags)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    dirs = glob.glob(str(out / "*"))
    assert len(dirs) == 1
    return dirs[0], r.stdout


def read_last_state(d):
    lines = open(os.path.join(d, "lastState.csv")).read().splitlines()
    assert lines[0] == "position_x, position_y, position_z "
    return np.array([[float(v) for v in l.split(",")] for l in lines[1:]])


def fmt6(a):
    return np.array([float("%g" % v) for v in a])


@pytest.mark.parametrize("algorithm,extra", [("naive", ["--block_size=128", "--opt_stage=2"]),
                                             ("BarnesHut", ["--theta=0.4", "--wg_size_barnes_hut=64"])])
def test_short_run_matches_oracle(nb, oracle, tmp_path, golden_dir, algorithm, extra):
    fixture = os.path.join(golden_dir, "solar_178.csv")
    names, classes, m, x, y, z, vx, vy, vz = load_csv(fixture)
    d, stdout = run_exe(tmp_path, fixture, "--dt=1h", "--t_end=5d", "--vs=1d", "--algorithm=" + algorithm,
                        "--energy=true", *extra)
    theta = 0.4 if algorithm == "BarnesHut" else 1.05
    ref = oracle.simulate(algorithm, m, x, y, z, vx, vy, vz, 1.0 / 24, 5.0, 1.0, theta=theta, energy=True)
    assert ref["n_snap"] == 6
    files = sorted(os.listdir(d))
    assert "simulation.pvd" in files and "times.json" in files and "lastState.csv" in files
    assert len([f for f in files if f.endswith(".vtp")]) == 6
    assert "Finished step 5" in stdout
    last = read_last_state(d)
    assert last.shape == (178, 3)
    # printed with 6 significant digits; allow one unit in the last printed digit
    for k, key in enumerate(("px", "py", "pz")):
        want = ref[key][5]
        assert np.all(np.abs(last[:, k] - want) <= 1.01e-5 * np.maximum(np.abs(want), 1e-30) + 1e-12)
    # times.json: same keys as the reference
    t = json.load(open(os.path.join(d, "times.json")))
    assert t["body count"] == 178 and "B200" in t["device"]
    assert len(t["Acceleration Kernel Time"]) == 1 + 120 and len(t["Leapfrog Part 1"]) == 120
    if algorithm == "BarnesHut":
        assert t["algorithm"] == "Barnes-Hut Algorithm" and t["theta"] == 0.4
        for key in ("Octree creation", "AABB creation", "Compute center of mass", "Total Time"):
            assert len(t[key]) == 121
    else:
        assert t["algorithm"] == "Naive Algorithm" and t["block size"] == 128 and t["optimization stage"] == 2
    # one vtp: structure and the energy field data
    vtp = open(os.path.join(d, "simulation_step5.vtp")).read()
    assert '<Piece NumberOfPoints="178" NumberOfVerts="178">' in vtp
    for name in ("position", "body_id", "velocity", "acceleration", "mass", "name", "orbit_class", "offsets",
                 "connectivity", "kinetic energy", "potential energy", "total energy", "virial equilibrium"):
        assert 'Name="%s"' % name in vtp
    ek = float(vtp.split('Name="kinetic energy"')[1].split("\n")[1])
    assert ek == pytest.approx(ref["energy"][5][0], rel=1e-5)
    # Sun is a star (15), Earth a planet (17): orbit_class block
    oc = vtp.split('Name="orbit_class"')[1].split("</DataArray>")[0].split("\n")[1:179]
    assert oc[0] == "15" and oc[2] == "17"


def test_config1_full_year_naive(nb, oracle, tmp_path, golden_dir):
    """BASELINE config 1: naive opt_stage 2, solar-system CSV, dt = 1h, t_end = 365d (8760 steps), vs = 1d."""
    fixture = os.path.join(golden_dir, "solar_178.csv")
    names, classes, m, x, y, z, vx, vy, vz = load_csv(fixture)
    d, _ = run_exe(tmp_path, fixture, "--dt=1h", "--t_end=365d", "--vs=1d", "--algorithm=naive", "--opt_stage=2")
    ref = oracle.simulate("naive", m, x, y, z, vx, vy, vz, 1.0 / 24, 365.0, 1.0)
    assert ref["n_steps"] == 8760 and ref["n_snap"] == 366
    last = read_last_state(d)
    sel = np.array([c in ("STA", "PLA", "DWA") for c in classes])
    assert sel.sum() >= 10
    for k, key in enumerate(("px", "py", "pz")):
        want = ref[key][365][sel]
        got = last[sel, k]
        # equal at printed precision (6 significant digits, one unit in the last place of slack)
        assert np.all(np.abs(got - want) <= 1.01e-5 * np.abs(want) + 1e-9)
    assert len([f for f in os.listdir(d) if f.endswith(".vtp")]) == 366


# ---- ParaView writer: byte compatibility with the reference's format ----------------------------------------------------
ORBIT_CLASSES = ["AMO", "APO", "ATE", "IEO", "MCA", "IMB", "MBA", "OMB", "CEN", "TJN", "TNO", "AST", "PAA", "HYA", "STA",
                 "DWA", "PLA", "SAT"]   # nBodyAlgorithm.cpp:298-342, numbered from 1; anything else -> 0


def g(v):
    """C++ `ostream << double` with the default precision (6 significant digits) == printf %g."""
    return "%g" % v


def expected_vtp(names, classes, m, px, py, pz, vx, vy, vz, anorm, energies=None):
    """Python restatement of the reference's .vtp layout (nBodyAlgorithm.cpp:156-245, 253-395)."""
    n = len(px)
    L = ['<?xml version="1.0"?>',
         '<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian" header_type="UInt64">', "<PolyData>",
         '<Piece NumberOfPoints="%d" NumberOfVerts="%d">' % (n, n), "<Points>",
         '<DataArray type="Float64" Name="position" NumberOfComponents="3" format="ascii">']
    L += ["%s %s %s" % (g(a), g(b), g(c)) for a, b, c in zip(px, py, pz)]
    L += ["</DataArray>", "</Points>", "<PointData>",
          '<DataArray type="Int32" Name="body_id" NumberOfComponents="1" format="ascii">']
    L += [str(j) for j in range(n)]
    L += ["</DataArray>", '<DataArray type="Float64" Name="velocity" NumberOfComponents="3" format="ascii">']
    L += ["%s %s %s" % (g(a), g(b), g(c)) for a, b, c in zip(vx, vy, vz)]
    L += ["</DataArray>", '<DataArray type="Float64" Name="acceleration" NumberOfComponents="1" format="ascii">']
    acc_first = len(L)
    L += [g(a) for a in anorm]
    acc_last = len(L)
    L += ["</DataArray>", '<DataArray type="Float64" Name="mass" NumberOfComponents="1" format="ascii">']
    L += [g(v) for v in m]
    L += ["</DataArray>", '<DataArray type="String" Name="name" NumberOfComponents="1" format="ascii">']
    L += ["".join("%d " % ord(ch) for ch in nm) + " 0" for nm in names]
    L += ["</DataArray>", '<DataArray type="Int32" Name="orbit_class" NumberOfComponents="1" format="ascii">']
    L += [str(ORBIT_CLASSES.index(c) + 1 if c in ORBIT_CLASSES else 0) for c in classes]
    L += ["</DataArray>", "</PointData>", "<Verts>", '<DataArray type="Int64" Name="offsets">',
          "".join("%d " % j for j in range(1, n + 1)), "</DataArray>", '<DataArray type="Int64" Name="connectivity">',
          "".join("%d " % j for j in range(n)), "</DataArray>", "</Verts>", "</Piece>", "<FieldData>"]
    e = energies if energies is not None else [0, 0, 0, 0]
    for name, val in zip(("kinetic energy", "potential energy", "total energy", "virial equilibrium"), e):
        L += ['<DataArray type="Float64" Name="%s" NumberOfTuples="1" format="ascii">' % name,
              g(val) if energies is not None else "0", "</DataArray>"]
    L += ["</FieldData>", "</PolyData>", "</VTKFile>"]
    return L, (acc_first, acc_last)


@pytest.mark.parametrize("stream", [False, True])
def test_vtp_and_pvd_are_byte_compatible(nb, oracle, tmp_path, golden_dir, stream):
    fixture = os.path.join(golden_dir, "solar_178.csv")
    names, classes, m, x, y, z, vx, vy, vz = load_csv(fixture)
    flags = ["--dt=6h", "--t_end=2d", "--vs=1d", "--algorithm=naive"] + (["--stream_output=true"] if stream else [])
    d, _ = run_exe(tmp_path, fixture, *flags)
    ref = oracle.simulate("naive", m, x, y, z, vx, vy, vz, 0.25, 2.0, 1.0)
    assert ref["n_snap"] == 3
    pvd = open(os.path.join(d, "simulation.pvd")).read().splitlines()
    assert pvd[:3] == ['<?xml version="1.0"?>',
                       '<VTKFile type="Collection" version="0.1" byte_order="LittleEndian" compressor="vtkZLibDataCompressor">',
                       "<Collection>"]
    assert pvd[3:6] == ['<DataSet timestep="%d" group="" part="0" file="simulation_step%d.vtp"/>' % (i, i) for i in range(3)]
    assert pvd[6:] == ["</Collection>", "</VTKFile>"]
    for step in range(3):
        got = open(os.path.join(d, "simulation_step%d.vtp" % step)).read().split("\n")
        assert got[-1] == ""          # file ends with a newline
        got = got[:-1]
        want, (a0, a1) = expected_vtp(names, classes, m, ref["px"][step], ref["py"][step], ref["pz"][step],
                                      ref["vx"][step], ref["vy"][step], ref["vz"][step], ref["anorm"][step])
        assert len(got) == len(want)
        if step == 0:
            # positions / adjusted velocities of step 0 are pure host arithmetic: every byte must match
            assert got[:a0] == want[:a0] and got[a1:] == want[a1:]
        else:
            assert got[a1:] == want[a1:]                      # masses, names, classes, cells, field data
            assert got[:6] == want[:6]
        # GPU-computed numbers: equal at the printed precision up to one unit in the last digit
        for gl, wl in zip(got[:a1], want[:a1]):
            if gl != wl:
                gv = np.array(gl.split(), dtype=float); wv = np.array(wl.split(), dtype=float)
                assert gv.shape == wv.shape and np.all(np.abs(gv - wv) <= 1.01e-5 * np.abs(wv) + 1e-300)
    last = open(os.path.join(d, "lastState.csv")).read().splitlines()
    assert last[0] == "position_x, position_y, position_z " and len(last) == 179


# ---- binary state input, checkpoint / resume, binary snapshots ------------------------------------------
def parse_appended_vtp(path):
    """Minimal reader of the VTK XML 'appended raw' layout written by --vtp_format=binary."""
    import re
    raw = open(path, "rb").read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", marker) + 1
    header = raw[:marker].decode()
    dtypes = {"Float64": "<f8", "Int32": "<i4", "Int64": "<i8"}
    arrays = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="([\w ]+)"(?: NumberOfComponents="(\d)")? format="appended" offset="(\d+)"/>', header):
        typ, name, comps, off = m.group(1), m.group(2), int(m.group(3) or 1), int(m.group(4))
        nbytes = int.from_bytes(raw[start + off:start + off + 8], "little")
        a = np.frombuffer(raw[start + off + 8:start + off + 8 + nbytes], dtype=dtypes[typ])
        arrays[name] = a.reshape(-1, comps) if comps > 1 else a
    fields = {m.group(1): float(m.group(2)) for m in
              re.finditer(r'Name="([\w ]+)" NumberOfTuples="1" format="ascii">\n(\S+)\n', header)}
    assert raw.rstrip().endswith(b"</VTKFile>")
    return header, arrays, fields


@pytest.mark.parametrize("algorithm", ["naive", "BarnesHut"])
def test_checkpoint_resume_is_bit_identical(nb, tmp_path, golden_dir, algorithm):
    """10 steps in one run == 5 steps + checkpoint + 5 steps from the checkpoint, to the last bit of x and v."""
    fixture = os.path.join(golden_dir, "solar_178.csv")
    common = ["--dt=6h", "--vs=1d", "--algorithm=" + algorithm]
    full = tmp_path / "full.nbstate"
    run_exe(tmp_path / "a", fixture, "--t_end=60h", "--checkpoint=" + str(full), *common)
    half = tmp_path / "half.nbstate"
    run_exe(tmp_path / "b", fixture, "--t_end=30h", "--checkpoint=" + str(half), *common)
    resumed = tmp_path / "resumed.nbstate"
    run_exe(tmp_path / "c", str(half), "--t_end=30h", "--checkpoint=" + str(resumed), *common)
    a, b, c = (nb.generators.read_state(p) for p in (full, half, resumed))
    assert a["time"] == 2.5 and b["time"] == 1.25 and c["time"] == 2.5
    assert c["names"] == a["names"] and a["names"][2] == "Earth"
    for k in ("m", "x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(a[k], c[k]), k
    assert not np.array_equal(a["x"], b["x"])


def test_checkpoint_at_every_visualised_step(nb, tmp_path, golden_dir):
    fixture = os.path.join(golden_dir, "solar_178.csv")
    ck = tmp_path / "ck.nbstate"
    d, _ = run_exe(tmp_path, fixture, "--dt=6h", "--t_end=2d", "--vs=1d", "--algorithm=naive", "--stream_output=true",
                   "--checkpoint=" + str(ck), "--checkpoint_every_vs=true")
    s = nb.generators.read_state(ck)
    assert s["time"] == 2.0
    last = read_last_state(d)
    assert np.all(np.abs(last[:, 0] - s["x"]) <= 1.01e-5 * np.abs(s["x"]) + 1e-12)


def test_binary_snapshots_hold_the_same_data_at_full_precision(nb, oracle, tmp_path, golden_dir):
    fixture = os.path.join(golden_dir, "solar_178.csv")
    names, classes, m, x, y, z, vx, vy, vz = load_csv(fixture)
    flags = ["--dt=6h", "--t_end=2d", "--vs=1d", "--algorithm=BarnesHut", "--theta=0.5", "--energy=true"]
    d, _ = run_exe(tmp_path, fixture, "--vtp_format=binary", *flags)
    ref = oracle.simulate("BarnesHut", m, x, y, z, vx, vy, vz, 0.25, 2.0, 1.0, theta=0.5, energy=True)
    for step in range(3):
        header, arr, fields = parse_appended_vtp(os.path.join(d, "simulation_step%d.vtp" % step))
        assert '<Piece NumberOfPoints="178" NumberOfVerts="178">' in header and 'Name="name"' not in header
        want_pos = np.stack([ref["px"][step], ref["py"][step], ref["pz"][step]], axis=1)
        want_vel = np.stack([ref["vx"][step], ref["vy"][step], ref["vz"][step]], axis=1)
        scale = np.abs(want_pos).max()
        assert arr["position"].shape == (178, 3) and np.all(np.abs(arr["position"] - want_pos) <= 1e-11 * scale)
        assert np.all(np.abs(arr["velocity"] - want_vel) <= 1e-11 * np.abs(want_vel).max())
        assert np.allclose(arr["acceleration"], ref["anorm"][step], rtol=1e-9, atol=0)
        assert np.array_equal(arr["mass"], m)
        assert np.array_equal(arr["body_id"], np.arange(178)) and np.array_equal(arr["connectivity"], np.arange(178))
        assert np.array_equal(arr["offsets"], np.arange(1, 179))
        assert arr["orbit_class"][0] == 15 and arr["orbit_class"][2] == 17
        assert fields["kinetic energy"] == pytest.approx(ref["energy"][step][0], rel=1e-12)
        assert fields["potential energy"] == pytest.approx(ref["energy"][step][1], rel=1e-12)
    if True:   # step 0 is host arithmetic only: positions exactly as read
        _, arr0, _ = parse_appended_vtp(os.path.join(d, "simulation_step0.vtp"))
        assert np.array_equal(arr0["position"], np.stack([x, y, z], axis=1))


def test_nameless_state_input_runs_and_writes_complete_arrays(nb, oracle, tmp_path):
    """Synthetic bodies enter through the binary state file (no names / classes) instead of a CSV."""
    n = 3000
    m, x, y, z, vx, vy, vz = nb.generators.plummer(n, seed=2)
    state = tmp_path / "plummer.nbstate"
    nb.generators.write_state(state, m, x, y, z, vx, vy, vz)
    d, _ = run_exe(tmp_path, str(state), "--dt=1h", "--t_end=3h", "--vs=1h", "--algorithm=BarnesHut", "--theta=0.5")
    ref = oracle.simulate("BarnesHut", m, x, y, z, vx, vy, vz, 1.0 / 24, 3.0 / 24, 1.0 / 24, theta=0.5)
    last = read_last_state(d)
    want = ref["px"][ref["n_snap"] - 1]
    assert np.all(np.abs(last[:, 0] - want) <= 1.01e-5 * np.abs(want) + 1e-12)
    vtp = open(os.path.join(d, "simulation_step1.vtp")).read()
    name_block = vtp.split('Name="name"')[1].split("</DataArray>")[0].split("\n")[1:-1]
    class_block = vtp.split('Name="orbit_class"')[1].split("</DataArray>")[0].split("\n")[1:-1]
    assert len(name_block) == n and set(name_block) == {" 0"}
    assert len(class_block) == n and set(class_block) == {"0"}
