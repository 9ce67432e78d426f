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
