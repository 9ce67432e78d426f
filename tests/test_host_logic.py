"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol (no compute calls), the
host C++ mirror of the reference's utilities behaves like the reference's own googletests, slice partitioning, and the
synthetic generators.  Restates reference tests/InputParserTest.cpp:4-32 and tests/TimeConverterTest.cpp:4-47."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "n-body-simulation_b200", "host")


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(ROOT, "tests", "_host_shim.so")
    srcs = [os.path.join(ROOT, "tests", "host_shim.cpp")] + [os.path.join(HOST, f) for f in
                                                             ("InputParser.cpp", "StateFile.cpp", "TimeConverter.cpp")]
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # Configuration.cpp needs nb_config_default from the CUDA library; the shim only needs initializeConfigValues,
    # so compile Configuration.cpp against the built library
    lib_dir = os.path.join(ROOT, "n-body-simulation_b200")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-I", HOST, "-I", os.path.join(ROOT, "include")]
                          + srcs + [os.path.join(HOST, "Configuration.cpp"), "-o", so, "-L", lib_dir, "-lnbody_b200",
                                    "-Wl,-rpath," + lib_dir])
    L = C.CDLL(so)
    L.shim_convert_time.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    L.shim_split.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.shim_convert_to_state.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
    L.shim_start_time.argtypes = [C.c_char_p]
    L.shim_start_time.restype = C.c_double
    return L


def split(shim, s):
    buf = C.create_string_buffer(4096)
    n = shim.shim_split(s.encode(), buf, 4096)
    parts = buf.value.decode().split("\x1f")
    assert len(parts) == n
    return parts


def convert(shim, s):
    out = C.c_double()
    rc = shim.shim_convert_time(s.encode(), C.byref(out))
    return None if rc else out.value


# ---- InputParserTest.cpp ---------------------------------------------------------------------------------------------
def test_split_basic_case(nb, shim):
    r = split(shim, "5,Name,Class,5000,0.5,-0.3,0.4,0.06,0.1,0.2")
    assert r == ["5", "Name", "Class", "5000", "0.5", "-0.3", "0.4", "0.06", "0.1", "0.2"]


def test_split_empty_entry(nb, shim):
    r = split(shim, "5,,Class,5000,0.5,-0.3,0.4,0.06,0.1,0.2")
    assert r[1] == "" and len(r) == 10 and r[9] == "0.2"


def test_split_trailing_empty_field(nb, shim):
    assert split(shim, "a,b,") == ["a", "b", ""]


# ---- TimeConverterTest.cpp ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("text,expect", [("1h", 1.0 / 24), ("10h", 10.0 / 24), ("1m", 30.4167), ("1y", 365.25),
                                         ("0.5y", 182.625), ("365d", 365.0)])
def test_time_conversion(nb, shim, text, expect):
    assert convert(shim, text) == pytest.approx(expect, rel=1e-7)


@pytest.mark.parametrize("text", ["0.5s", "0.5", "1y 5m", "h", ""])
def test_time_conversion_invalid(nb, shim, text):
    assert convert(shim, text) is None


def test_parse_solar_fixture(nb, shim, golden_dir):
    n = 178
    arrs = [np.zeros(n) for _ in range(7)]
    names = C.create_string_buffer(1 << 16)
    shim.shim_parse_csv.argtypes = [C.c_char_p, C.c_int] + [C.POINTER(C.c_double)] * 7 + [C.c_char_p, C.c_int]
    got = shim.shim_parse_csv(os.path.join(golden_dir, "solar_178.csv").encode(), n,
                              *[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs], names, 1 << 16)
    assert got == 178
    tags = names.value.decode().split("\x1f")
    assert tags[0] == "Sun|STA" and arrs[0][0] == 1.98847e30
    assert tags[2] == "Earth|PLA" and arrs[0][2] == pytest.approx(5.97219e24)
    assert np.hypot(arrs[1][2], arrs[2][2]) == pytest.approx(0.9833, abs=2e-3)  # Earth near perihelion on 1 Jan 2000


def _parse(shim, path, n):
    arrs = [np.zeros(n) for _ in range(7)]
    names = C.create_string_buffer(1 << 16)
    shim.shim_parse_csv.argtypes = [C.c_char_p, C.c_int] + [C.POINTER(C.c_double)] * 7 + [C.c_char_p, C.c_int]
    got = shim.shim_parse_csv(str(path).encode(), n, *[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs], names, 1 << 16)
    return got, arrs, names.value.decode()


def test_state_file_round_trip_keeps_every_bit(nb, shim, golden_dir, tmp_path):
    """CSV -> binary state (C++ writer) -> C++ reader and Python reader: all seven arrays bitwise, names kept."""
    csv = os.path.join(golden_dir, "solar_178.csv")
    state = tmp_path / "solar.nbstate"
    assert shim.shim_convert_to_state(csv.encode(), str(state).encode(), 12.5) == 178
    n0, a0, tags0 = _parse(shim, csv, 178)
    n1, a1, tags1 = _parse(shim, state, 178)
    assert n0 == n1 == 178 and tags0 == tags1
    for u, v in zip(a0, a1):
        assert np.array_equal(u, v)
    assert shim.shim_start_time(str(state).encode()) == 12.5 and shim.shim_start_time(csv.encode()) == 0.0
    py = nb.generators.read_state(state)
    assert py["time"] == 12.5 and py["names"][2] == "Earth" and py["classes"][0] == "STA"
    for u, k in zip(a0, ("m", "x", "y", "z", "vx", "vy", "vz")):
        assert np.array_equal(u, py[k])


def test_state_file_written_by_python_is_read_by_the_host(nb, shim, tmp_path):
    m, x, y, z, vx, vy, vz = nb.generators.plummer(1000, seed=4)
    path = tmp_path / "p.nbstate"
    nb.generators.write_state(path, m, x, y, z, vx, vy, vz, time=3.0)
    n, arrs, tags = _parse(shim, path, 1000)
    assert n == 1000 and tags == ""            # no name table
    for u, v in zip(arrs, (m, x, y, z, vx, vy, vz)):
        assert np.array_equal(u, v)
    assert shim.shim_start_time(str(path).encode()) == 3.0


def test_state_file_errors(nb, shim, tmp_path):
    m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(64, seed=1)
    good = tmp_path / "g.nbstate"
    nb.generators.write_state(good, m, x, y, z, vx, vy, vz)
    raw = good.read_bytes()
    cut = tmp_path / "cut.nbstate"
    cut.write_bytes(raw[:-16])                  # truncated array
    assert _parse(shim, cut, 64)[0] == -1
    ver = tmp_path / "ver.nbstate"
    ver.write_bytes(raw[:8] + (2).to_bytes(4, "little") + raw[12:])   # unknown version
    assert _parse(shim, ver, 64)[0] == -1
    empty = tmp_path / "empty.nbstate"
    nb.generators.write_state(empty, *[np.zeros(0)] * 7)
    assert _parse(shim, empty, 1)[0] == 0       # valid, no bodies (main() rejects it like an empty CSV)
    with pytest.raises(ValueError):
        nb.generators.write_state(tmp_path / "bad", m, x[:5], y, z, vx, vy, vz)


def test_config_values_match_oracle(nb, shim, oracle):
    for n, sp, kp in [(3, 3, 3), (178, 16, 16), (1 << 20, 16, 16), (14999, 16, 16), (15000, 16, 16)]:
        a, b = C.c_uint(), C.c_uint()
        shim.shim_init_config(n, sp, kp, C.byref(a), C.byref(b))
        assert (a.value, b.value) == oracle.init_config(n, sp, kp)


# ---- C ABI surface ---------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(nb):
    header = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    declared = set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", header))
    declared -= {"nb_ctx"}
    L = nb.load_library()
    binding = __import__("importlib").import_module("n-body-simulation_b200.binding")
    assert declared == set(binding.EXPORTED_SYMBOLS)
    for sym in sorted(declared):
        assert hasattr(L, sym), sym
    assert L.nb_abi_version() == 1


def test_clean_rebuild_from_sources(nb, tmp_path):
    """The in-tree library is reused when it is newer than its sources (build(): mtime check).  This is the clean build:
    every .cu compiled for sm_100a from scratch into a scratch directory (nvcc cross-compiles without a GPU), linked with
    no undefined symbols, exporting every entry point the header declares, carrying sm_100a code only."""
    import ctypes
    import subprocess
    build = __import__("importlib").import_module("n-body-simulation_b200.build")
    lib = build.build_library(force=True, out_dir=str(tmp_path))
    assert os.path.dirname(lib) == str(tmp_path) and os.path.getsize(lib) > 100000
    L = ctypes.CDLL(lib)
    binding = __import__("importlib").import_module("n-body-simulation_b200.binding")
    for sym in binding.EXPORTED_SYMBOLS:
        assert hasattr(L, sym), sym
    assert L.nb_abi_version() == 1
    elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf[:400]


def test_default_config_matches_reference_defaults(nb, oracle):
    cfg = nb.default_config()
    assert cfg.G == oracle.gravitational_constant() and cfg.epsilon2 == oracle.epsilon2()
    assert (cfg.theta, cfg.block_size, cfg.opt_stage, cfg.sort_bodies, cfg.wg_size_barnes_hut) == (1.05, 64, 2, 1, 64)
    assert (cfg.num_wi_aabb, cfg.num_wi_octree, cfg.num_wi_top_octree, cfg.num_wi_com, cfg.max_level_top_octree) == (
        1024, 640, 1024, 1024, 7)
    assert (cfg.storage_size_param, cfg.stack_size_param) == (16, 16)


def test_no_cpu_fallback(nb):
    """Without a GPU the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nb.NBodyError) as e:
        nb.Context()
    assert e.value.status == -2


def test_opt_stage_range_checked(nb):
    cfg = nb.default_config(opt_stage=3)   # main.cpp:154-157: must be 0, 1 or 2
    h = C.c_void_p()
    assert nb.load_library().nb_create(C.byref(cfg), C.byref(h)) == -1


def test_slice_bounds_cover_and_balance(nb):
    for n in (1, 7, 178, 1 << 20, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [nb.slice_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            chunk = -(-n // world)
            for r, (b, e) in enumerate(spans):
                assert b == min(r * chunk, n) and e == min((r + 1) * chunk, n)
            assert sum(e - b for b, e in spans) == n


def test_host_executable_rejects_bad_arguments(nb):
    exe = os.path.join(ROOT, "n-body-simulation_b200", "N_Body_Simulation")
    assert os.path.exists(exe)
    fixture = os.path.join(ROOT, "tests", "golden", "solar_178.csv")
    base = ["--file=" + fixture, "--dt=1h", "--t_end=1d", "--vs=1d", "--vs_dir=/tmp/nb_out"]
    r = subprocess.run([exe] + base + ["--algorithm=fast"], capture_output=True, text=True)
    assert r.returncode != 0 and "Algorithm must either be <naive> or <BarnesHut>" in r.stderr
    r = subprocess.run([exe] + base + ["--algorithm=naive", "--opt_stage=5"], capture_output=True, text=True)
    assert r.returncode != 0 and "Optimization stage must be 0,1 or 2" in r.stderr
    r = subprocess.run([exe] + base[:-1] + ["--algorithm=naive"], capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([exe] + base + ["--algorithm=naive", "--dt=1x"], capture_output=True, text=True)
    assert r.returncode != 0 and "Time values have to be of the format" in r.stderr


# ---- generators ----------------------------------------------------------------------------------------------------------------
def test_generators_are_deterministic_and_rank_independent(nb):
    a = nb.generators.plummer(1000, seed=3)
    b = nb.generators.plummer(1000, seed=3)
    c = nb.generators.plummer(2000, seed=3)
    for u, v, w in zip(a, b, c):
        assert np.array_equal(u, v)
    # counter based: body i does not depend on N (positions; masses scale with 1/N)
    assert np.array_equal(a[1], c[1][:1000]) and np.array_equal(a[4], c[4][:1000])
    assert not np.array_equal(a[1], nb.generators.plummer(1000, seed=4)[1])


def test_plummer_shape(nb):
    m, x, y, z, vx, vy, vz = nb.generators.plummer(20000, seed=1)
    r = np.sqrt(x * x + y * y + z * z)
    assert r.max() <= 50.0 and np.median(r) == pytest.approx(1.3048, rel=0.05)   # half-mass radius 1.305 a
    assert m.sum() == pytest.approx(nb.generators.M_SUN)
    assert len(np.unique(np.stack([x, y, z], 1), axis=0)) == 20000


def test_uniform_sphere_shape(nb):
    m, x, y, z, vx, vy, vz = nb.generators.uniform_sphere(20000, seed=1)
    r = np.sqrt(x * x + y * y + z * z)
    assert r.max() <= 1.0 and np.median(r) == pytest.approx(0.5 ** (1 / 3), rel=0.03)
    assert not vx.any()


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch types) and a C program must
    link against the library using only what the header declares."""
    header = os.path.join(ROOT, "include", "nbody_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", header])
    src = tmp_path / "probe.c"
    src.write_text('#include "nbody_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { nb_config c; nb_ctx *ctx = 0; nb_config_default(&c);\n'
                   '  int rc = nb_create(&c, &ctx);\n'
                   '  printf("%d %d %s %.17g\\n", nb_abi_version(), rc, nb_status_string(rc), c.G);\n'
                   '  if (ctx) nb_destroy(ctx); return 0; }\n')
    exe = tmp_path / "probe"
    lib_dir = os.path.join(ROOT, "n-body-simulation_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir,
                           "-lnbody_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0
    fields = out.stdout.split()
    assert fields[0] == "1"
    import torch
    if not torch.cuda.is_available():
        assert fields[1] == "-2"          # NB_ERR_NO_DEVICE: no CPU fallback
