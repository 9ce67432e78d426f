"""Seeded synthetic bodies for the benchmarks and parity tests.

Counter-based: every random number is a SplitMix64 hash of (seed, body id, draw index), so any rank / the CPU oracle
generates bit-identical bodies without communicating.  Units are the reference's: AU, days, kg.
"""
import numpy as np

M_SUN = 1.98847e30  # kg
_G_AU_DAY = 1.488180711053671e-34  # AU^3 kg^-1 day^-2 (nBodyAlgorithm.hpp:55-61)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x.copy()
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def _uniform(seed, ids, draw):
    """U(0,1) (open interval) for body `ids` and draw index `draw` (arrays broadcast)."""
    with np.errstate(over="ignore"):
        k = _splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + ids.astype(np.uint64))
        k = _splitmix64(k ^ (np.asarray(draw, dtype=np.uint64) * np.uint64(0xA24BAED4963EE407)))
    return ((k >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def uniform_sphere(n, seed=1, radius=1.0, total_mass=M_SUN, velocity_scale=0.0, first_id=0, n_total=None):
    """Uniform-density sphere of radius `radius` AU, equal masses; cold (v=0) unless velocity_scale > 0.
    first_id / n_total: bodies first_id .. first_id + n - 1 of a set of n_total (the generator is counter-based, so a
    large set can be produced in chunks without large temporaries)."""
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    r = radius * np.cbrt(_uniform(seed, ids, 0))
    cos_t = 2.0 * _uniform(seed, ids, 1) - 1.0
    phi = 2.0 * np.pi * _uniform(seed, ids, 2)
    sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
    x, y, z = r * sin_t * np.cos(phi), r * sin_t * np.sin(phi), r * cos_t
    m = np.full(n, total_mass / (n if n_total is None else n_total))
    if velocity_scale > 0:
        sigma = velocity_scale * np.sqrt(_G_AU_DAY * total_mass / radius)
        vx = sigma * (2.0 * _uniform(seed, ids, 3) - 1.0)
        vy = sigma * (2.0 * _uniform(seed, ids, 4) - 1.0)
        vz = sigma * (2.0 * _uniform(seed, ids, 5) - 1.0)
    else:
        vx = np.zeros(n); vy = np.zeros(n); vz = np.zeros(n)
    if n_total is not None:
        return m, x, y, z, vx, vy, vz
    return _dedup(m, x, y, z, vx, vy, vz)


def plummer(n, seed=1, a=1.0, total_mass=M_SUN, r_max=50.0):
    """Plummer sphere, scale radius `a` AU, radii rejected above r_max*a, velocities by Aarseth's rejection method."""
    ids = np.arange(n, dtype=np.uint64)
    r = np.empty(n)
    todo = np.arange(n)
    draw = 0
    while todo.size:
        u = _uniform(seed, ids[todo], 10 + draw)
        rr = a / np.sqrt(np.maximum(u ** (-2.0 / 3.0) - 1.0, 1e-300))
        ok = rr <= r_max * a
        r[todo[ok]] = rr[ok]
        todo = todo[~ok]
        draw += 1
    cos_t = 2.0 * _uniform(seed, ids, 1) - 1.0
    phi = 2.0 * np.pi * _uniform(seed, ids, 2)
    sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
    x, y, z = r * sin_t * np.cos(phi), r * sin_t * np.sin(phi), r * cos_t
    # speed: q in (0,1) with density g(q) = q^2 (1-q^2)^(7/2), accepted when 0.1*u2 < g(q)
    q = np.empty(n)
    todo = np.arange(n)
    draw = 0
    while todo.size:
        q1 = _uniform(seed, ids[todo], 1000 + 2 * draw)
        q2 = _uniform(seed, ids[todo], 1001 + 2 * draw)
        ok = 0.1 * q2 < q1 * q1 * (1.0 - q1 * q1) ** 3.5
        q[todo[ok]] = q1[ok]
        todo = todo[~ok]
        draw += 1
    v = q * np.sqrt(2.0 * _G_AU_DAY * total_mass / a) * (1.0 + (r / a) ** 2) ** (-0.25)
    cos_v = 2.0 * _uniform(seed, ids, 3) - 1.0
    phi_v = 2.0 * np.pi * _uniform(seed, ids, 4)
    sin_v = np.sqrt(np.maximum(0.0, 1.0 - cos_v * cos_v))
    vx, vy, vz = v * sin_v * np.cos(phi_v), v * sin_v * np.sin(phi_v), v * cos_v
    m = np.full(n, total_mass / n)
    return _dedup(m, x, y, z, vx, vy, vz)


def _dedup(m, x, y, z, vx, vy, vz):
    """Coincident bodies are unsupported by the reference (unbounded splitting): nudge exact duplicates apart."""
    key = np.stack([x, y, z], axis=1)
    _, first = np.unique(key, axis=0, return_index=True) if key.shape[0] < (1 << 22) else (None, None)
    if first is not None and first.size != key.shape[0]:
        dup = np.setdiff1d(np.arange(key.shape[0]), first)
        x[dup] += 1e-9 * (1.0 + np.arange(dup.size))
    return m, x, y, z, vx, vy, vz


def solar_like(n=178, seed=3):
    """A small star + satellites system with unequal masses (stand-in when the CSV fixture is not used)."""
    ids = np.arange(n, dtype=np.uint64)
    a = 0.3 + 40.0 * _uniform(seed, ids, 0) ** 2
    phi = 2.0 * np.pi * _uniform(seed, ids, 1)
    inc = 0.05 * (2.0 * _uniform(seed, ids, 2) - 1.0)
    x, y, z = a * np.cos(phi), a * np.sin(phi), a * inc
    m = 10.0 ** (20.0 + 7.0 * _uniform(seed, ids, 3))
    vc = np.sqrt(_G_AU_DAY * M_SUN / a)
    vx, vy, vz = -vc * np.sin(phi), vc * np.cos(phi), np.zeros(n)
    x[0] = y[0] = z[0] = 0.0
    vx[0] = vy[0] = vz[0] = 0.0
    m[0] = M_SUN
    return m, x, y, z, vx, vy, vz


# ---- binary body-state files (host/StateFile.hpp) ---------------------------------------------------------------------
STATE_MAGIC = b"NBSTATE1"


def write_state(path, m, x, y, z, vx, vy, vz, time=0.0, names=None, classes=None):
    """Writes the bodies as the binary SoA state file `N_Body_Simulation --file=` accepts (no CSV round trip)."""
    import struct
    arrays = [np.ascontiguousarray(a, dtype="<f8") for a in (m, x, y, z, vx, vy, vz)]
    n = arrays[0].size
    if any(a.size != n for a in arrays):
        raise ValueError("state arrays differ in length")
    table = names is not None and classes is not None and n > 0
    with open(path, "wb") as f:
        f.write(struct.pack("<8sIIQd32x", STATE_MAGIC, 1, 1 if table else 0, n, float(time)))
        for a in arrays:
            f.write(a.tobytes())
        if table:
            for nm, cl in zip(names, classes):
                f.write(nm.encode() + b"\0" + cl.encode() + b"\0")


def read_state(path):
    """-> dict(m, x, y, z, vx, vy, vz, time, names, classes) of a binary state file / checkpoint."""
    import struct
    with open(path, "rb") as f:
        magic, version, flags, n, time = struct.unpack("<8sIIQd32x", f.read(64))
        if magic != STATE_MAGIC or version != 1:
            raise ValueError("not a version-1 body-state file: %s" % path)
        out = {"time": time}
        for k in ("m", "x", "y", "z", "vx", "vy", "vz"):
            out[k] = np.frombuffer(f.read(8 * n), dtype="<f8").copy()
            if out[k].size != n:
                raise ValueError("state file is truncated: %s" % path)
        out["names"], out["classes"] = None, None
        if flags & 1:
            fields = f.read().split(b"\0")
            out["names"] = [b.decode() for b in fields[0:2 * n:2]]
            out["classes"] = [b.decode() for b in fields[1:2 * n:2]]
    return out
