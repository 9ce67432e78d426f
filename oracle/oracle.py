"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (n-body-simulation_b200/) never imports this module.
Every function restates a reference function; see the citations in nbody_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile liboracle.so with oracle/Makefile (g++, OpenMP, -ffp-contract=off)."""
    src = os.path.join(_HERE, "nbody_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_gravitational_constant.restype = C.c_double
        L.orc_epsilon2.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        L.orc_tree_create.restype = C.c_void_p
        L.orc_tree_create.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_tree_destroy.argtypes = [C.c_void_p]
        L.orc_tree_build.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_tree_build_ordered.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, _up]
        L.orc_tree_num_nodes.restype = C.c_uint32
        L.orc_tree_num_nodes.argtypes = [C.c_void_p]
        L.orc_tree_max_depth.argtypes = [C.c_void_p]
        L.orc_tree_aabb.argtypes = [C.c_void_p, _dp]
        for name, rt in [("body_of_node", _up), ("body_count", _up), ("octants", _up), ("is_leaf", C.POINTER(C.c_int)),
                         ("sum_masses", _dp), ("com_x", _dp), ("com_y", _dp), ("com_z", _dp), ("edge", _dp),
                         ("min_x", _dp), ("min_y", _dp), ("min_z", _dp), ("sorted_bodies", _up)]:
            f = getattr(L, "orc_tree_" + name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        L.orc_tree_canonical.argtypes = [C.c_void_p, _up, _u64p, _u64p, _up, _up, _up] + [_dp] * 8
        L.orc_bh_accel.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp,
                                   _dp, _u64p, C.c_int]
        L.orc_bh_accel_sample.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_uint32, _up,
                                          _dp, _dp, _dp, _u64p, C.c_int]
        L.orc_naive_accel_rows.argtypes = [C.c_uint32, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_uint32,
                                           C.c_uint32, _dp, _dp, _dp, C.c_int]
        L.orc_leapfrog_part1.argtypes = [C.c_uint32, C.c_double] + [_dp] * 12
        L.orc_leapfrog_part2.argtypes = [C.c_uint32, C.c_double] + [_dp] * 9
        L.orc_energy.argtypes = [C.c_uint32, C.c_double] + [_dp] * 7 + [_dp, _dp, _dp]
        L.orc_accel_norm.argtypes = [C.c_uint32, _dp, _dp, _dp, _dp]
        L.orc_adjust_velocities.argtypes = [C.c_uint32] + [_dp] * 7
        L.orc_aabb.argtypes = [C.c_uint32, _dp, _dp, _dp, C.c_int, _dp]
        L.orc_init_config.argtypes = [C.c_uint32, C.c_int, C.c_int, _up, _up]
        L.orc_prepare_subtrees.argtypes = [C.c_uint32, _up, C.c_uint32, _up, _up, _up]
        L.orc_sort_bodies_for_subtrees.argtypes = [C.c_uint32, _up, _up, _up, C.c_uint32, _up, _up]
        L.orc_simulate.argtypes = ([C.c_int, C.c_uint32] + [_dp] * 7 + [C.c_double] * 4 + [C.c_int] * 5 +
                                   [C.c_uint32] + [_dp] * 8 + [_up, _u64p])
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_up)


def gravitational_constant():
    return lib().orc_gravitational_constant()


def epsilon2():
    return lib().orc_epsilon2()


def max_threads():
    return lib().orc_max_threads()


def init_config(n, storage_param=16, stack_param=16):
    s, k = C.c_uint32(), C.c_uint32()
    lib().orc_init_config(n, storage_param, stack_param, C.byref(s), C.byref(k))
    return s.value, k.value


def naive_accel(m, x, y, z, eps2=None, G=None, rows=None, nthreads=0):
    """NaiveAlgorithm::computeAccelerations_opt_{0,1,2}; rows=(i0,i1) restricts the target rows."""
    L = lib()
    eps2 = L.orc_epsilon2() if eps2 is None else eps2
    G = L.orc_gravitational_constant() if G is None else G
    m, pm = _d(m); x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    n = x.shape[0]
    i0, i1 = (0, n) if rows is None else rows
    ax = np.zeros(n); ay = np.zeros(n); az = np.zeros(n)
    L.orc_naive_accel_rows(n, pm, px, py, pz, eps2, G, i0, i1, ax.ctypes.data_as(_dp), ay.ctypes.data_as(_dp),
                           az.ctypes.data_as(_dp), nthreads)
    return ax, ay, az


def leapfrog_part1(dt, x, y, z, vx, vy, vz, ax, ay, az):
    """In place on x,y,z; returns the half-step velocities."""
    n = x.shape[0]
    vh = [np.zeros(n) for _ in range(3)]
    args = [a.ctypes.data_as(_dp) for a in (x, y, z, vx, vy, vz, *vh, ax, ay, az)]
    lib().orc_leapfrog_part1(n, dt, *args)
    return vh


def leapfrog_part2(dt, vx, vy, vz, vhx, vhy, vhz, ax, ay, az):
    n = vx.shape[0]
    args = [a.ctypes.data_as(_dp) for a in (vx, vy, vz, vhx, vhy, vhz, ax, ay, az)]
    lib().orc_leapfrog_part2(n, dt, *args)


def energy(m, x, y, z, vx, vy, vz, G=None, per_body=False):
    L = lib()
    G = L.orc_gravitational_constant() if G is None else G
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m, x, y, z, vx, vy, vz)]
    n = arrs[0].shape[0]
    out = np.zeros(4)
    ek = np.zeros(n) if per_body else None
    ep = np.zeros(n) if per_body else None
    L.orc_energy(n, G, *[a.ctypes.data_as(_dp) for a in arrs], out.ctypes.data_as(_dp),
                 ek.ctypes.data_as(_dp) if per_body else None, ep.ctypes.data_as(_dp) if per_body else None)
    return (out, ek, ep) if per_body else out


def accel_norm(ax, ay, az):
    ax, pax = _d(ax); ay, pay = _d(ay); az, paz = _d(az)
    out = np.zeros(ax.shape[0])
    lib().orc_accel_norm(ax.shape[0], pax, pay, paz, out.ctypes.data_as(_dp))
    return out


def adjust_velocities(m, vx, vy, vz):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m, vx, vy, vz)]
    n = arrs[0].shape[0]
    outs = [np.zeros(n) for _ in range(3)]
    lib().orc_adjust_velocities(n, *[a.ctypes.data_as(_dp) for a in arrs + outs])
    return outs


def aabb(x, y, z, work_items=1024):
    x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    out = np.zeros(7)
    lib().orc_aabb(x.shape[0], px, py, pz, work_items, out.ctypes.data_as(_dp))
    return out  # min xyz, max xyz, edge


def prepare_subtrees(subtree_of_body, node_count):
    s, ps = _u(subtree_of_body)
    counts = np.zeros(node_count, dtype=np.uint32)
    subtrees = np.zeros(node_count, dtype=np.uint32)
    cnt = C.c_uint32()
    lib().orc_prepare_subtrees(s.shape[0], ps, node_count, counts.ctypes.data_as(_up), subtrees.ctypes.data_as(_up),
                               C.byref(cnt))
    return counts, subtrees, cnt.value


def sort_bodies_for_subtrees(subtree_of_body, counts, subtrees, n_subtrees):
    s, ps = _u(subtree_of_body)
    counts, pc = _u(counts)
    subtrees, pt = _u(subtrees)
    start = np.zeros(n_subtrees, dtype=np.uint32)
    sorted_bodies = np.zeros(s.shape[0], dtype=np.uint32)
    lib().orc_sort_bodies_for_subtrees(s.shape[0], ps, pc, pt, n_subtrees, start.ctypes.data_as(_up),
                                       sorted_bodies.ctypes.data_as(_up))
    return start, sorted_bodies


def morton_order(x, y, z, bits=20):
    """Permutation that visits the bodies along a Z-order curve of a 2^bits grid (locality only: nothing depends on it)."""
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    lo = min(x.min(), y.min(), z.min(), 0.0)
    hi = max(x.max(), y.max(), z.max(), 0.0)
    scale = ((1 << bits) - 1) / (hi - lo if hi > lo else 1.0)
    q = [np.clip((a - lo) * scale, 0, (1 << bits) - 1).astype(np.uint64) for a in (x, y, z)]
    code = spread(q[0]) | (spread(q[1]) << np.uint64(1)) | (spread(q[2]) << np.uint64(2))
    return np.argsort(code, kind="stable").astype(np.uint32)


class Tree:
    """BarnesHutOctree (canonical tree via sequential insertion) + its COM and in-order sort."""

    def __init__(self, m, x, y, z, storage_param=16, aabb_work_items=1024, insertion_order=None):
        """insertion_order: None (bodies 0..N-1), an explicit permutation, or "morton" (approximate Morton order computed
        here: the canonical tree is the same for every order, a cache-friendly one builds large trees much faster)."""
        L = lib()
        self.m, self._pm = _d(m)
        self.x, self._px = _d(x)
        self.y, self._py = _d(y)
        self.z, self._pz = _d(z)
        self.N = self.x.shape[0]
        self.S = storage_param * self.N
        self._h = C.c_void_p(L.orc_tree_create(self.N, self.S))
        if insertion_order is None:
            rc = L.orc_tree_build(self._h, self._px, self._py, self._pz, self._pm, aabb_work_items)
        else:
            if isinstance(insertion_order, str):
                insertion_order = morton_order(self.x, self.y, self.z)
            order, po = _u(insertion_order)
            assert order.shape[0] == self.N
            rc = L.orc_tree_build_ordered(self._h, self._px, self._py, self._pz, self._pm, aabb_work_items, po)
        if rc:
            raise RuntimeError("oracle tree build failed: %s" % {1: "node storage overflow", 2: "depth guard"}[rc])
        self.num_nodes = L.orc_tree_num_nodes(self._h)
        self.max_depth = L.orc_tree_max_depth(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_destroy(self._h)
            self._h = None

    def _arr(self, name, dtype, n):
        p = getattr(lib(), "orc_tree_" + name)(self._h)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)

    def aabb(self):
        out = np.zeros(7)
        lib().orc_tree_aabb(self._h, out.ctypes.data_as(_dp))
        return out

    @property
    def body_of_node(self):
        return self._arr("body_of_node", np.uint32, self.num_nodes)

    @property
    def body_count(self):
        return self._arr("body_count", np.uint32, self.num_nodes)

    @property
    def sum_masses(self):
        return self._arr("sum_masses", np.float64, self.num_nodes)

    @property
    def sorted_bodies(self):
        return self._arr("sorted_bodies", np.uint32, self.N)

    def canonical(self):
        """dict of arrays, one entry per node, in DFS order (children by ascending octant code)."""
        n = self.num_nodes
        out = dict(depth=np.zeros(n, np.uint32), path_hi=np.zeros(n, np.uint64), path_lo=np.zeros(n, np.uint64),
                   kind=np.zeros(n, np.uint32), body=np.zeros(n, np.uint32), count=np.zeros(n, np.uint32))
        for k in ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz"):
            out[k] = np.zeros(n)
        lib().orc_tree_canonical(self._h, out["depth"].ctypes.data_as(_up), out["path_hi"].ctypes.data_as(_u64p),
                                 out["path_lo"].ctypes.data_as(_u64p), out["kind"].ctypes.data_as(_up),
                                 out["body"].ctypes.data_as(_up), out["count"].ctypes.data_as(_up),
                                 *[out[k].ctypes.data_as(_dp) for k in
                                   ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz")])
        return out

    def accel(self, theta, eps2=None, G=None, sort_bodies=True, stats=False, nthreads=0, x=None, y=None, z=None):
        """BarnesHutAlgorithm::computeAccelerations on this tree."""
        L = lib()
        eps2 = L.orc_epsilon2() if eps2 is None else eps2
        G = L.orc_gravitational_constant() if G is None else G
        n = self.N
        ax = np.zeros(n); ay = np.zeros(n); az = np.zeros(n)
        st = np.zeros((n, 5), dtype=np.uint64) if stats else None
        L.orc_bh_accel(self._h, self._px, self._py, self._pz, theta, eps2, G, int(sort_bodies),
                       ax.ctypes.data_as(_dp), ay.ctypes.data_as(_dp), az.ctypes.data_as(_dp),
                       st.ctypes.data_as(_u64p) if stats else None, nthreads)
        return (ax, ay, az, st) if stats else (ax, ay, az)

    def accel_sample(self, theta, ids, eps2=None, G=None, stats=False, nthreads=0):
        """The same traversal for the sampled bodies `ids` only (outputs indexed by sample position)."""
        L = lib()
        eps2 = L.orc_epsilon2() if eps2 is None else eps2
        G = L.orc_gravitational_constant() if G is None else G
        ids, pids = _u(ids)
        k = ids.shape[0]
        ax = np.zeros(k); ay = np.zeros(k); az = np.zeros(k)
        st = np.zeros((k, 5), dtype=np.uint64) if stats else None
        L.orc_bh_accel_sample(self._h, self._px, self._py, self._pz, theta, eps2, G, k, pids, ax.ctypes.data_as(_dp),
                              ay.ctypes.data_as(_dp), az.ctypes.data_as(_dp),
                              st.ctypes.data_as(_u64p) if stats else None, nthreads)
        return (ax, ay, az, st) if stats else (ax, ay, az)


def simulate(algorithm, m, x, y, z, vx, vy, vz, dt, t_end, vs, theta=1.05, energy=False, sort_bodies=True,
             storage_param=16, aabb_work_items=1024, nthreads=0, max_snap=None):
    """{Naive,BarnesHut}Algorithm::startSimulation. Returns dict with per-snapshot arrays."""
    L = lib()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m, x, y, z, vx, vy, vz)]
    n = arrs[0].shape[0]
    if max_snap is None:
        max_snap = int(t_end / vs) + 3
    snaps = {k: np.zeros((max_snap, n)) for k in ("px", "py", "pz", "vx", "vy", "vz", "anorm")}
    en = np.zeros((max_snap, 4))
    ns, nst = C.c_uint32(), C.c_uint64()
    alg = {"naive": 0, "BarnesHut": 1}[algorithm]
    rc = L.orc_simulate(alg, n, *[a.ctypes.data_as(_dp) for a in arrs], dt, t_end, vs, theta, int(energy),
                        int(sort_bodies), storage_param, aabb_work_items, nthreads, max_snap,
                        *[snaps[k].ctypes.data_as(_dp) for k in ("px", "py", "pz", "vx", "vy", "vz", "anorm")],
                        en.ctypes.data_as(_dp), C.byref(ns), C.byref(nst))
    if rc:
        raise RuntimeError("oracle simulate failed rc=%d" % rc)
    out = {k: v[:ns.value] for k, v in snaps.items()}
    out["energy"] = en[:ns.value]
    out["n_snap"] = ns.value
    out["n_steps"] = nst.value
    return out
