"""ctypes binding of oracle/_ref/libnbody_ref.so: the reference's OWN, unmodified sources compiled with g++ through the
host SYCL subset in oracle/sycl_shim (recipe: oracle/Makefile target `ref`; driver: oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, tools/make_golden.py and bench.py's CPU legs.  The product
(n-body-simulation_b200/) never imports this module.  /root/reference exists only in the build container; the GPU box
gets the prebuilt oracle/_ref/ (git-ignored, not gpurun-ignored).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_LIB_PATH = os.path.join(REF_DIR, "libnbody_ref.so")
EXE_PATH = os.path.join(REF_DIR, "N_Body_Simulation")
REFERENCE_ROOT = os.environ.get("NBODY_REFERENCE_ROOT", "/root/reference")

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    """make -C oracle ref (only where the reference's sources exist; elsewhere the prebuilt files are used)."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "main.cpp")):
        if force:
            subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "ref", "REF=" + REFERENCE_ROOT])
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available():
    return os.path.exists(_LIB_PATH) or build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_epsilon2.restype = C.c_double
        L.ref_gravitational_constant.restype = C.c_double
        L.ref_convert_to_earth_days.restype = C.c_double
        L.ref_convert_to_earth_days.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.ref_configure.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_double] + [C.c_int] * 10
        L.ref_config_storage_size.restype = C.c_uint32
        L.ref_config_stack_size.restype = C.c_uint32
        L.ref_naive_accel.argtypes = [C.c_int, C.c_uint32] + [_dp] * 7
        L.ref_energy.argtypes = [C.c_uint32] + [_dp] * 8
        L.ref_tree_create.restype = C.c_void_p
        L.ref_tree_create.argtypes = [C.c_int]
        L.ref_tree_destroy.argtypes = [C.c_void_p]
        L.ref_tree_aabb.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_tree_build.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_tree_num_nodes.restype = C.c_uint32
        L.ref_tree_num_nodes.argtypes = [C.c_void_p]
        L.ref_tree_get_aabb.argtypes = [C.c_void_p, _dp]
        for name, rt in [("body_of_node", _up), ("body_count", _up), ("octants", _up), ("is_leaf", C.POINTER(C.c_int)),
                         ("sum_masses", _dp), ("com_x", _dp), ("com_y", _dp), ("com_z", _dp), ("edge", _dp),
                         ("sorted_bodies", _up)]:
            f = getattr(L, "ref_tree_" + name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        L.ref_tree_canonical.argtypes = [C.c_void_p, _up, _u64p, _u64p, _up, _up, _up] + [_dp] * 8
        L.ref_bh_accel.argtypes = [C.c_uint32] + [_dp] * 7 + [_up]
        L.ref_sim_run.restype = C.c_void_p
        L.ref_sim_run.argtypes = [C.c_int, C.c_uint32] + [_dp] * 7 + [C.c_double] * 3 + [C.c_char_p]
        L.ref_sim_run_csv.restype = C.c_void_p
        L.ref_sim_run_csv.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int] + [C.c_double] * 3 + [C.c_char_p]
        L.ref_sim_destroy.argtypes = [C.c_void_p]
        L.ref_sim_num_bodies.restype = C.c_uint32
        L.ref_sim_num_bodies.argtypes = [C.c_void_p]
        L.ref_sim_num_snapshots.restype = C.c_uint32
        L.ref_sim_num_snapshots.argtypes = [C.c_void_p]
        L.ref_sim_get.argtypes = [C.c_void_p, C.c_uint32, C.c_int, _dp]
        L.ref_sim_energy.argtypes = [C.c_void_p, C.c_uint32, _dp]
        L.ref_sim_write_output.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _check(rc):
    if rc:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())


def configure(n, storage_param=16, stack_param=16, theta=1.05, block_size=64, opt_stage=2, sort_bodies=True,
              wg_size_barnes_hut=64, num_wi_octree=640, num_wi_top_octree=1024, num_wi_AABB=1024, num_wi_com=1024,
              max_level_top_octree=7, energy=False):
    """The configuration:: setters src/main.cpp:122-159 calls (defaults = Configuration.cpp:5-22)."""
    lib().ref_configure(n, storage_param, stack_param, theta, block_size, opt_stage, int(sort_bodies),
                        wg_size_barnes_hut, num_wi_octree, num_wi_top_octree, num_wi_AABB, num_wi_com,
                        max_level_top_octree, int(energy))
    return lib().ref_config_storage_size(), lib().ref_config_stack_size()


def gravitational_constant():
    return lib().ref_gravitational_constant()


def epsilon2():
    return lib().ref_epsilon2()


def max_threads():
    return lib().ref_max_threads()


def set_threads(n):
    lib().ref_set_threads(n)


def convert_to_earth_days(text):
    failed = C.c_int()
    v = lib().ref_convert_to_earth_days(text.encode(), C.byref(failed))
    if failed.value:
        raise ValueError(lib().ref_last_error().decode())
    return v


def naive_accel(m, x, y, z, opt_stage=0, block_size=64):
    """NaiveAlgorithm::computeAccelerations_opt_{0,1,2} of the reference itself."""
    m, pm = _d(m); x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    n = x.shape[0]
    configure(n, block_size=block_size, opt_stage=opt_stage)
    ax = np.zeros(n); ay = np.zeros(n); az = np.zeros(n)
    _check(lib().ref_naive_accel(opt_stage, n, pm, px, py, pz, ax.ctypes.data_as(_dp), ay.ctypes.data_as(_dp),
                                 az.ctypes.data_as(_dp)))
    return ax, ay, az


def energy(m, x, y, z, vx, vy, vz):
    """nBodyAlgorithm::computeEnergy -> (kinetic, potential, total, virial)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m, x, y, z, vx, vy, vz)]
    n = arrs[0].shape[0]
    configure(n)
    out = np.zeros(4)
    _check(lib().ref_energy(n, *[a.ctypes.data_as(_dp) for a in arrs], out.ctypes.data_as(_dp)))
    return out


class Tree:
    """The reference's octree built by its own buildOctree (AABB, insertion, centre of mass, in-order sort).

    builder: "subtrees" (ParallelOctreeTopDownSubtrees, the default) or "synchronized" (ParallelOctreeTopDownSynchronized).
    """

    def __init__(self, m, x, y, z, builder="subtrees", storage_param=16, stack_param=16, **cfg):
        L = lib()
        self.m, self._pm = _d(m)
        self.x, self._px = _d(x)
        self.y, self._py = _d(y)
        self.z, self._pz = _d(z)
        self.N = self.x.shape[0]
        self.S, _ = configure(self.N, storage_param, stack_param, **cfg)
        self._h = C.c_void_p(L.ref_tree_create({"subtrees": 0, "synchronized": 1}[builder]))
        if not self._h:
            raise RuntimeError("reference: " + L.ref_last_error().decode())
        _check(L.ref_tree_build(self._h, self._pm, self._px, self._py, self._pz))
        self.num_nodes = L.ref_tree_num_nodes(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_tree_destroy(self._h)
            self._h = None

    def _arr(self, name, dtype, n):
        p = getattr(lib(), "ref_tree_" + name)(self._h)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)

    def aabb(self):
        out = np.zeros(7)
        lib().ref_tree_get_aabb(self._h, out.ctypes.data_as(_dp))
        return out  # min xyz, max xyz, edge

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
        """Same record layout and order as oracle.Tree.canonical()."""
        n = self.num_nodes
        out = dict(depth=np.zeros(n, np.uint32), path_hi=np.zeros(n, np.uint64), path_lo=np.zeros(n, np.uint64),
                   kind=np.zeros(n, np.uint32), body=np.zeros(n, np.uint32), count=np.zeros(n, np.uint32))
        names = ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz")
        for k in names:
            out[k] = np.zeros(n)
        lib().ref_tree_canonical(self._h, out["depth"].ctypes.data_as(_up), out["path_hi"].ctypes.data_as(_u64p),
                                 out["path_lo"].ctypes.data_as(_u64p), out["kind"].ctypes.data_as(_up),
                                 out["body"].ctypes.data_as(_up), out["count"].ctypes.data_as(_up),
                                 *[out[k].ctypes.data_as(_dp) for k in names])
        return out


def aabb(x, y, z, num_wi_AABB=1024):
    """BarnesHutOctree::computeMinMaxValuesAABB -> min xyz, max xyz, edge."""
    x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    n = x.shape[0]
    configure(n, num_wi_AABB=num_wi_AABB)
    h = C.c_void_p(lib().ref_tree_create(0))
    out = np.zeros(7)
    try:
        _check(lib().ref_tree_aabb(h, px, py, pz, out.ctypes.data_as(_dp)))
    finally:
        lib().ref_tree_destroy(h)
    return out


def bh_accel(m, x, y, z, theta, sort_bodies=True, storage_param=16, stack_param=16, **cfg):
    """buildOctree + BarnesHutAlgorithm::computeAccelerations as the reference's time loop calls them."""
    m, pm = _d(m); x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    n = x.shape[0]
    configure(n, storage_param, stack_param, theta=theta, sort_bodies=sort_bodies, **cfg)
    ax = np.zeros(n); ay = np.zeros(n); az = np.zeros(n)
    nodes = C.c_uint32()
    _check(lib().ref_bh_accel(n, pm, px, py, pz, ax.ctypes.data_as(_dp), ay.ctypes.data_as(_dp),
                              az.ctypes.data_as(_dp), C.byref(nodes)))
    return ax, ay, az, nodes.value


def _collect(h, write_output):
    L = lib()
    n = L.ref_sim_num_bodies(h)
    ns = L.ref_sim_num_snapshots(h)
    keys = ("px", "py", "pz", "vx", "vy", "vz", "anorm")
    out = {k: np.zeros((ns, n)) for k in keys}
    en = np.zeros((ns, 4))
    for s in range(ns):
        for w, k in enumerate(keys):
            if L.ref_sim_get(h, s, w, out[k][s].ctypes.data_as(_dp)):
                raise RuntimeError("reference: snapshot %d of %s missing" % (s, k))
        L.ref_sim_energy(h, s, en[s].ctypes.data_as(_dp))
    out["energy"] = en
    out["n_snap"] = ns
    if write_output:
        _check(L.ref_sim_write_output(h))
    return out


def simulate(algorithm, m, x, y, z, vx, vy, vz, dt, t_end, vs, theta=1.05, energy=False, sort_bodies=True,
             storage_param=16, stack_param=16, opt_stage=2, block_size=64, output_directory=None, **cfg):
    """{Naive,BarnesHut}Algorithm::startSimulation of the reference; returns its snapshot maps as arrays.
    With output_directory, also runs generateParaViewOutput into <output_directory>/<ctime>/."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (m, x, y, z, vx, vy, vz)]
    n = arrs[0].shape[0]
    configure(n, storage_param, stack_param, theta=theta, block_size=block_size, opt_stage=opt_stage,
              sort_bodies=sort_bodies, energy=energy, **cfg)
    L = lib()
    h = L.ref_sim_run({"naive": 0, "BarnesHut": 1}[algorithm], n, *[a.ctypes.data_as(_dp) for a in arrs], dt, t_end, vs,
                      (output_directory or ".").encode())
    if not h:
        raise RuntimeError("reference: " + L.ref_last_error().decode())
    h = C.c_void_p(h)
    try:
        return _collect(h, output_directory is not None)
    finally:
        L.ref_sim_destroy(h)


def simulate_csv(algorithm, csv_path, dt, t_end, vs, theta=1.05, energy=False, sort_bodies=True, storage_param=16,
                 stack_param=16, opt_stage=2, block_size=64, output_directory=None, **cfg):
    """The same with the bodies read by the reference's InputParser (names / classes kept for the writers)."""
    L = lib()
    configure(1, storage_param, stack_param, theta=theta, block_size=block_size, opt_stage=opt_stage,
              sort_bodies=sort_bodies, energy=energy, **cfg)
    h = L.ref_sim_run_csv({"naive": 0, "BarnesHut": 1}[algorithm], csv_path.encode(), storage_param, stack_param, dt,
                          t_end, vs, (output_directory or ".").encode())
    if not h:
        raise RuntimeError("reference: " + L.ref_last_error().decode())
    h = C.c_void_p(h)
    try:
        return _collect(h, output_directory is not None)
    finally:
        L.ref_sim_destroy(h)


def run_executable(args, cwd=None, timeout=600):
    """The reference's own main() (src/main.cpp) with its cxxopts flags; returns the CompletedProcess."""
    return subprocess.run([EXE_PATH] + list(args), cwd=cwd, capture_output=True, text=True, timeout=timeout)
