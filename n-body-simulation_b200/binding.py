"""ctypes binding of include/nbody_b200.h (tests, bench.py and smoke() call the product through this, i.e. through the
C ABI).  No compute happens in Python and there is no fallback: if libnbody_b200.so is missing or no B200 is visible the
calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NB_LIB: developer override used by the A/B scripts under tools/ to time an earlier build of the library
_LIB = os.environ.get("NB_LIB") or os.path.join(_HERE, "libnbody_b200.so")

NB_COMM_ID_BYTES = 128
NB_T_COUNT = 10
TIMER_NAMES = ["Acceleration Kernel Time", "Leapfrog Part 1", "Leapfrog Part 2", "AABB creation",
               "Sort bodies for subtrees", "Build subtrees", "Compute center of mass", "Octree creation", "Energy",
               "Allgather"]

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class NBConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("G", C.c_double), ("epsilon2", C.c_double),
                ("theta", C.c_double), ("block_size", C.c_int32), ("opt_stage", C.c_int32), ("sort_bodies", C.c_int32),
                ("wg_size_barnes_hut", C.c_int32), ("storage_size_param", C.c_int32), ("stack_size_param", C.c_int32),
                ("num_wi_aabb", C.c_int32), ("num_wi_octree", C.c_int32), ("num_wi_top_octree", C.c_int32),
                ("num_wi_com", C.c_int32), ("max_level_top_octree", C.c_int32), ("precise_rsqrt", C.c_int32),
                ("world_size", C.c_int32), ("rank", C.c_int32), ("reserved", C.c_int32 * 8)]


class NBTreeInfo(C.Structure):
    _fields_ = [("num_bodies", C.c_uint64), ("num_nodes_materialised", C.c_uint64), ("num_internal", C.c_uint64),
                ("num_nodes_canonical", C.c_uint64), ("max_depth", C.c_uint32), ("reserved", C.c_uint32),
                ("aabb_min", C.c_double * 3), ("aabb_max", C.c_double * 3), ("aabb_edge", C.c_double)]


class NBodyError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("nbody_b200 status %d: %s" % (status, text))
        self.status = status


# every symbol include/nbody_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "nb_config_default", "nb_abi_version", "nb_status_string", "nb_create", "nb_destroy", "nb_last_error",
    "nb_synchronize", "nb_device_name", "nb_set_theta", "nb_set_block_size", "nb_set_sort_bodies",
    "nb_set_precise_rsqrt", "nb_set_bodies", "nb_set_positions", "nb_num_bodies", "nb_naive_accel", "nb_bh_build",
    "nb_bh_accel", "nb_bh_accel_range", "nb_leapfrog_part1", "nb_leapfrog_part2", "nb_leapfrog_part2_part1", "nb_advance", "nb_energy",
    "nb_get_positions", "nb_get_velocities", "nb_get_accelerations", "nb_get_acceleration_norms",
    "nb_op_naive_accelerations", "nb_op_barnes_hut_accelerations", "nb_bh_tree_info", "nb_bh_aabb",
    "nb_bh_export_canonical", "nb_bh_sorted_bodies", "nb_bh_enable_stats", "nb_bh_get_stats",
    "nb_util_group_by_subtree", "nb_enable_timers", "nb_get_timers", "nb_timer_name", "nb_comm_get_unique_id",
    "nb_comm_init", "nb_comm_p2p_enabled", "nb_slice_bounds", "nb_event_record", "nb_event_elapsed_ms", "nb_measure_fp64_peak", "nb_launch_count", "nb_device_pointers",
]

_lib = None


def library_path():
    return _LIB


def load_library():
    """dlopen libnbody_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise NBodyError(-2, "libnbody_b200.so not built: run `python n-body-simulation_b200/build.py`")
    L = C.CDLL(_LIB)
    vp = C.c_void_p
    L.nb_config_default.argtypes = [C.POINTER(NBConfig)]
    L.nb_config_default.restype = None
    L.nb_status_string.restype = C.c_char_p
    L.nb_status_string.argtypes = [C.c_int]
    L.nb_create.argtypes = [C.POINTER(NBConfig), C.POINTER(vp)]
    L.nb_destroy.argtypes = [vp]
    L.nb_destroy.restype = None
    L.nb_last_error.argtypes = [vp]
    L.nb_last_error.restype = C.c_char_p
    L.nb_synchronize.argtypes = [vp]
    L.nb_device_name.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.nb_set_theta.argtypes = [vp, C.c_double]
    L.nb_set_block_size.argtypes = [vp, C.c_int]
    L.nb_set_sort_bodies.argtypes = [vp, C.c_int]
    L.nb_set_precise_rsqrt.argtypes = [vp, C.c_int]
    L.nb_set_bodies.argtypes = [vp, C.c_uint64] + [_dp] * 7
    L.nb_set_positions.argtypes = [vp, _dp, _dp, _dp]
    L.nb_num_bodies.argtypes = [vp]
    L.nb_num_bodies.restype = C.c_uint64
    for f in ("nb_naive_accel", "nb_bh_build", "nb_bh_accel"):
        getattr(L, f).argtypes = [vp]
    L.nb_bh_accel_range.argtypes = [vp, C.c_uint64, C.c_uint64]
    for f in ("nb_leapfrog_part1", "nb_leapfrog_part2", "nb_leapfrog_part2_part1"):
        getattr(L, f).argtypes = [vp, C.c_double]
    L.nb_advance.argtypes = [vp, C.c_int, C.c_double, C.c_uint32, _dp]
    L.nb_energy.argtypes = [vp, _dp]
    for f in ("nb_get_positions", "nb_get_velocities", "nb_get_accelerations"):
        getattr(L, f).argtypes = [vp, _dp, _dp, _dp]
    L.nb_get_acceleration_norms.argtypes = [vp, _dp]
    L.nb_op_naive_accelerations.argtypes = [vp, C.c_uint64] + [_dp] * 7
    L.nb_op_barnes_hut_accelerations.argtypes = [vp, C.c_uint64] + [_dp] * 7
    L.nb_bh_tree_info.argtypes = [vp, C.POINTER(NBTreeInfo)]
    L.nb_bh_aabb.argtypes = [vp, _dp]
    L.nb_bh_export_canonical.argtypes = [vp, _u32p, _u64p, _u64p, _u32p, _u32p, _u32p] + [_dp] * 8
    L.nb_bh_sorted_bodies.argtypes = [vp, _u32p]
    L.nb_bh_enable_stats.argtypes = [vp, C.c_int]
    L.nb_bh_get_stats.argtypes = [vp, _u64p, _u64p, _u32p]
    L.nb_util_group_by_subtree.argtypes = [vp, C.c_uint32, _u32p, C.c_uint32, _u32p, _u32p, _u32p, _u32p, _u32p]
    L.nb_enable_timers.argtypes = [vp, C.c_int]
    L.nb_get_timers.argtypes = [vp, _dp]
    L.nb_timer_name.argtypes = [C.c_int]
    L.nb_timer_name.restype = C.c_char_p
    L.nb_comm_get_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.nb_comm_init.argtypes = [vp, C.POINTER(C.c_uint8), C.c_int, C.c_int]
    L.nb_comm_p2p_enabled.argtypes = [vp]
    L.nb_slice_bounds.argtypes = [C.c_uint64, C.c_int, C.c_int, _u64p, _u64p]
    L.nb_slice_bounds.restype = None
    L.nb_event_record.argtypes = [vp, C.c_int]
    L.nb_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.nb_measure_fp64_peak.argtypes = [vp, _dp]
    L.nb_launch_count.argtypes = [vp]
    L.nb_launch_count.restype = C.c_uint64
    L.nb_device_pointers.argtypes = [vp, C.POINTER(C.c_void_p)]
    _lib = L
    return L


def default_config(**overrides):
    cfg = NBConfig()
    load_library().nb_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if k == "ipt":
            cfg.reserved[0] = int(v)
        elif k == "unfused_advance":
            cfg.reserved[1] = int(v)
        elif k == "naive_segments":
            cfg.reserved[4] = int(v)
        elif k == "static_slices":
            cfg.reserved[5] = int(v)
        elif k == "sort_variant":
            cfg.reserved[6] = int(v)
        elif k == "com_variant":
            cfg.reserved[7] = int(v)
        elif k == "walk_variant":
            cfg.reserved[3] = int(v)
        elif k == "naive_variant":
            cfg.reserved[2] = int(v)
        else:
            setattr(cfg, k, v)
    return cfg


def slice_bounds(n, world, rank):
    b, e = C.c_uint64(), C.c_uint64()
    load_library().nb_slice_bounds(n, world, rank, C.byref(b), C.byref(e))
    return b.value, e.value


def comm_unique_id():
    buf = (C.c_uint8 * NB_COMM_ID_BYTES)()
    rc = load_library().nb_comm_get_unique_id(buf)
    if rc:
        raise NBodyError(rc, "nb_comm_get_unique_id failed (libnccl.so.2 missing?)")
    return bytes(buf)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One nb_ctx: a CUDA device + stream (+ NCCL communicator) owning the simulation state."""

    def __init__(self, cfg=None, **overrides):
        self.L = load_library()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        h = C.c_void_p()
        rc = self.L.nb_create(C.byref(self.cfg), C.byref(h))
        if rc:
            raise NBodyError(rc, self.L.nb_status_string(rc).decode())
        self.h = h
        self.n = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.nb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise NBodyError(rc, self.L.nb_last_error(self.h).decode() or self.L.nb_status_string(rc).decode())

    # ---- state -------------------------------------------------------------------------------------------------
    def set_bodies(self, m, x, y, z, vx=None, vy=None, vz=None):
        arrs = [_f64(a) for a in (m, x, y, z)]
        vel = [None if a is None else _f64(a) for a in (vx, vy, vz)]
        self.n = arrs[0].shape[0]
        self._ck(self.L.nb_set_bodies(self.h, self.n, *[_d(a) for a in arrs], *[_d(a) for a in vel]))

    def set_positions(self, x, y, z):
        arrs = [_f64(a) for a in (x, y, z)]
        self._ck(self.L.nb_set_positions(self.h, *[_d(a) for a in arrs]))

    def _get3(self, fn):
        out = [np.empty(self.n) for _ in range(3)]
        self._ck(fn(self.h, *[_d(a) for a in out]))
        return out

    def positions(self):
        return self._get3(self.L.nb_get_positions)

    def velocities(self):
        return self._get3(self.L.nb_get_velocities)

    def accelerations(self):
        return self._get3(self.L.nb_get_accelerations)

    def acceleration_norms(self):
        out = np.empty(self.n)
        self._ck(self.L.nb_get_acceleration_norms(self.h, _d(out)))
        return out

    def synchronize(self):
        self._ck(self.L.nb_synchronize(self.h))

    def device_name(self):
        buf = C.create_string_buffer(256)
        self._ck(self.L.nb_device_name(self.h, buf, 256))
        return buf.value.decode()

    def device_pointers(self):
        p = (C.c_void_p * 10)()
        self._ck(self.L.nb_device_pointers(self.h, p))
        return dict(zip(("x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az", "m"), [int(v or 0) for v in p]))

    # ---- knobs ----------------------------------------------------------------------------------------------------
    def set_theta(self, theta):
        self._ck(self.L.nb_set_theta(self.h, theta))

    def set_block_size(self, bs):
        self._ck(self.L.nb_set_block_size(self.h, bs))

    def set_precise_rsqrt(self, p):
        self._ck(self.L.nb_set_precise_rsqrt(self.h, int(p)))

    # ---- operators --------------------------------------------------------------------------------------------------
    def naive_accel(self):
        self._ck(self.L.nb_naive_accel(self.h))

    def bh_build(self):
        self._ck(self.L.nb_bh_build(self.h))

    def bh_accel(self):
        self._ck(self.L.nb_bh_accel(self.h))

    def bh_accel_range(self, begin, end):
        self._ck(self.L.nb_bh_accel_range(self.h, C.c_uint64(begin), C.c_uint64(end)))

    def leapfrog_part1(self, dt):
        self._ck(self.L.nb_leapfrog_part1(self.h, dt))

    def leapfrog_part2(self, dt):
        self._ck(self.L.nb_leapfrog_part2(self.h, dt))

    def advance(self, algorithm, dt, nsteps, timers=False):
        """nsteps complete leapfrog steps (part 1, forces, part 2 each); algorithm: "naive" or "BarnesHut"."""
        ms = (C.c_double * 16)() if timers else None
        self._ck(self.L.nb_advance(self.h, {"naive": 0, "BarnesHut": 1}[algorithm], dt, nsteps, ms))
        return list(ms) if timers else None

    def leapfrog_part2_part1(self, dt):
        self._ck(self.L.nb_leapfrog_part2_part1(self.h, dt))

    def energy(self):
        out = np.zeros(4)
        self._ck(self.L.nb_energy(self.h, _d(out)))
        return out

    def op_naive_accelerations(self, m, x, y, z, out=None):
        arrs = [_f64(a) for a in (m, x, y, z)]
        n = arrs[0].shape[0]
        self.n = n
        out = out if out is not None else [np.empty(n) for _ in range(3)]
        self._ck(self.L.nb_op_naive_accelerations(self.h, n, *[_d(a) for a in arrs], *[_d(a) for a in out]))
        return out

    def op_barnes_hut_accelerations(self, m, x, y, z, out=None):
        arrs = [_f64(a) for a in (m, x, y, z)]
        n = arrs[0].shape[0]
        self.n = n
        out = out if out is not None else [np.empty(n) for _ in range(3)]
        self._ck(self.L.nb_op_barnes_hut_accelerations(self.h, n, *[_d(a) for a in arrs], *[_d(a) for a in out]))
        return out

    # ---- Barnes-Hut inspection -----------------------------------------------------------------------------------------
    def bh_aabb(self):
        out = np.zeros(7)
        self._ck(self.L.nb_bh_aabb(self.h, _d(out)))
        return out

    def bh_tree_info(self):
        info = NBTreeInfo()
        self._ck(self.L.nb_bh_tree_info(self.h, C.byref(info)))
        return info

    def bh_export_canonical(self):
        n = self.bh_tree_info().num_nodes_canonical
        out = dict(depth=np.zeros(n, np.uint32), path_hi=np.zeros(n, np.uint64), path_lo=np.zeros(n, np.uint64),
                   kind=np.zeros(n, np.uint32), body=np.zeros(n, np.uint32), count=np.zeros(n, np.uint32))
        names = ("edge", "minx", "miny", "minz", "mass", "comx", "comy", "comz")
        for k in names:
            out[k] = np.zeros(n)
        self._ck(self.L.nb_bh_export_canonical(
            self.h, out["depth"].ctypes.data_as(_u32p), out["path_hi"].ctypes.data_as(_u64p),
            out["path_lo"].ctypes.data_as(_u64p), out["kind"].ctypes.data_as(_u32p), out["body"].ctypes.data_as(_u32p),
            out["count"].ctypes.data_as(_u32p), *[_d(out[k]) for k in names]))
        return out

    def bh_sorted_bodies(self):
        out = np.zeros(self.n, np.uint32)
        self._ck(self.L.nb_bh_sorted_bodies(self.h, out.ctypes.data_as(_u32p)))
        return out

    def bh_enable_stats(self, enable=True):
        self._ck(self.L.nb_bh_enable_stats(self.h, int(enable)))

    def bh_stats(self, per_body=False):
        tv, ta = C.c_uint64(), C.c_uint64()
        pb = np.zeros(self.n, np.uint32) if per_body else None
        self._ck(self.L.nb_bh_get_stats(self.h, C.byref(tv), C.byref(ta),
                                        pb.ctypes.data_as(_u32p) if per_body else None))
        return (tv.value, ta.value, pb) if per_body else (tv.value, ta.value)

    def util_group_by_subtree(self, subtree_of_body, node_count):
        s = np.ascontiguousarray(subtree_of_body, dtype=np.uint32)
        n = s.shape[0]
        counts = np.zeros(node_count, np.uint32)
        subtrees = np.zeros(node_count, np.uint32)
        start = np.zeros(node_count, np.uint32)
        sorted_bodies = np.zeros(n, np.uint32)
        cnt = C.c_uint32()
        self._ck(self.L.nb_util_group_by_subtree(self.h, n, s.ctypes.data_as(_u32p), node_count,
                                                 counts.ctypes.data_as(_u32p), subtrees.ctypes.data_as(_u32p),
                                                 C.byref(cnt), start.ctypes.data_as(_u32p),
                                                 sorted_bodies.ctypes.data_as(_u32p)))
        return counts, subtrees[:cnt.value], cnt.value, start[:cnt.value], sorted_bodies

    # ---- timers / measurement ---------------------------------------------------------------------------------------------
    def enable_timers(self, enable=True):
        self._ck(self.L.nb_enable_timers(self.h, int(enable)))

    def timers(self):
        ms = np.zeros(NB_T_COUNT)
        self._ck(self.L.nb_get_timers(self.h, _d(ms)))
        return dict(zip(TIMER_NAMES, ms.tolist()))

    def event_record(self, slot):
        self._ck(self.L.nb_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        v = C.c_double()
        self._ck(self.L.nb_event_elapsed_ms(self.h, a, b, C.byref(v)))
        return v.value

    def measure_fp64_peak(self):
        v = C.c_double()
        self._ck(self.L.nb_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def launch_count(self):
        return int(self.L.nb_launch_count(self.h))

    # ---- multi-GPU ------------------------------------------------------------------------------------------------------------
    def comm_init(self, unique_id, world_size, rank):
        buf = (C.c_uint8 * NB_COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self.L.nb_comm_init(self.h, buf, world_size, rank))

    def p2p_enabled(self):
        return bool(self.L.nb_comm_p2p_enabled(self.h))
