"""ctypes binding of libsmash_b200.so (include/smash_b200.h).

The library is the product: if it is missing or no CUDA device is present, every compute call raises
``RuntimeError`` -- there is no CPU fallback and this module never imports anything from ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

GNP, GNS = 16, 8
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMASH_B200_LIB", os.path.join(_HERE, "libsmash_b200.so"))

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)

STRUCTURES = {"gr-a": 1, "gr-b": 2, "gr-c": 3, "gr-d": 4, "vic-a": 5}
JOBS_FUN = {"nse": 1, "kge": 2, "kge2": 3, "se": 4, "rmse": 5, "logarithmic": 6, "Crc": 7, "Cfp2": 8, "Cfp10": 9, "Cfp50": 10,
            "Cfp90": 11, "Erc": 12, "Elt": 13, "Epf": 14}
JREG_FUN = {"prior": 1, "smoothing": 2, "hard_smoothing": 3}
MAPPING = {"hyper-linear": 1, "hyper-polynomial": 2}


class SmashSetup(C.Structure):
    _fields_ = [
        ("structure", C.c_int32), ("dt", C.c_float), ("ntime_step", C.c_int32), ("nd", C.c_int32), ("ncpu", C.c_int32),
        ("sparse_storage", C.c_int32), ("save_qsim_domain", C.c_int32), ("save_net_prcp_domain", C.c_int32),
        ("njf", C.c_int32), ("jobs_fun", c_int_p), ("wjobs_fun", c_float_p),
        ("njr", C.c_int32), ("jreg_fun", c_int_p), ("wjreg_fun", c_float_p),
        ("wjreg", C.c_float), ("mapping", C.c_int32), ("denormalize_forward", C.c_int32), ("nhyper", C.c_int32),
        ("optimize_start_step", C.c_int32),
        ("optim_parameters", C.c_int32 * GNP), ("optim_states", C.c_int32 * GNS),
        ("lb_parameters", C.c_float * GNP), ("ub_parameters", C.c_float * GNP),
        ("lb_states", C.c_float * GNS), ("ub_states", C.c_float * GNS),
        ("wgauge", c_float_p), ("mask_event", c_int_p),
    ]


class SmashMesh(C.Structure):
    _fields_ = [
        ("dx", C.c_float), ("nrow", C.c_int32), ("ncol", C.c_int32), ("ng", C.c_int32), ("nac", C.c_int32),
        ("flwdir", c_int_p), ("flwacc", c_int_p), ("active_cell", c_int_p), ("local_active_cell", c_int_p),
        ("path", c_int_p), ("gauge_pos", c_int_p), ("rowcol_to_ind_sparse", c_int_p), ("area", c_float_p),
    ]


class SmashInputData(C.Structure):
    _fields_ = [
        ("qobs", c_float_p), ("prcp", c_float_p), ("pet", c_float_p), ("sparse_prcp", c_float_p),
        ("sparse_pet", c_float_p), ("descriptor", c_float_p), ("forcing_version", C.c_uint64), ("mean_prcp", c_float_p),
    ]


class SmashParameters(C.Structure):
    _fields_ = [("v", c_float_p * GNP)]


class SmashStates(C.Structure):
    _fields_ = [("v", c_float_p * GNS)]


class SmashOutput(C.Structure):
    _fields_ = [
        ("qsim", c_float_p), ("qsim_domain", c_float_p), ("sparse_qsim_domain", c_float_p),
        ("net_prcp_domain", c_float_p), ("sparse_net_prcp_domain", c_float_p),
        ("cost", C.c_float), ("cost_jobs", C.c_float), ("cost_jreg", C.c_float), ("fstates", SmashStates),
    ]


# every symbol include/smash_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "smash_b200_forward", "smash_b200_forward_b", "smash_b200_hyper_forward", "smash_b200_hyper_forward_b",
    "smash_b200_compute_multiple_run", "smash_b200_last_error", "smash_b200_version", "smash_b200_device_count",
    "smash_b200_set_device", "smash_b200_clear_cache", "smash_b200_set_option", "smash_b200_plan_create",
    "smash_b200_plan_destroy", "smash_b200_plan_set_forcing", "smash_b200_plan_set_fields",
    "smash_b200_plan_run_forward", "smash_b200_plan_run_gradient", "smash_b200_plan_run_hyper_gradient", "smash_b200_plan_get_qsim",
    "smash_b200_plan_get_gradient", "smash_b200_plan_checksum", "smash_b200_plan_info", "smash_b200_plan_order",
    "smash_b200_mesh_order", "smash_b200_mesh_chains", "smash_b200_mesh_tick_schedule", "smash_b200_flow_accumulation",
    "smash_b200_gauge_masks", "smash_b200_compute_mean_forcing", "smash_b200_adjust_interception_store", "smash_b200_mlp_forward", "smash_b200_mlp_create", "smash_b200_mlp_destroy",
    "smash_b200_mlp_run_forward", "smash_b200_mlp_run_backward", "smash_b200_plan_kernel_times", "smash_b200_plan_stat",
    "smash_b200_comm_unique_id", "smash_b200_comm_create", "smash_b200_comm_destroy", "smash_b200_comm_allreduce",
    "smash_b200_comm_allgather", "smash_b200_comm_rank", "smash_b200_comm_world", "smash_b200_comm_last_error",
]

_lib = None


def lib():
    """Load libsmash_b200.so (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  smash_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.smash_b200_last_error.restype = C.c_char_p
        L.smash_b200_version.restype = C.c_char_p
        L.smash_b200_set_option.argtypes = [C.c_char_p, C.c_longlong]
        L.smash_b200_plan_destroy.restype = None
        L.smash_b200_plan_destroy.argtypes = [C.c_void_p]
        L.smash_b200_clear_cache.restype = None
        L.smash_b200_plan_stat.restype = C.c_double
        L.smash_b200_plan_stat.argtypes = [C.c_void_p, C.c_char_p]
        L.smash_b200_mlp_destroy.restype = None
        L.smash_b200_mlp_destroy.argtypes = [C.c_void_p]
        L.smash_b200_comm_last_error.restype = C.c_char_p
        L.smash_b200_comm_destroy.restype = None
        L.smash_b200_comm_destroy.argtypes = [C.c_void_p]
        L.smash_b200_comm_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
        L.smash_b200_comm_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"libsmash_b200 error {rc}: {lib().smash_b200_last_error().decode()}")


def _fp(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def _f32(a, keep):
    """Fortran-contiguous float32 view of ``a`` (no copy when it already is one)."""
    b = np.asfortranarray(a, dtype=np.float32)
    keep.append(b)
    return b


def _i32(a, keep):
    b = np.asfortranarray(a, dtype=np.int32)
    keep.append(b)
    return b


class Packed:
    """ctypes structs + the NumPy arrays that back them (kept alive for the duration of the call)."""

    def __init__(self):
        self.keep = []


def pack_setup(setup, mesh, pk: Packed) -> SmashSetup:
    o = setup._optimize
    s = SmashSetup()
    if setup.structure not in STRUCTURES:
        raise ValueError(f"unknown structure {setup.structure!r}")
    # gr-a: every entry point.  gr-b, gr-c, gr-d, vic-a (md_forward_structure.f90:216-931): forward runs (forward,
    # compute_multiple_run, the plan API); the library answers SMASH_B200_EUNSUPPORTED to their adjoint and descriptor mappings
    s.structure = STRUCTURES[setup.structure]
    s.dt = float(setup.dt)
    s.ntime_step = int(setup._ntime_step)
    s.nd = int(setup._nd)
    s.ncpu = int(setup._ncpu)
    s.sparse_storage = int(bool(setup.sparse_storage))
    s.save_qsim_domain = int(bool(setup.save_qsim_domain))
    s.save_net_prcp_domain = int(bool(setup.save_net_prcp_domain))
    jf = [str(x).strip() for x in np.atleast_1d(o.jobs_fun)][: int(o.njf)]
    for name in jf:
        if name not in JOBS_FUN:
            raise RuntimeError(f"jobs_fun {name!r} is not implemented by smash_b200")
    jfc = _i32(np.array([JOBS_FUN[x] for x in jf], dtype=np.int32), pk.keep)
    wjf = _f32(np.atleast_1d(o.wjobs_fun)[: len(jf)], pk.keep)
    s.njf, s.jobs_fun, s.wjobs_fun = len(jf), _ip(jfc), _fp(wjf)
    jr = [str(x).strip() for x in np.atleast_1d(o.jreg_fun)][: int(o.njr)]
    for name in jr:
        if name not in JREG_FUN:
            raise RuntimeError(f"jreg_fun {name!r} is not implemented by smash_b200")
    jrc = _i32(np.array([JREG_FUN[x] for x in jr], dtype=np.int32), pk.keep)
    wjr = _f32(np.atleast_1d(o.wjreg_fun)[: len(jr)], pk.keep)
    s.njr, s.jreg_fun, s.wjreg_fun = len(jr), _ip(jrc), _fp(wjr)
    s.wjreg = float(o.wjreg)
    s.mapping = MAPPING.get(str(o.mapping).strip(), 0)
    s.denormalize_forward = int(bool(o.denormalize_forward))
    s.nhyper = int(o.nhyper)
    s.optimize_start_step = int(o.optimize_start_step)
    if any(JOBS_FUN[x] >= 7 for x in jf) and mesh.ng > 0:                   # signature objectives read the event mask
        s.mask_event = _ip(_i32(o.mask_event, pk.keep))
    s.optim_parameters[:] = [int(x) for x in o.optim_parameters]
    s.optim_states[:] = [int(x) for x in o.optim_states]
    s.lb_parameters[:] = [float(x) for x in o.lb_parameters]
    s.ub_parameters[:] = [float(x) for x in o.ub_parameters]
    s.lb_states[:] = [float(x) for x in o.lb_states]
    s.ub_states[:] = [float(x) for x in o.ub_states]
    wg = _f32(np.atleast_1d(o.wgauge), pk.keep) if mesh.ng > 0 else None
    s.wgauge = _fp(wg)
    return s


def pack_mesh(mesh, setup, pk: Packed) -> SmashMesh:
    m = SmashMesh()
    m.dx = float(mesh.dx)
    m.nrow, m.ncol, m.ng, m.nac = int(mesh.nrow), int(mesh.ncol), int(mesh.ng), int(mesh.nac)
    # The converted arrays are cached on the mesh object.  The key holds the source arrays themselves (identity compared, and
    # kept alive so that an id cannot be recycled); rebinding any of them -- e.g. mesh._local_active_cell for a basin shard --
    # makes a new image.  In-place edits of a keyed array are not seen: rebind or delete mesh._b200_cache.
    cache = getattr(mesh, "_b200_cache", None)
    srcs = (mesh.flwdir, mesh.flwacc, mesh.path, mesh.active_cell, getattr(mesh, "_local_active_cell", None),
            getattr(mesh, "gauge_pos", None), getattr(mesh, "area", None))
    key = tuple(srcs)
    if cache is None or len(cache[0]) != len(key) or any(x is not y for x, y in zip(cache[0], key)):
        arrs = {}
        arrs["flwdir"] = np.asfortranarray(mesh.flwdir, dtype=np.int32)
        arrs["flwacc"] = np.asfortranarray(mesh.flwacc, dtype=np.int32)
        arrs["active_cell"] = np.asfortranarray(mesh.active_cell, dtype=np.int32)
        lac = getattr(mesh, "_local_active_cell", None)
        arrs["local_active_cell"] = None if lac is None else np.asfortranarray(lac, dtype=np.int32)
        # Python side is 0-based (f90wrap index handler); the C ABI takes the Fortran memory image
        arrs["path"] = np.asfortranarray(np.asarray(mesh.path, dtype=np.int32) + 1)
        arrs["gauge_pos"] = np.asfortranarray(np.asarray(mesh.gauge_pos, dtype=np.int32) + 1) if mesh.ng > 0 else None
        arrs["area"] = np.asfortranarray(mesh.area, dtype=np.float32) if mesh.ng > 0 else None
        cache = (key, arrs)
        try:
            mesh._b200_cache = cache
        except AttributeError:
            pass
    a = cache[1]
    pk.keep.append(a)
    m.flwdir, m.flwacc, m.active_cell = _ip(a["flwdir"]), _ip(a["flwacc"]), _ip(a["active_cell"])
    m.local_active_cell = _ip(a["local_active_cell"])
    m.path, m.gauge_pos, m.area = _ip(a["path"]), _ip(a["gauge_pos"]), _fp(a["area"])
    rts = getattr(mesh, "_rowcol_to_ind_sparse", None)
    m.rowcol_to_ind_sparse = _ip(_i32(rts, pk.keep)) if rts is not None else None
    return m


def pack_input(input_data, setup, mesh, pk: Packed) -> SmashInputData:
    i = SmashInputData()
    if mesh.ng > 0 and getattr(input_data, "qobs", None) is not None:
        i.qobs = _fp(_f32(input_data.qobs, pk.keep))
    if setup.sparse_storage:
        i.sparse_prcp = _fp(_f32(input_data.sparse_prcp, pk.keep))
        i.sparse_pet = _fp(_f32(input_data.sparse_pet, pk.keep))
    else:
        i.prcp = _fp(_f32(input_data.prcp, pk.keep))
        i.pet = _fp(_f32(input_data.pet, pk.keep))
    if setup._nd > 0 and getattr(input_data, "descriptor", None) is not None:
        i.descriptor = _fp(_f32(input_data.descriptor, pk.keep))
    i.forcing_version = int(getattr(input_data, "_forcing_version", 0))
    if mesh.ng > 0 and getattr(input_data, "mean_prcp", None) is not None:
        i.mean_prcp = _fp(_f32(input_data.mean_prcp, pk.keep))
    return i


def _planes(obj, names, struct, pk: Packed, writeback: list | None):
    """Bind every field of a ParametersDT / StatesDT-like object.  Arrays that are not float32
    Fortran-contiguous are copied; `writeback` collects (obj, name, copy) to restore in-place semantics."""
    s = struct()
    for k, name in enumerate(names):
        a = getattr(obj, name, None)
        if a is None:
            continue
        b = np.asfortranarray(a, dtype=np.float32)
        if b is not a and writeback is not None:
            writeback.append((obj, name, b))
        pk.keep.append(b)
        s.v[k] = _fp(b)
    return s


PARAM_NAMES = ("ci", "cp", "beta", "cft", "cst", "alpha", "exc", "b", "cusl1", "cusl2", "clsl", "ks", "ds", "dsm",
               "ws", "lr")
STATE_NAMES = ("hi", "hp", "hft", "hst", "husl1", "husl2", "hlsl", "hlr")


def pack_parameters(p, pk, writeback=None):
    return _planes(p, PARAM_NAMES, SmashParameters, pk, writeback)


def pack_states(s, pk, writeback=None):
    return _planes(s, STATE_NAMES, SmashStates, pk, writeback)


def pack_output(output, setup, mesh, pk: Packed, writeback: list):
    o = SmashOutput()
    for name in ("qsim", "qsim_domain", "sparse_qsim_domain", "net_prcp_domain", "sparse_net_prcp_domain"):
        a = getattr(output, name, None)
        if a is None or (name == "qsim" and mesh.ng == 0):
            continue
        b = np.asfortranarray(a, dtype=np.float32)
        if b is not a:
            writeback.append((output, name, b))
        pk.keep.append(b)
        setattr(o, name, _fp(b))
    o.fstates = pack_states(output.fstates, pk, writeback)
    return o


def finish_output(o: SmashOutput, output, writeback):
    for obj, name, arr in writeback:
        dst = getattr(obj, name)
        if isinstance(dst, np.ndarray) and dst.shape == arr.shape:
            dst[...] = arr
        else:
            setattr(obj, name, arr)
    output.cost = np.float32(o.cost)
    output.cost_jobs = np.float32(o.cost_jobs)
    output.cost_jreg = np.float32(o.cost_jreg)
