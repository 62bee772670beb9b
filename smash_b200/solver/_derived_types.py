"""Host-side mirror of the reference's wrapped derived types.

The reference generates these classes with f90wrap from ``smash/solver/derived_type/mwd_*.f90``; here they
are plain Python objects holding Fortran-ordered NumPy arrays with the same attribute names, dtypes,
shapes and defaults, so that code written against ``smash.solver._mwd_*`` keeps working:

* ``SetupDT`` / ``Optimize_SetupDT``  mwd_setup.f90:57-228 (private fields appear with a leading underscore,
  ``f90wrap_utils/finalize_f90wrap.py:49-120``)
* ``MeshDT``                         mwd_mesh.f90:45-120 (``path`` and ``gauge_pos`` are 0-based on the Python
  side like the f90wrap index handlers, ``_f90wrap_decorator.py:72-106``)
* ``Input_DataDT``                   mwd_input_data.f90:32-110
* ``ParametersDT`` / ``Hyper_ParametersDT``  mwd_parameters.f90:58-230
* ``StatesDT`` / ``Hyper_StatesDT``          mwd_states.f90:49-170
* ``OutputDT``                       mwd_output.f90:36-130
"""
from __future__ import annotations

import copy as _copy

import numpy as np

GNP = 16  # md_constant.f90:32
GNS = 8   # md_constant.f90:33

# md_constant.f90:35-69
GPARAMETERS_NAME = ("ci", "cp", "beta", "cft", "cst", "alpha", "exc", "b", "cusl1", "cusl2", "clsl", "ks",
                    "ds", "dsm", "ws", "lr")
GSTATES_NAME = ("hi", "hp", "hft", "hst", "husl1", "husl2", "hlsl", "hlr")
# md_constant.f90:70-136
GLB_PARAMETERS = np.array([1e-6, 1e-6, 1e-6, 1e-6, 1e-6, 1e-6, -50.0, 1e-6, 1e-6, 1e-6, 1e-6, 1e-6, 1e-6, 1e-6,
                           1e-6, 1e-6], dtype=np.float32)
GUB_PARAMETERS = np.array([1e2, 1e3, 1e3, 1e3, 1e4, 0.999999, 50.0, 1e1, 2e3, 2e3, 2e3, 1e4, 0.999999, 30.0,
                           0.999999, 1e3], dtype=np.float32)
GLB_STATES = np.array([1e-6] * 8, dtype=np.float32)
GUB_STATES = np.array([0.999999] * 7 + [10000.0], dtype=np.float32)

# mwd_parameters.f90:150-167 / mwd_states.f90:117-126
_PARAMETERS_DEFAULT = dict(ci=1e-6, cp=200.0, beta=1000.0, cft=500.0, cst=500.0, alpha=0.9, exc=0.0, b=0.3,
                           cusl1=100.0, cusl2=500.0, clsl=2000.0, ks=20.0, ds=0.02, dsm=0.33, ws=0.8, lr=5.0)
_STATES_DEFAULT = dict(hi=0.01, hp=0.01, hft=0.01, hst=0.01, husl1=0.01, husl2=0.01, hlsl=0.01, hlr=0.000001)


def _farray(shape, value, dtype=np.float32):
    return np.full(shape, value, dtype=dtype, order="F")


def _coerce(cur, value):
    """What an f90wrap property setter does to ``value`` given the Fortran type behind the attribute: integer and logical scalars are
    converted to the component's type (``setup._ntime_step = 1440.0`` stores the integer 1440), an array component keeps its type
    and -- when the right-hand side is a scalar -- its storage (``parameters.lr = 5`` fills the plane).
    An array of another shape replaces the component (the Python mirror has no fixed allocation)."""
    if isinstance(cur, (bool, np.bool_)):
        return bool(value), False
    if isinstance(cur, (int, np.integer)) and not isinstance(value, (str, bytes)):
        return int(value), False
    if isinstance(cur, np.ndarray) and not isinstance(value, (str, bytes)):
        arr = np.asarray(value)
        if arr.shape == cur.shape:
            # kept as given (the packing layer converts to the solver's kinds; the float64 oracle tests rely on it)
            return value if isinstance(value, np.ndarray) else np.asfortranarray(arr, dtype=cur.dtype), False
        if arr.ndim == 0 and cur.dtype.kind in "iuf" and arr.dtype.kind in "iufb":
            cur[...] = arr
            return cur, True
        return value, False
    return value, False


class _DT:
    """copy() mirrors the f90wrap ``copy`` methods added by finalize_f90wrap.py; attribute assignment follows the typed
    f90wrap setters (_coerce)."""

    def __setattr__(self, name, value):
        cur = self.__dict__.get(name)
        if cur is not None:
            value, done = _coerce(cur, value)
            if done:
                return
        object.__setattr__(self, name, value)

    def copy(self):
        return _copy.deepcopy(self)


class Optimize_SetupDT(_DT):
    def __init__(self, ntime_step=0, nd=0, ng=0, mapping="...", njf=0, njr=0):
        self.algorithm = "..."
        self.jobs_fun = np.array(["..."] * njf, dtype="U20")
        self.wjobs_fun = np.zeros(njf, dtype=np.float32)
        self.wjreg = np.float32(0.0)
        self.jreg_fun = np.array(["..."] * njr, dtype="U20")
        self.wjreg_fun = np.ones(njr, dtype=np.float32)
        self.reg_descriptors_for_params = np.zeros((GNP, nd), dtype=np.int32, order="F")
        self.reg_descriptors_for_states = np.zeros((GNS, nd), dtype=np.int32, order="F")
        self.njf = njf
        self.njr = njr
        self.verbose = True
        self.mapping = mapping
        self.denormalize_forward = False
        self.nhyper = {"hyper-linear": 1 + nd, "hyper-polynomial": 1 + 2 * nd}.get(mapping, 0)
        self.optimize_start_step = 1
        self.maxiter = 100
        self.optim_parameters = np.zeros(GNP, dtype=np.int32)
        self.optim_states = np.zeros(GNS, dtype=np.int32)
        self.lb_parameters = GLB_PARAMETERS.copy()
        self.ub_parameters = GUB_PARAMETERS.copy()
        self.lb_states = GLB_STATES.copy()
        self.ub_states = GUB_STATES.copy()
        self.wgauge = np.full(ng, 1.0 / ng if ng else 0.0, dtype=np.float32)
        self.mask_event = np.zeros((ng, ntime_step), dtype=np.int32, order="F")


class SetupDT(_DT):
    def __init__(self, nd=0, ng=0):
        self.structure = "gr-a"
        self.dt = np.float32(3600.0)
        self.start_time = "..."
        self.end_time = "..."
        self.sparse_storage = False
        self.read_qobs = False
        self.qobs_directory = "..."
        self.read_prcp = False
        self.prcp_format = "tif"
        self.prcp_yyyymmdd_access = False
        self.prcp_conversion_factor = np.float32(1.0)
        self.prcp_directory = "..."
        self.read_pet = False
        self.pet_format = "tif"
        self.pet_conversion_factor = np.float32(1.0)
        self.pet_directory = "..."
        self.daily_interannual_pet = False
        self.mean_forcing = True
        self.read_descriptor = False
        self.descriptor_format = "tif"
        self.descriptor_directory = "..."
        self.descriptor_name = np.array(["..."] * nd, dtype="U20")
        self.save_qsim_domain = False
        self.save_net_prcp_domain = False
        self._ntime_step = 0
        self._nd = nd
        self._ncpu = 1
        self._parameters_name = np.array(GPARAMETERS_NAME, dtype="U10")
        self._states_name = np.array(GSTATES_NAME, dtype="U10")
        self._optimize = Optimize_SetupDT(0, nd, ng)


class MeshDT(_DT):
    def __init__(self, setup, nrow, ncol, ng):
        self.dx = np.float32(0.0)
        self.nrow, self.ncol, self.ng = int(nrow), int(ncol), int(ng)
        self.nac = 0
        self.xmin = 0
        self.ymax = 0
        self.flwdir = _farray((nrow, ncol), -99, np.int32)
        self.flwacc = _farray((nrow, ncol), -99, np.int32)
        self.path = _farray((2, nrow * ncol), -100, np.int32)  # 0-based view of the Fortran -99 fill
        self.active_cell = _farray((nrow, ncol), 1, np.int32)
        if ng > 0:
            self.flwdst = _farray((nrow, ncol), -99.0)
            self.gauge_pos = np.zeros((ng, 2), dtype=np.int32, order="F")
            self.code = np.array(["..."] * ng, dtype="U20")
            self.area = np.zeros(ng, dtype=np.float32)
        self._rowcol_to_ind_sparse = None
        self._local_active_cell = _farray((nrow, ncol), 1, np.int32)


class Input_DataDT(_DT):
    def __init__(self, setup, mesh):
        T = setup._ntime_step
        if mesh.ng > 0:
            self.qobs = _farray((mesh.ng, T), -99.0)
        if setup.sparse_storage:
            self.sparse_prcp = _farray((mesh.nac, T), -99.0)
            self.sparse_pet = _farray((mesh.nac, T), -99.0)
        else:
            self.prcp = _farray((mesh.nrow, mesh.ncol, T), -99.0)
            self.pet = _farray((mesh.nrow, mesh.ncol, T), -99.0)
        if setup._nd > 0:
            self.descriptor = _farray((mesh.nrow, mesh.ncol, setup._nd), 0.0)
        if setup.mean_forcing and mesh.ng > 0:
            self.mean_prcp = _farray((mesh.ng, T), -99.0)
            self.mean_pet = _farray((mesh.ng, T), -99.0)
        # smash_b200 extension: non-zero = "forcing arrays are immutable under this tag" (device copy is reused)
        self._forcing_version = 0


class ParametersDT(_DT):
    def __init__(self, mesh):
        for name in GPARAMETERS_NAME:
            setattr(self, name, _farray((mesh.nrow, mesh.ncol), _PARAMETERS_DEFAULT[name]))


class StatesDT(_DT):
    def __init__(self, mesh):
        for name in GSTATES_NAME:
            setattr(self, name, _farray((mesh.nrow, mesh.ncol), _STATES_DEFAULT[name]))


class Hyper_ParametersDT(_DT):
    def __init__(self, setup):
        nh = setup._optimize.nhyper
        for name in GPARAMETERS_NAME:
            setattr(self, name, _farray((nh, 1), 0.0))


class Hyper_StatesDT(_DT):
    def __init__(self, setup):
        nh = setup._optimize.nhyper
        for name in GSTATES_NAME:
            setattr(self, name, _farray((nh, 1), 0.0))


class OutputDT(_DT):
    def __init__(self, setup, mesh):
        T = setup._ntime_step
        if mesh.ng > 0:
            self.qsim = _farray((mesh.ng, T), -99.0)  # mwd_output.f90:76
        if setup.save_qsim_domain:
            if setup.sparse_storage:
                self.sparse_qsim_domain = _farray((mesh.nac, T), -99.0)
            else:
                self.qsim_domain = _farray((mesh.nrow, mesh.ncol, T), -99.0)
        if setup.save_net_prcp_domain:
            if setup.sparse_storage:
                self.sparse_net_prcp_domain = _farray((mesh.nac, T), -99.0)
            else:
                self.net_prcp_domain = _farray((mesh.nrow, mesh.ncol, T), -99.0)
        self.cost = np.float32(0.0)
        self.cost_jobs = np.float32(0.0)
        self.cost_jreg = np.float32(0.0)
        self._cost_jobs_initial = np.float32(0.0)
        self._cost_jreg_initial = np.float32(0.0)
        self.fstates = StatesDT(mesh)


def compute_rowcol_to_ind_sparse(mesh):
    """mw_sparse_storage.f90:12-49: sparse index k (1-based) of each active cell, in ``path`` order."""
    ind = np.zeros((mesh.nrow, mesh.ncol), dtype=np.int32, order="F")
    rows, cols = mesh.path[0], mesh.path[1]
    ok = (rows >= 0) & (cols >= 0)
    r, c = rows[ok], cols[ok]
    act = mesh.active_cell[r, c] == 1
    ind[r[act], c[act]] = np.arange(1, int(act.sum()) + 1, dtype=np.int32)
    mesh._rowcol_to_ind_sparse = ind
    return ind
