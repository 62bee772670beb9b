"""Drop-in for ``smash.solver._mw_derived_type_update`` (routine/mw_derived_type_update.f90:13-137)."""
from __future__ import annotations

import numpy as np

from ._derived_types import GLB_PARAMETERS, GLB_STATES, GUB_PARAMETERS, GUB_STATES


def reset_optimize_setup(this):
    """mw_derived_type_update.f90:13-59: every field back to its default, array shapes kept."""
    this.algorithm = "..."
    this.jobs_fun = np.array(["..."] * len(this.jobs_fun), dtype="U20")
    this.wjobs_fun = np.zeros(len(this.wjobs_fun), dtype=np.float32)
    this.wjreg = np.float32(0.0)
    this.jreg_fun = np.array(["..."] * len(this.jreg_fun), dtype="U20")
    this.wjreg_fun = np.ones(len(this.wjreg_fun), dtype=np.float32)
    this.reg_descriptors_for_params[...] = 0
    this.reg_descriptors_for_states[...] = 0
    this.njf = 0
    this.njr = 0
    this.verbose = True
    this.mapping = "..."
    this.denormalize_forward = False
    this.nhyper = 0
    this.optimize_start_step = 1
    this.maxiter = 100
    this.optim_parameters = np.zeros_like(this.optim_parameters)
    this.optim_states = np.zeros_like(this.optim_states)
    this.lb_parameters, this.ub_parameters = GLB_PARAMETERS.copy(), GUB_PARAMETERS.copy()
    this.lb_states, this.ub_states = GLB_STATES.copy(), GUB_STATES.copy()
    ng = len(this.wgauge)
    this.wgauge = np.full(ng, 1.0 / ng if ng else 0.0, dtype=np.float32)
    this.mask_event = np.zeros_like(this.mask_event)


def update_optimize_setup_optimize_args(this, mapping, ntime_step, nd, ng, njf):
    """mw_derived_type_update.f90:61-112."""
    this.mapping = mapping
    m = str(mapping).strip()
    if m == "hyper-linear":
        this.nhyper = 1 + int(nd)
    elif m == "hyper-polynomial":
        this.nhyper = 1 + 2 * int(nd)
    this.njf = int(njf)
    if this.njf != len(this.jobs_fun):
        this.jobs_fun = np.array(["..."] * this.njf, dtype="U20")
        this.wjobs_fun = np.zeros(this.njf, dtype=np.float32)
    if int(ng) != len(this.wgauge):
        this.wgauge = np.full(int(ng), 1.0 / ng if ng else 0.0, dtype=np.float32)
    if this.mask_event.shape != (int(ng), int(ntime_step)):
        this.mask_event = np.zeros((int(ng), int(ntime_step)), dtype=np.int32, order="F")


def update_optimize_setup_optimize_options(this, njr):
    """mw_derived_type_update.f90:114-135."""
    this.njr = int(njr)
    if this.njr != len(this.jreg_fun):
        this.jreg_fun = np.array(["..."] * this.njr, dtype="U20")
        this.wjreg_fun = np.ones(this.njr, dtype=np.float32)
