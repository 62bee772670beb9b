"""Drop-in for ``smash.solver._mw_derived_type_copy`` (routine/mw_derived_type_copy.f90:18-110): ``copy = this``."""
from __future__ import annotations

import copy as _copy


def _assign(this, copy):
    copy.__dict__.clear()
    copy.__dict__.update(_copy.deepcopy(this.__dict__))


copy_setup = copy_mesh = copy_input_data = copy_parameters = copy_states = copy_output = _assign
