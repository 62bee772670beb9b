"""``smash.solver._mwd_parameters`` (derived_type/mwd_parameters.f90)."""
from ._derived_types import Hyper_ParametersDT, ParametersDT  # noqa: F401
