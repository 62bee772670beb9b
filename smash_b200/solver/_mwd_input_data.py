"""``smash.solver._mwd_input_data`` (derived_type/mwd_input_data.f90)."""
from ._derived_types import Input_DataDT  # noqa: F401
