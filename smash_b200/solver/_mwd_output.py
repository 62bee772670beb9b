"""``smash.solver._mwd_output`` (derived_type/mwd_output.f90)."""
from ._derived_types import OutputDT  # noqa: F401
